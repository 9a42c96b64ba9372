// Weight-gradient GEMM on tcgen05 / TMEM, fed by TMA (sm_100a), and the layout pass that feeds it.
//
// The reference gets dL/dW of every convolution from ATen autograd over F.conv2d / F.conv_transpose2d
// (stylegan2.py:129,276,285,291 -> cuDNN wgrad).  Here the gradient of a conv weight is a plain GEMM whose
// contraction runs over PIXELS:
//
//   dW[t][a][b] += sum_{l < L}  A[slab_a(t)*Ca + a][l + off_a(t)] * Bm[slab_b(t)*Cb + b][l + off_b(t)]
//
// with both operands stored channel-major / pixel-linear ("CPL": [channels][B*Hq*Wq] bf16, rows 16-byte aligned), so a
// K chunk of 64 pixels of 128 (or N) channels is one 2-D TMA box that lands as a K-major SWIZZLE_128B UMMA operand.
// All geometry lives in the layout pass: the conv's zero padding is a zero halo in the pixel grid, a conv tap is a
// linear offset (ky*Wq + kx) into that grid, a stride-2 conv reads one of four parity planes (slabs) of its input, and
// out-of-range coordinates are TMA zero fill.  One kernel therefore serves conv2d (stride 1/2) and conv_transpose2d.
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-5 epilogue (TMEM lane quarter =
// warp % 4).  Persistent CTAs over work items (m-tile, n-tile, tap, k-slice); the fp32 accumulator of an item is added to
// dW with vector atomics (split-K over pixels: a 128 x 256 tile of one tap would otherwise be the whole job of one SM).
//
// Algorithmic FLOPs per launch: 2 * L_valid * Ca * Cb * ntaps.
#include <stdlib.h>

#include "common.cuh"

namespace fm {

constexpr int WG_BM = 128;
constexpr int WG_BK = 64;
constexpr int WG_THREADS = 192;
constexpr int WG_MAX_STAGES = 8;

struct WgradParams {
  int Ca, Cb;
  int m_tiles, n_tiles, ntaps, ksplit, kper, nchunks, num_items;
  int stages;
  float* dw;
  long long dw_tap_stride;
  int dw_row_stride;
  int32_t off_a[FM_MAX_TAPS], off_b[FM_MAX_TAPS];
  int8_t slab_a[FM_MAX_TAPS], slab_b[FM_MAX_TAPS];
};

template <int BN>
struct WgradCfg {
  static constexpr int A_BYTES = WG_BM * WG_BK * 2;     // 16 KB
  static constexpr int B_BYTES = BN * WG_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (192 * 1024 / STAGE_BYTES) > WG_MAX_STAGES ? WG_MAX_STAGES : (192 * 1024 / STAGE_BYTES);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 /*barriers*/ + 1024 /*alignment slack*/;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
};

template <int BN>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradParams p) {
  using Cfg = WgradCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                         // [STAGES] TMA -> MMA
  uint64_t* empty_bar = bars + WG_MAX_STAGES;        // [STAGES] MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * WG_MAX_STAGES;    // [2] MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * WG_MAX_STAGES + 2;   // [2] epilogue -> MMA
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * WG_MAX_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < Cfg::STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(s_tmem, Cfg::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  pdl_wait();
  pdl_launch_dependents();

  // item -> (m-tile, n-tile, tap, k-slice); k-slice fastest so that the CTAs working on one dW tile run concurrently
  auto decode = [&](int it, int& mt, int& nt, int& tap, int& c0, int& c1) {
    const int ks = it % p.ksplit; it /= p.ksplit;
    tap = it % p.ntaps; it /= p.ntaps;
    nt = it % p.n_tiles;
    mt = it / p.n_tiles;
    c0 = ks * p.kper;
    c1 = min(p.nchunks, c0 + p.kper);
  };

  if (warp == 0) {
    // ============================== TMA producer ==============================
    int stage = 0;
    uint32_t phase = 0;
    for (int it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      int mt, nt, tap, c0, c1;
      decode(it, mt, nt, tap, c0, c1);
      const int ra = p.slab_a[tap] * p.Ca + mt * WG_BM, rb = p.slab_b[tap] * p.Cb + nt * BN;
      const int oa = p.off_a[tap], ob = p.off_b[tap];
      for (int c = c0; c < c1; ++c) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[stage], c * WG_BK + oa, ra);
          tma_load_2d(sa + Cfg::A_BYTES, &tmB, &full_bar[stage], c * WG_BK + ob, rb);
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    const uint32_t idesc = umma_idesc_bf16(WG_BM, BN);
    constexpr uint32_t dhi = umma_desc_hi_sw128(1024);
    const uint32_t ring = smem_u32(smem);
    int stage = 0, buf = 0;
    uint32_t phase = 0, aphase = 0;
    for (int it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      int mt, nt, tap, c0, c1;
      decode(it, mt, nt, tap, c0, c1);
      mbar_wait(&tempty_bar[buf], aphase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + buf * BN;
      for (int c = c0; c < c1; ++c) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = ring + stage * Cfg::STAGE_BYTES;
        if (elect_one()) {
          umma_bf16_x4(tmem_d, umma_desc_lo(sa), dhi, umma_desc_lo(sa + Cfg::A_BYTES), dhi, idesc, c > c0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (c == c1 - 1) umma_commit(&tfull_bar[buf]);
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
      if (++buf == 2) { buf = 0; aphase ^= 1; }
    }
  } else {
    // ============================== epilogue (4 warps) ==============================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int buf = 0;
    uint32_t aphase = 0;
    const bool vec_ok = (p.dw_row_stride & 3) == 0 && (p.dw_tap_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dw) & 15) == 0;
    for (int it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      int mt, nt, tap, c0, c1;
      decode(it, mt, nt, tap, c0, c1);
      const int a = mt * WG_BM + row;
      float* base = p.dw + static_cast<long long>(tap) * p.dw_tap_stride + static_cast<long long>(a) * p.dw_row_stride + nt * BN;
      mbar_wait(&tfull_bar[buf], aphase);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN;
#pragma unroll 1
      for (int cb = 0; cb < BN; cb += 16) {
        uint32_t acc[16];
        tmem_ld_32x16(tmem_acc + cb, acc);
        tmem_ld_wait();
        if (c1 > c0 && a < p.Ca) {
          const int b0 = nt * BN + cb;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            if (vec_ok && b0 + j + 3 < p.Cb) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(base + cb + j), "f"(__uint_as_float(acc[j])),
                           "f"(__uint_as_float(acc[j + 1])), "f"(__uint_as_float(acc[j + 2])), "f"(__uint_as_float(acc[j + 3]))
                           : "memory");
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (b0 + j + e < p.Cb) atomicAdd(base + cb + j + e, __uint_as_float(acc[j + e]));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
      if (++buf == 2) { buf = 0; aphase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN>
static int launch_wgrad(const CUtensorMap& tmA, const CUtensorMap& tmB, const WgradParams& p, cudaStream_t st) {
  using Cfg = WgradCfg<BN>;
  static SmemOptIn opt_in;
  FM_CUDA_OK(smem_opt_in(opt_in, wgrad_kernel<BN>, Cfg::SMEM_BYTES));
  const int sms = sm_count();
  const unsigned grid = static_cast<unsigned>(p.num_items < sms ? p.num_items : sms);
  FM_CUDA_OK(launch_pdl(wgrad_kernel<BN>, dim3(grid), dim3(WG_THREADS), Cfg::SMEM_BYTES, st, tmA, tmB, p));
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

// dst[(slab*C + c)][b*Hq*Wq + yq*Wq + xq] = scale[b,c] * src[b, c, yq*s + py - y0, xq*s + px - x0]   (0 outside the image),
// slab = py*s + px.  One thread per 8 consecutive xq (one 16-byte store).
__global__ void __launch_bounds__(256) nchw_to_cpl_kernel(__nv_bfloat16* __restrict__ dst, const float* __restrict__ src,
                                                         const float* __restrict__ scale, int B, int C, int H, int W, int s,
                                                         int y0, int x0, int Hq, int Wq, int64_t total8) {
  pdl_wait();
  pdl_launch_dependents();
  const int wq8 = Wq >> 3;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total8; idx += stride) {
    int64_t t = idx;
    const int xg = static_cast<int>(t % wq8); t /= wq8;
    const int yq = static_cast<int>(t % Hq); t /= Hq;
    const int b = static_cast<int>(t % B); t /= B;
    const int c = static_cast<int>(t % C);
    const int slab = static_cast<int>(t / C);
    const int py = slab / s, px = slab - py * s;
    const int y = yq * s + py - y0;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (y >= 0 && y < H) {
      const float sc = scale ? __ldg(scale + static_cast<int64_t>(b) * C + c) : 1.f;
      const float* row = src + ((static_cast<int64_t>(b) * C + c) * H + y) * W;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int x = (xg * 8 + j) * s + px - x0;
        if (x >= 0 && x < W) v[j] = __ldg(row + x) * sc;
      }
    }
    uint4 w;
    w.x = pack_bf16x2(v[0], v[1]); w.y = pack_bf16x2(v[2], v[3]);
    w.z = pack_bf16x2(v[4], v[5]); w.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(dst + idx * 8) = w;
  }
}

}  // namespace fm

using namespace fm;

extern "C" int fm_nchw_to_cpl_bf16(void* dst, const float* src, const float* scale_bc, int B, int C, int H, int W, int s, int y0,
                                   int x0, int Hq, int Wq, void* stream) {
  FM_CHECK_ARG(dst && src && B > 0 && C > 0 && H > 0 && W > 0, "fm_nchw_to_cpl_bf16: bad args");
  FM_CHECK_ARG((s == 1 || s == 2) && Hq > 0 && Wq > 0 && Wq % 8 == 0, "fm_nchw_to_cpl_bf16: s must be 1 or 2, Wq a multiple of 8");
  FM_CHECK_ARG((reinterpret_cast<uintptr_t>(dst) & 15) == 0, "fm_nchw_to_cpl_bf16: dst must be 16-byte aligned");
  const int64_t total8 = static_cast<int64_t>(s) * s * C * B * Hq * (Wq / 8);
  int64_t blocks = (total8 + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  FM_CUDA_OK(launch_pdl(nchw_to_cpl_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream),
                        static_cast<__nv_bfloat16*>(dst), src, scale_bc, B, C, H, W, s, y0, x0, Hq, Wq, total8));
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

extern "C" int fm_wgrad_gemm(const fm_wgrad_desc* d, void* stream) {
  FM_CHECK_ARG(d != nullptr && d->a && d->b && d->dw, "fm_wgrad_gemm: null pointer");
  FM_CHECK_ARG(d->Ca > 0 && d->Cb > 0 && d->L > 0 && d->ntaps >= 1 && d->ntaps <= FM_MAX_TAPS, "fm_wgrad_gemm: bad sizes");
  FM_CHECK_ARG(d->La % 8 == 0 && d->Lb % 8 == 0 && d->La > 0 && d->Lb > 0, "fm_wgrad_gemm: row lengths must be multiples of 8 elements");
  FM_CHECK_ARG(d->La < 0x7FFFFF00LL && d->Lb < 0x7FFFFF00LL, "fm_wgrad_gemm: row length exceeds the TMA coordinate range");
  FM_CHECK_ARG((reinterpret_cast<uintptr_t>(d->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->b) & 15) == 0,
               "fm_wgrad_gemm: operands must be 16-byte aligned");
  FM_CHECK_ARG(d->nslabs_a >= 1 && d->nslabs_b >= 1 && d->dw_row_stride >= d->Cb, "fm_wgrad_gemm: bad slab counts / dw_row_stride");
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) { set_error("fm_wgrad_gemm: cuTensorMapEncodeTiled driver entry point unavailable"); return FM_ERR_NO_DEVICE; }

  int bn = d->Cb > 128 ? 256 : (d->Cb > 64 ? 128 : (d->Cb > 32 ? 64 : (d->Cb > 16 ? 32 : 16)));
  WgradParams p{};
  p.Ca = d->Ca; p.Cb = d->Cb;
  p.m_tiles = (d->Ca + WG_BM - 1) / WG_BM;
  p.n_tiles = (d->Cb + bn - 1) / bn;
  p.ntaps = d->ntaps;
  p.nchunks = static_cast<int>((d->L + WG_BK - 1) / WG_BK);
  for (int i = 0; i < d->ntaps; ++i) {
    FM_CHECK_ARG(d->tap_slab_a[i] >= 0 && d->tap_slab_a[i] < d->nslabs_a && d->tap_slab_b[i] >= 0 && d->tap_slab_b[i] < d->nslabs_b,
                 "fm_wgrad_gemm: tap %d: slab out of range", i);
    p.off_a[i] = d->tap_off_a[i]; p.off_b[i] = d->tap_off_b[i];
    p.slab_a[i] = d->tap_slab_a[i]; p.slab_b[i] = d->tap_slab_b[i];
  }
  // split-K over pixels: aim at >= 2 items per SM, keep >= 8 chunks per item
  const int64_t items0 = static_cast<int64_t>(p.m_tiles) * p.n_tiles * p.ntaps;
  int ksplit = d->ksplit;
  if (ksplit <= 0) {
    const int sms = sm_count();
    ksplit = static_cast<int>((2 * static_cast<int64_t>(sms) + items0 - 1) / items0);
    const int max_split = p.nchunks / 8 > 0 ? p.nchunks / 8 : 1;
    if (ksplit > max_split) ksplit = max_split;
    if (ksplit < 1) ksplit = 1;
  }
  if (ksplit > p.nchunks) ksplit = p.nchunks;
  p.kper = (p.nchunks + ksplit - 1) / ksplit;
  p.ksplit = (p.nchunks + p.kper - 1) / p.kper;
  const int64_t items = items0 * p.ksplit;
  FM_CHECK_ARG(items < 0x7FFFFFFF, "fm_wgrad_gemm: too many work items");
  p.num_items = static_cast<int>(items);
  p.dw = d->dw; p.dw_tap_stride = d->dw_tap_stride; p.dw_row_stride = d->dw_row_stride;

  CUtensorMap tmA, tmB;
  {
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(d->La), static_cast<cuuint64_t>(d->nslabs_a) * d->Ca};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(d->La) * 2};
    const cuuint32_t box[2] = {WG_BK, WG_BM};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d->a), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("fm_wgrad_gemm: cuTensorMapEncodeTiled(A) failed with CUresult %d", (int)r); return FM_ERR_CUDA; }
  }
  {
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(d->Lb), static_cast<cuuint64_t>(d->nslabs_b) * d->Cb};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(d->Lb) * 2};
    const cuuint32_t box[2] = {WG_BK, static_cast<cuuint32_t>(bn)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d->b), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("fm_wgrad_gemm: cuTensorMapEncodeTiled(B) failed with CUresult %d", (int)r); return FM_ERR_CUDA; }
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (bn) {
    case 16: return launch_wgrad<16>(tmA, tmB, p, st);
    case 32: return launch_wgrad<32>(tmA, tmB, p, st);
    case 64: return launch_wgrad<64>(tmA, tmB, p, st);
    case 128: return launch_wgrad<128>(tmA, tmB, p, st);
    default: return launch_wgrad<256>(tmA, tmB, p, st);
  }
}
