// Weight gradient of a convolution on tcgen05 / TMEM, fed by TMA (sm_100a).
//
// The reference gets dL/dW of every convolution from ATen autograd over F.conv2d / F.conv_transpose2d
// (stylegan2.py:129,276,285,291 -> cuDNN wgrad).  Here it is one GEMM whose contraction runs over PIXELS:
//
//   dW[t][a][b] += sum_{(n,oy,ox)}  A[n, oy*sa + dya(t), ox*sa + dxa(t), a] * Bm[n, oy*sb + dyb(t), ox*sb + dxb(t), b]
//
// with both operands in the layout the forward / dgrad kernels already use: NHWC bf16.  A K chunk is 64 grid pixels
// (tw x th x tb, like an igemm tile); for every 64 channels of an operand ONE 4-D TMA box (64 ch x tw x th x tb, the
// tap's shift in the pixel coordinates, the conv stride as TMA element stride, zero padding = out-of-bounds fill) lands
// as 64 rows (pixels = K) of 128 bytes (channels = M or N): the canonical **MN-major** SWIZZLE_128B UMMA operand
// (cute::UMMA Layout_MN_SW128_Atom: 8 K-rows of 64 MN-elements per 1024-byte atom; stride between 8-row groups along
// K = SBO = 1024 B; stride between 64-channel boxes along M/N = LBO = 8192 B).  So the operands need no transposed copy:
// the instruction descriptor just marks A and B as MN-major.
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-5 epilogue (TMEM lane quarter =
// warp % 4).  Persistent CTAs over work items (m-tile, n-tile, tap group, k-slice); the fp32 accumulators of an item are
// added to dW with vector atomics (split-K over pixels: a 128 x 256 tile of one tap would otherwise be one SM's whole job).
//
// Tap groups: the taps of a conv shift only ONE operand (x); the other (dL/dy) is the same for all of them.  An item takes
// a group of T taps (T accumulators of BN columns in TMEM, T*BN <= 512): per 64-pixel chunk the shared operand is loaded
// once and only the T shifted tiles stream -- a one-tap item streams 16 KB + N*128 B per 2N tensor-pipe clocks (96 B/clk
// per SM at N = 256, 128 at N = 128, against the ~42 B/clk L2->SM budget of the chip), a group of T taps
// (16 KB + T*N*128 B) per 2*N*T clocks (80 at N = 256 / T = 2, 83 at N = 128 / T = 3).
//
// Algorithmic FLOPs per launch: 2 * B*GH*GW * Ca * Cb * ntaps.
#include <stdlib.h>

#include "common.cuh"

namespace fm {

constexpr int WG_BM = 128;
constexpr int WG_BK = 64;                 // pixels per K chunk
constexpr int WG_BOX_BYTES = 64 * 128;    // one TMA box: 64 pixels x 64 channels bf16
constexpr int WG_THREADS = 192;
constexpr int WG_MAX_STAGES = 8;

struct WgradParams {
  int Ca, Cb;
  int m_tiles, n_tiles, ntaps, ksplit, kper, nchunks, num_items;
  int tg, ngroups;                 // taps per item, number of tap groups
  int share_a;                     // 1: the A operand is common to the taps of a group (B tiles stream), 0: the B operand is
  int stages, stage_bytes, a_slot_bytes, b_slot_bytes, nbuf;
  int tw, th, tb, tiles_x, tiles_y;
  int sa, sb;
  float* dw;
  long long dw_tap_stride;
  int dw_row_stride;
  int8_t dya[FM_MAX_TAPS], dxa[FM_MAX_TAPS], dyb[FM_MAX_TAPS], dxb[FM_MAX_TAPS];
};

template <int BN>
struct WgradCfg {
  static constexpr int NB_BOXES = BN >= 64 ? BN / 64 : 1;
  static constexpr int A_BYTES = 2 * WG_BOX_BYTES;                 // 128 channels
  static constexpr int B_BYTES = NB_BOXES * WG_BOX_BYTES;
  static constexpr int RING_BYTES = 192 * 1024;
  static constexpr int SMEM_BYTES = RING_BYTES + 256 /*barriers*/ + 1024 /*alignment slack*/;
  static constexpr int TMEM_COLS = 512;    // up to T accumulators of BN columns (x2 when they fit twice)
};

// Instruction descriptor: D = f32, A = B = bf16, both MN-major (bits 15, 16), dense, M x N.
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int M, int N) {
  return umma_idesc_bf16(M, N) | (1u << 15) | (1u << 16);
}
// MN-major SWIZZLE_128B operand descriptor halves: LBO = 8192 B between 64-channel boxes, SBO = 1024 B between 8-pixel groups.
__device__ __forceinline__ uint32_t umma_desc_lo_mn(uint32_t smem_addr) { return ((smem_addr & 0x3FFFF) >> 4) | ((8192u >> 4) << 16); }

// The four K=16 steps of one 64-pixel chunk: both start addresses advance 16 pixel rows = 2048 bytes (+128 in the >>4 field).
__device__ __forceinline__ void umma_bf16_x4_mn(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "setp.eq.u32 q, 0, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "add.u32 al, %1, 128;\n\tadd.u32 bl, %3, 128;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t"
      "add.u32 al, %1, 256;\n\tadd.u32 bl, %3, 256;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t"
      "add.u32 al, %1, 384;\n\tadd.u32 bl, %3, 384;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int BN>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradParams p) {
  using Cfg = WgradCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::RING_BYTES);
  uint64_t* full_bar = bars;                         // [STAGES] TMA -> MMA
  uint64_t* empty_bar = bars + WG_MAX_STAGES;        // [STAGES] MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * WG_MAX_STAGES;    // [2] MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * WG_MAX_STAGES + 2;   // [2] epilogue -> MMA
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * WG_MAX_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < WG_MAX_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
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
  // (`tap` is the first tap of the item's group; the group holds min(tg, ntaps - tap) taps)
  auto decode = [&](int it, int& mt, int& nt, int& tap, int& c0, int& c1) {
    const int ks = it % p.ksplit; it /= p.ksplit;
    tap = (it % p.ngroups) * p.tg; it /= p.ngroups;
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
      const int nt_g = min(p.tg, p.ntaps - tap);           // taps in this group
      const int na = p.share_a ? 1 : nt_g, nb = p.share_a ? nt_g : 1;
      const uint32_t tx = static_cast<uint32_t>(na) * Cfg::A_BYTES + static_cast<uint32_t>(nb) * Cfg::B_BYTES;
      // chunk c -> tile (bx, by, bb) of the contraction grid
      int bx = c0 % p.tiles_x, t = c0 / p.tiles_x;
      int by = t % p.tiles_y, bb = t / p.tiles_y;
      for (int c = c0; c < c1; ++c) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* sa = smem + stage * p.stage_bytes;
          uint8_t* sb = sa + p.a_slot_bytes;
          const int gx = bx * p.tw, gy = by * p.th, gb = bb * p.tb;
          mbar_arrive_expect_tx(&full_bar[stage], tx);
          for (int i = 0; i < na; ++i) {
            const int dya = p.dya[tap + i], dxa = p.dxa[tap + i];
#pragma unroll
            for (int j = 0; j < 2; ++j)
              tma_load_4d(sa + i * Cfg::A_BYTES + j * WG_BOX_BYTES, &tmA, &full_bar[stage], mt * WG_BM + j * 64, gx * p.sa + dxa,
                          gy * p.sa + dya, gb);
          }
          for (int i = 0; i < nb; ++i) {
            const int dyb = p.dyb[tap + i], dxb = p.dxb[tap + i];
#pragma unroll
            for (int j = 0; j < Cfg::NB_BOXES; ++j)
              tma_load_4d(sb + i * Cfg::B_BYTES + j * WG_BOX_BYTES, &tmB, &full_bar[stage], nt * BN + j * 64, gx * p.sb + dxb,
                          gy * p.sb + dyb, gb);
          }
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
        if (++bx == p.tiles_x) { bx = 0; if (++by == p.tiles_y) { by = 0; ++bb; } }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    const uint32_t idesc = umma_idesc_bf16_mn(WG_BM, BN);
    constexpr uint32_t dhi = umma_desc_hi_sw128(1024);      // SBO = 1024 B
    const uint32_t ring = smem_u32(smem);
    int stage = 0, buf = 0;
    uint32_t phase = 0, aphase = 0;
    for (int it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      int mt, nt, tap, c0, c1;
      decode(it, mt, nt, tap, c0, c1);
      mbar_wait(&tempty_bar[buf], aphase ^ 1);
      tc_fence_after();
      const int nt_g = min(p.tg, p.ntaps - tap);
      const uint32_t tmem_d = tmem_base + buf * (p.tg * BN);
      for (int c = c0; c < c1; ++c) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = ring + stage * p.stage_bytes, sb = sa + p.a_slot_bytes;
        if (elect_one()) {
          for (int i = 0; i < nt_g; ++i)
            umma_bf16_x4_mn(tmem_d + i * BN, umma_desc_lo_mn(sa + (p.share_a ? 0 : i) * Cfg::A_BYTES), dhi,
                            umma_desc_lo_mn(sb + (p.share_a ? i : 0) * Cfg::B_BYTES), dhi, idesc, c > c0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (c == c1 - 1) umma_commit(&tfull_bar[buf]);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (++buf == p.nbuf) { buf = 0; aphase ^= 1; }
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
      const int nt_g = min(p.tg, p.ntaps - tap);
      mbar_wait(&tfull_bar[buf], aphase);
      tc_fence_after();
      for (int i = 0; i < nt_g; ++i) {
        float* base = p.dw + static_cast<long long>(tap + i) * p.dw_tap_stride + static_cast<long long>(a) * p.dw_row_stride + nt * BN;
        const uint32_t tmem_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * (p.tg * BN) + i * BN;
#pragma unroll 1
        for (int cb = 0; cb < BN; cb += 16) {
          uint32_t acc[16];
          tmem_ld_32x16(tmem_acc + cb, acc);
          tmem_ld_wait();
          if (a < p.Ca) {
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
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
      if (++buf == p.nbuf) { buf = 0; aphase ^= 1; }
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

static int encode_operand(EncodeTiledFn encode, CUtensorMap* tm, const fm_wgrad_operand& o, int B, int tw, int th, int tb, int s,
                          const char* which) {
  const cuuint64_t dims[4] = {static_cast<cuuint64_t>(o.C), static_cast<cuuint64_t>(o.W), static_cast<cuuint64_t>(o.H),
                              static_cast<cuuint64_t>(B)};
  const cuuint64_t strides[3] = {static_cast<cuuint64_t>(o.cstride) * 2, static_cast<cuuint64_t>(o.cstride) * o.W * 2,
                                 static_cast<cuuint64_t>(o.cstride) * o.W * o.H * 2};
  // with an element stride s TMA loads ceil(box/s) elements: box = n*s loads n
  const cuuint32_t box[4] = {64, static_cast<cuuint32_t>(tw * s), static_cast<cuuint32_t>(th * s), static_cast<cuuint32_t>(tb)};
  const cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(s), static_cast<cuuint32_t>(s), 1};
  CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(o.ptr), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("fm_conv_wgrad: cuTensorMapEncodeTiled(%s) failed with CUresult %d", which, (int)r); return FM_ERR_CUDA; }
  return FM_OK;
}

}  // namespace fm

using namespace fm;

extern "C" int fm_conv_wgrad(const fm_wgrad_desc* d, void* stream) {
  FM_CHECK_ARG(d != nullptr && d->a.ptr && d->b.ptr && d->dw, "fm_conv_wgrad: null pointer");
  FM_CHECK_ARG(d->B > 0 && d->GH > 0 && d->GW > 0 && d->ntaps >= 1 && d->ntaps <= FM_MAX_TAPS, "fm_conv_wgrad: bad sizes");
  for (const fm_wgrad_operand* o : {&d->a, &d->b}) {
    FM_CHECK_ARG(o->C > 0 && o->H > 0 && o->W > 0 && o->cstride >= o->C && o->cstride % 8 == 0,
                 "fm_conv_wgrad: operand channel stride must be a multiple of 8 and >= C");
    FM_CHECK_ARG(o->stride >= 1 && o->stride <= 2, "fm_conv_wgrad: operand stride must be 1 or 2");
    FM_CHECK_ARG((reinterpret_cast<uintptr_t>(o->ptr) & 15) == 0, "fm_conv_wgrad: operands must be 16-byte aligned");
  }
  FM_CHECK_ARG(d->dw_row_stride >= d->b.C, "fm_conv_wgrad: dw_row_stride < Cb");
  {
    // K chunks are 64-pixel tiles that may reach past the grid: one operand must BE the grid (same size, stride 1, no
    // shift), so that its out-of-bounds zero fill masks those pixels (for a conv: dL/dy)
    auto is_grid = [&](const fm_wgrad_operand& o, const int8_t* dy, const int8_t* dx) {
      if (o.H != d->GH || o.W != d->GW || o.stride != 1) return false;
      for (int i = 0; i < d->ntaps; ++i)
        if (dy[i] != 0 || dx[i] != 0) return false;
      return true;
    };
    FM_CHECK_ARG(is_grid(d->a, d->tap_dy_a, d->tap_dx_a) || is_grid(d->b, d->tap_dy_b, d->tap_dx_b),
                 "fm_conv_wgrad: one operand must be the contraction grid itself (H = GH, W = GW, stride 1, zero tap offsets)");
  }
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) { set_error("fm_conv_wgrad: cuTensorMapEncodeTiled driver entry point unavailable"); return FM_ERR_NO_DEVICE; }

  const int Cb = d->b.C;
  const int bn = Cb > 128 ? 256 : (Cb > 64 ? 128 : (Cb > 32 ? 64 : (Cb > 16 ? 32 : 16)));
  WgradParams p{};
  p.Ca = d->a.C; p.Cb = Cb;
  p.m_tiles = (p.Ca + WG_BM - 1) / WG_BM;
  p.n_tiles = (Cb + bn - 1) / bn;
  p.ntaps = d->ntaps;
  p.sa = d->a.stride; p.sb = d->b.stride;
  // K chunk = 64 pixels of the contraction grid [B][GH][GW]: tw x th x tb
  int tw = 1; while (tw < d->GW && tw < 64) tw <<= 1;
  int th = 1; while (th < d->GH && tw * th < 64) th <<= 1;
  const int tb = 64 / (tw * th);
  p.tw = tw; p.th = th; p.tb = tb;
  p.tiles_x = (d->GW + tw - 1) / tw;
  p.tiles_y = (d->GH + th - 1) / th;
  const int64_t nchunks = static_cast<int64_t>(p.tiles_x) * p.tiles_y * ((d->B + tb - 1) / tb);
  FM_CHECK_ARG(nchunks < 0x7FFFFFFF, "fm_conv_wgrad: too many K chunks");
  p.nchunks = static_cast<int>(nchunks);
  for (int i = 0; i < d->ntaps; ++i) {
    p.dya[i] = d->tap_dy_a[i]; p.dxa[i] = d->tap_dx_a[i]; p.dyb[i] = d->tap_dy_b[i]; p.dxb[i] = d->tap_dx_b[i];
  }
  // tap groups: the operand whose shift is the same for every tap (the grid operand, dL/dy) is loaded once per chunk
  {
    static const int env_tg = []() { const char* e = getenv("FM3D_WGRAD_TG"); return e ? atoi(e) : 0; }();
    bool same_a = true, same_b = true;
    for (int i = 1; i < d->ntaps; ++i) {
      same_a = same_a && p.dya[i] == p.dya[0] && p.dxa[i] == p.dxa[0];
      same_b = same_b && p.dyb[i] == p.dyb[0] && p.dxb[i] == p.dxb[0];
    }
    const int a_bytes = 2 * WG_BOX_BYTES, b_bytes = (bn >= 64 ? bn / 64 : 1) * WG_BOX_BYTES;
    int tg = 1;
    p.share_a = same_a ? 1 : 0;
    if (d->ntaps > 1 && (same_a || same_b)) {
      const int streamed = same_a ? b_bytes : a_bytes, shared = same_a ? a_bytes : b_bytes;
      tg = 512 / bn;                                              // T accumulators of BN columns
      while (tg > 1 && shared + tg * streamed > 96 * 1024) --tg;  // at least two stages in the 192 KB ring
      if (tg > d->ntaps) tg = d->ntaps;
      // balanced groups: 9 taps as 3 x 3 rather than 4 + 4 + 1
      const int ng = (d->ntaps + tg - 1) / tg;
      tg = (d->ntaps + ng - 1) / ng;
      if (env_tg > 0 && env_tg < tg) tg = env_tg;
    }
    p.tg = tg;
    p.ngroups = (d->ntaps + tg - 1) / tg;
    p.a_slot_bytes = (p.share_a ? 1 : tg) * a_bytes;
    p.b_slot_bytes = (p.share_a ? tg : 1) * b_bytes;
    p.stage_bytes = p.a_slot_bytes + p.b_slot_bytes;
    p.stages = (192 * 1024) / p.stage_bytes;
    if (p.stages > WG_MAX_STAGES) p.stages = WG_MAX_STAGES;
    p.nbuf = (2 * tg * bn <= 512) ? 2 : 1;
  }
  // split-K over pixels: aim at >= 2 items per SM, keep >= 8 chunks per item
  const int64_t items0 = static_cast<int64_t>(p.m_tiles) * p.n_tiles * p.ngroups;
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
  FM_CHECK_ARG(items < 0x7FFFFFFF, "fm_conv_wgrad: too many work items");
  p.num_items = static_cast<int>(items);
  p.dw = d->dw; p.dw_tap_stride = d->dw_tap_stride; p.dw_row_stride = d->dw_row_stride;

  CUtensorMap tmA, tmB;
  int rc = encode_operand(encode, &tmA, d->a, d->B, tw, th, tb, d->a.stride, "A");
  if (rc != FM_OK) return rc;
  rc = encode_operand(encode, &tmB, d->b, d->B, tw, th, tb, d->b.stride, "B");
  if (rc != FM_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (bn) {
    case 16: return launch_wgrad<16>(tmA, tmB, p, st);
    case 32: return launch_wgrad<32>(tmA, tmB, p, st);
    case 64: return launch_wgrad<64>(tmA, tmB, p, st);
    case 128: return launch_wgrad<128>(tmA, tmB, p, st);
    default: return launch_wgrad<256>(tmA, tmB, p, st);
  }
}
