// Bandwidth-bound helper kernels of the synthesis path (sm_100a): style affine, epilogue
// tables (demodulation), layout converts, the fused blur+noise+bias+lrelu pass after the
// stride-2 transposed conv, the ToRGB tail and weight preparation.
#include <type_traits>

#include "common.cuh"

namespace fm {

// ------------------------------------------------------------------------------------
// Small "NT" GEMMs of the modulation path:  C[b][o] = sum_k A[b][k] * W[o][k]  with B <= a few
// dozen samples.  lane = sample; the A rows of a 32-sample chunk sit in shared memory (padded, so
// the per-lane float4 reads are conflict-free); W values are warp-uniform broadcast loads.  Each
// warp produces 4 output channels per pass: no shuffles, 1 LDS.128 + 4 LDG.128 per 16 FMAs.
//   style_affine : A = latent[:, idx, :], W = modulation weight  (stylegan2.py:165-175,240,257)
//   build_tables : A = s^2,               W = wsq                 (stylegan2.py:258-262)
// grid (ceil(max_out/SG_CH), n_layers), 256 threads, dynamic smem 32*(K+4)*4 bytes.  The op is latency
// bound (a dependent chain of K/4 steps per warp), so CTAs are kept small: 32 output channels each
// gives ~4 CTAs per SM for the generator's 20 layers.
// ------------------------------------------------------------------------------------
constexpr int SG_CH = 32;   // output channels per CTA (8 warps x 4)
template <bool SQUARE>
__device__ __forceinline__ void stage_rows(float* s_a, const float* __restrict__ a, int64_t row_stride, int b0, int B, int K,
                                           int KP) {
  // s_a[r][k] for r = 0..31 (sample b0 + r), zero beyond B and beyond K
  for (int i = threadIdx.x; i < 32 * KP; i += blockDim.x) {
    const int r = i / KP, k = i - r * KP;
    float v = 0.f;
    if (b0 + r < B && k < K) v = __ldg(a + static_cast<int64_t>(b0 + r) * row_stride + k);
    s_a[i] = SQUARE ? v * v : v;
  }
}

__device__ __forceinline__ void dot4x4(const float* s_row, const float* __restrict__ w, int K, int o_valid, float (&acc)[4]) {
  // acc[c] = sum_k s_row[k] * w[c*K + k], c < o_valid (rows beyond are clamped by the caller)
  if ((K & 3) == 0) {
    // The W rows are warp-uniform.  Loading them as uniform (broadcast) LDG.128s fetched 16 bytes per instruction and
    // kept only a few hundred bytes in flight per warp: the kernels ran at 260 GB/s, bound by load latency.  Now
    // lane l fetches quad l of a 128-element k-block of each row (512 coalesced bytes per instruction, the next block
    // already in flight) and the quads are handed round with shuffles.
    const int lane = threadIdx.x & 31;
    const float* wr[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) wr[c] = w + static_cast<size_t>(min(c, o_valid - 1)) * K;
    auto fetch = [&](int kb, float4 (&dst)[4]) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        dst[c] = kb + 4 * lane < K ? __ldg(reinterpret_cast<const float4*>(wr[c] + kb + 4 * lane)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    float4 cur[4], nxt[4] = {};
    fetch(0, cur);
    for (int kb = 0; kb < K; kb += 128) {
      if (kb + 128 < K) fetch(kb + 128, nxt);
      const int nq = min(32, (K - kb) >> 2);
#pragma unroll 4
      for (int q = 0; q < nq; ++q) {
        const float4 av = *reinterpret_cast<const float4*>(s_row + kb + 4 * q);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float wx = __shfl_sync(0xffffffffu, cur[c].x, q), wy = __shfl_sync(0xffffffffu, cur[c].y, q);
          const float wz = __shfl_sync(0xffffffffu, cur[c].z, q), ww = __shfl_sync(0xffffffffu, cur[c].w, q);
          acc[c] = fmaf(av.x, wx, fmaf(av.y, wy, fmaf(av.z, wz, fmaf(av.w, ww, acc[c]))));
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) cur[c] = nxt[c];
    }
  } else {
    for (int k = 0; k < K; ++k) {
      const float av = s_row[k];
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[c] = fmaf(av, __ldg(w + static_cast<size_t>(min(c, o_valid - 1)) * K + k), acc[c]);
    }
  }
}

__global__ void __launch_bounds__(256) style_affine_kernel(const fm_style_layer* __restrict__ layers,
                                                           const float* __restrict__ latent, int B, int n_latent,
                                                           int D, float scale) {
  extern __shared__ __align__(16) float s_a[];
  const fm_style_layer L = layers[blockIdx.y];
  if (blockIdx.x * SG_CH >= L.cin) return;
  const int KP = ((D + 3) & ~3) + 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b0 = 0; b0 < B; b0 += 32) {
    __syncthreads();
    stage_rows<false>(s_a, latent + static_cast<int64_t>(L.latent_idx) * D, static_cast<int64_t>(n_latent) * D, b0, B, D, KP);
    __syncthreads();
#pragma unroll 1
    for (int pass = 0; pass < SG_CH / 32; ++pass) {
      const int o = blockIdx.x * SG_CH + warp * (SG_CH / 8) + pass * 4;
      if (o >= L.cin) break;
      const int nv = min(4, L.cin - o);
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      dot4x4(s_a + lane * KP, L.wmod + static_cast<size_t>(o) * D, D, nv, acc);
      if (b0 + lane < B) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < nv) L.s[static_cast<size_t>(b0 + lane) * L.cin + o + c] = fmaf(acc[c], scale, __ldg(L.bmod + o + c));
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// Epilogue tables: T[b][o] = (d, bias, slope, gain*s_next, rgb weights...)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) build_tables_kernel(const fm_table_layer* __restrict__ layers, int B) {
  extern __shared__ __align__(16) float s_a[];
  const fm_table_layer L = layers[blockIdx.y];
  if (blockIdx.x * SG_CH >= L.cout) return;
  const int KP = ((L.cin + 3) & ~3) + 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float rgb_scale = rsqrtf(static_cast<float>(L.cout));   // ToRGB fan_in = cout*1*1 (stylegan2.py:232-233)
  for (int b0 = 0; b0 < B; b0 += 32) {
    if (L.wsq) {
      __syncthreads();
      stage_rows<true>(s_a, L.s, L.cin, b0, B, L.cin, KP);
      __syncthreads();
    }
#pragma unroll 1
    for (int pass = 0; pass < SG_CH / 32; ++pass) {
      const int o = blockIdx.x * SG_CH + warp * (SG_CH / 8) + pass * 4;
      if (o >= L.cout) break;
      const int nv = min(4, L.cout - o);
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      if (L.wsq) dot4x4(s_a + lane * KP, L.wsq + static_cast<size_t>(o) * L.cin, L.cin, nv, acc);
      const int b = b0 + lane;
      if (b >= B) continue;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c >= nv) break;
        const int oc = o + c;
        float* t = L.tab + (static_cast<size_t>(b) * L.cout + oc) * 8;
        float4 t0, t1 = make_float4(0.f, 0.f, 0.f, 0.f);
        t0.x = L.wsq ? rsqrtf(acc[c] + 1e-8f) : 1.f;                   // stylegan2.py:261 (literal 1e-8)
        t0.y = L.act_bias ? __ldg(L.act_bias + oc) : 0.f;
        t0.z = L.slope;
        t0.w = L.gain * (L.s_next ? __ldg(L.s_next + static_cast<size_t>(b) * L.cout + oc) : 1.f);
        if (L.wrgb) {
          const float sr = __ldg(L.s_rgb + static_cast<size_t>(b) * L.cout + oc) * L.gain;
          t1.x = rgb_scale * __ldg(L.wrgb + oc) * sr;
          t1.y = rgb_scale * __ldg(L.wrgb + L.cout + oc) * sr;
          t1.z = rgb_scale * __ldg(L.wrgb + 2 * L.cout + oc) * sr;
        }
        *reinterpret_cast<float4*>(t) = t0;
        *reinterpret_cast<float4*>(t + 4) = t1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// layout converts
// ------------------------------------------------------------------------------------
// fp32 NCHW -> bf16 NHWC through a shared-memory tile of 64 pixels x 64 channels: the loads are coalesced along the
// pixels of a channel plane, the stores along the channels of a pixel (a warp writes whole 128-byte lines; the first
// version stored 16 bytes per thread 2*cs bytes apart -- half-used sectors, and this pass is the most-launched kernel of
// a training iteration: every differentiable conv converts its operands).  grid = (pixel tiles, channel tiles, B).
constexpr int NT_PIX = 64, NT_CH = 64, NT_ROW = NT_CH * 2 + 8;   // row pitch 136 B: 8-byte reads stay aligned, 2-way conflicts at most
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(__nv_bfloat16* __restrict__ out, const float* __restrict__ x,
                                                           const float* __restrict__ scale, int C, int HW, int cs) {
  __shared__ __align__(16) uint8_t tile[NT_PIX * NT_ROW];
  const int p0 = blockIdx.x * NT_PIX, c0 = blockIdx.y * NT_CH;
  const int64_t b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // All 16 + 8 loads of the thread are issued before the first is used.  Written as one loop (load, scale, store to the
  // tile per channel) ptxas kept the iterations in order -- eight waits on DRAM per CTA, 3 loads in flight per thread:
  // ncu put 72 % of the stall samples on the eight multiplies and the pass moved 4.0 TB/s (Little's law: 24 KB in flight
  // per SM).
  float v0[NT_CH / 8], v1[NT_CH / 8], sc[NT_CH / 8];
#pragma unroll
  for (int k = 0; k < NT_CH / 8; ++k) {
    const int c = c0 + warp + 8 * k;
    const bool ok = c < C;
    const float* src = x + (b * C + (ok ? c : 0)) * HW;
    sc[k] = (ok && scale) ? __ldg(scale + b * C + c) : 1.f;
    v0[k] = (ok && p0 + lane < HW) ? __ldg(src + p0 + lane) : 0.f;
    v1[k] = (ok && p0 + lane + 32 < HW) ? __ldg(src + p0 + lane + 32) : 0.f;
  }
#pragma unroll
  for (int k = 0; k < NT_CH / 8; ++k) {
    const int cl = warp + 8 * k;
    *reinterpret_cast<__nv_bfloat16*>(tile + lane * NT_ROW + cl * 2) = __float2bfloat16_rn(v0[k] * sc[k]);
    *reinterpret_cast<__nv_bfloat16*>(tile + (lane + 32) * NT_ROW + cl * 2) = __float2bfloat16_rn(v1[k] * sc[k]);
  }
  __syncthreads();
  // 64 pixels x 16 quads of 4 channels (8 bytes): consecutive threads write consecutive quads of a pixel
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int id = threadIdx.x + 256 * k;
    const int pl = id >> 4, q = id & 15;
    const int ch = c0 + q * 4;
    if (p0 + pl < HW && ch < cs)
      *reinterpret_cast<uint2*>(out + (b * HW + p0 + pl) * cs + ch) = *reinterpret_cast<const uint2*>(tile + pl * NT_ROW + q * 8);
  }
}

__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(float* __restrict__ out, const __nv_bfloat16* __restrict__ x,
                                                           const float* __restrict__ inv_scale, int C, int HW, int cs,
                                                           int64_t total) {
  const int groups = (C + 7) / 8;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int pix = static_cast<int>(idx % HW);
    const int64_t t = idx / HW;
    const int g = static_cast<int>(t % groups);
    const int64_t b = t / groups;
    const uint4 w = __ldg(reinterpret_cast<const uint4*>(x + (b * HW + pix) * cs + g * 8));
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      if (c < C) {
        const float2 f = unpack_bf16x2(ws[j >> 1]);
        float v = (j & 1) ? f.y : f.x;
        if (inv_scale) v /= __ldg(inv_scale + b * C + c);
        out[(b * C + c) * HW + pix] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// blur (4x4 FIR, pad (1,1)) + demod + noise + bias + lrelu + next style, NHWC bf16.
// t [B, OH+1, OW+1, cs] -> out [B, OH, OW, cs].  Replaces Blur (stylegan2.py:279 ->
// upfirdn2d mode 1), NoiseInjection (:312) and FusedLeakyReLU (:371) after the stride-2
// transposed conv.
// The op moves 4 B per output element and needs ~11 fp32 operations for it, so it is bound by HBM
// only if (a) enough bytes are in flight and (b) the instruction count is cut to the bone:
//  (a) a CTA owns a 64-column x 64-channel strip and streams down BL_ROWS output rows; TMA
//      (cp.async.bulk.tensor, out-of-bounds = the blur's zero padding) stages 2 input rows
//      (67 px x 128 B each) per mbarrier stage into a 4-stage ring: ~50 KB of loads in flight per
//      CTA from ONE issuing thread, nothing is fetched twice.
//  (b) a thread owns 4 adjacent columns x 4 channels: 7 LDS.64 feed 4 outputs, the horizontal pass
//      and the rolling window of 4 vertical partial sums (statically renamed, 4 rows per unrolled
//      step) run on packed fp32 pairs (FFMA2), the epilogue on FFMA2/FMUL2/FMNMX.
// Rank-1 kernels (the default [1,3,3,1] outer product) take the separable path.
// Algorithmic bytes per output pixel-channel: 2 B read ((OH+1)(OW+1)/(OH*OW) ~ 1) + 2 B write.
// ------------------------------------------------------------------------------------
// A CTA owns CB channels x TW columns with CB * TW = 4096: CB = 128 when the tensor has at least 128 channels, so a
// staged pixel is 256 contiguous bytes (two CTAs no longer split every 256-byte pixel of the 128-channel layers between
// them -- DRAM sees whole rows of the strip), else CB = 64.
template <int CB> struct BlurCfg {
  static constexpr int TW = 4096 / CB;             // output columns per CTA
  static constexpr int IW = TW + 3;                // staged input columns
  static constexpr int PIX = CB * 2;               // bytes per staged pixel
  static constexpr int STAGE_BYTES = 2 * IW * PIX; // BL_SR rows
  static constexpr int SMEM = 4 * STAGE_BYTES;     // BL_NST stages
};
constexpr int BL_TW = 64;                        // (CB = 64) output columns per CTA
constexpr int BL_IW = BL_TW + 3;                 // staged input columns
constexpr int BL_SR = 2;                         // input rows per stage
constexpr int BL_NST = 4;                        // ring stages
constexpr int BL_STAGE_BYTES = BL_SR * BL_IW * 128;
constexpr int BL_ROWS = 64;                      // default output rows per strip (see fm_blur_act_nhwc)
constexpr int BL_SMEM = BL_NST * BL_STAGE_BYTES;

template <bool SEP, int CB>
__global__ void __launch_bounds__(256, 2) blur_act_nhwc_kernel(const __grid_constant__ CUtensorMap tmT,
                                                               __nv_bfloat16* __restrict__ out,
                                                               const float* __restrict__ kernel, const float* __restrict__ tab,
                                                               const float* __restrict__ noise, int noise_bstride,
                                                               const float* __restrict__ noise_w, int OH, int OW, int C, int cs,
                                                               int tiles_x, int tiles_y, int cblocks, int strip_rows) {
  extern __shared__ __align__(128) uint8_t bl_smem[];
  __shared__ __align__(8) uint64_t s_full[BL_NST];
  __shared__ float s_k[16];
  __shared__ float s_kv[4], s_kh[4];
  pdl_wait();
  pdl_launch_dependents();
  int bid = blockIdx.x;
  const int cb = bid % cblocks; bid /= cblocks;
  const int tx = bid % tiles_x; bid /= tiles_x;
  const int ty = bid % tiles_y;
  const int b = bid / tiles_y;
  using BC = BlurCfg<CB>;
  const int c0 = cb * CB, x0 = tx * BC::TW;
  const int Y0 = ty * strip_rows;
  const int Y1 = min(OH, Y0 + strip_rows);
  const int tid = threadIdx.x;
  if (tid < 16) {
    const int a = tid >> 2, bb = tid & 3;
    s_k[tid] = kernel[(3 - a) * 4 + (3 - bb)];   // flipped taps (true convolution)
  }
  if (tid < 4) {
    // rank-1 kernel: k[a][b] = rowsum[a] * colsum[b] / total
    float tot = 0.f;
    for (int i = 0; i < 16; ++i) tot += kernel[i];
    float rs = 0.f, csum = 0.f;
    for (int j = 0; j < 4; ++j) { rs += kernel[(3 - tid) * 4 + j]; csum += kernel[j * 4 + (3 - tid)]; }
    s_kv[tid] = rs;
    s_kh[tid] = csum / tot;
  }
  const int nrows_in = (Y1 - Y0) + 3;                       // input rows Y0-1 .. Y1+1
  const int nstages = (nrows_in + BL_SR - 1) / BL_SR;
  auto issue = [&](int s) {                                  // thread 0 only
    const int slot = s % BL_NST;
    fence_proxy_async();                                     // generic reads of this slot precede the async refill
    mbar_arrive_expect_tx(&s_full[slot], BC::STAGE_BYTES);
    tma_load_4d(bl_smem + slot * BC::STAGE_BYTES, &tmT, &s_full[slot], c0, x0 - 1, Y0 - 1 + s * BL_SR, b);
  };
  if (tid == 0) {
    tma_prefetch_desc(&tmT);
    for (int i = 0; i < BL_NST; ++i) mbar_init(&s_full[i], 1);
    fence_barrier_init();
    for (int s = 0; s < BL_NST && s < nstages; ++s) issue(s);
  }
  __syncthreads();

  const int g = tid & (CB / 4 - 1);       // 4-channel group inside the channel block
  const int cg = tid / (CB / 4);          // 4-column group inside the column tile
  const int X = x0 + 4 * cg;
  const int ch = c0 + g * 4;
  const bool active = X < OW && ch < cs;
  const float nw = noise ? (noise_w ? __ldg(noise_w) : 1.f) : 0.f;
  const float* nzp = noise ? noise + static_cast<size_t>(noise_bstride ? b : 0) * OH * OW + X : nullptr;
  const bool nz_vec = (OW & 3) == 0;

  // epilogue table of this thread's 4 channels as pairs: v = acc*d + bias (+noise); lrelu; * post
  f32x2 td[2], tb[2], ts[2], tp[2];
#pragma unroll
  for (int pr = 0; pr < 2; ++pr) {
    float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
    if (ch + 2 * pr < C) t0 = __ldg(reinterpret_cast<const float4*>(tab + (static_cast<size_t>(b) * C + ch + 2 * pr) * 8));
    if (ch + 2 * pr + 1 < C) t1 = __ldg(reinterpret_cast<const float4*>(tab + (static_cast<size_t>(b) * C + ch + 2 * pr + 1) * 8));
    td[pr] = f2_pack(t0.x, t1.x); tb[pr] = f2_pack(t0.y, t1.y); ts[pr] = f2_pack(t0.z, t1.z); tp[pr] = f2_pack(t0.w, t1.w);
  }
  f32x2 kh2[4], kv2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { kh2[i] = f2_pack(s_kh[i], s_kh[i]); kv2[i] = f2_pack(s_kv[i], s_kv[i]); }

  // Input row j (relative, iy = Y0-1+j) feeds output row Y0 + j - a with vertical tap a; its partial sum
  // sits in slot (j - a) & 3, and tap 3 closes output row Y0 + j - 3 (slot (j + 1) & 3).
  f32x2 acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[i][c][0] = acc[i][c][1] = 0ull;

  const uint8_t* my = bl_smem + (4 * cg) * BC::PIX + g * 8;
  __nv_bfloat16* orow0 = out + (static_cast<size_t>(b) * OH * OW + X) * cs + ch;

  // The strip's noise ((Y1 - Y0) rows x TW columns of fp32) is staged in shared memory once, with every load in flight
  // at the same time, while the first TMA stages land.  Requested row by row (even one row ahead) it was an L2 round
  // trip per row on the critical path: ncu put a third of the kernel's stall samples on the first use of that register.
  float* s_nz = reinterpret_cast<float*>(bl_smem + BC::SMEM);
  const bool nz_smem = noise != nullptr && nz_vec;
  if (nz_smem) {
    constexpr int Q = BC::TW / 4;
    const int n4 = (Y1 - Y0) * Q;
    const float* nbase = noise + static_cast<size_t>(noise_bstride ? b : 0) * OH * OW;
    for (int i = tid; i < n4; i += 256) {
      const int r = i / Q, q = i - r * Q;
      float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (x0 + 4 * q < OW) v4 = __ldg(reinterpret_cast<const float4*>(nbase + static_cast<size_t>(Y0 + r) * OW + x0 + 4 * q));
      *reinterpret_cast<float4*>(s_nz + r * BC::TW + 4 * q) = v4;
    }
    // visible to the other warps after the first per-stage __syncthreads below, two input rows before the first output row
  }

  auto load_noise = [&](int Y) -> float4 {
    float4 n4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nzp && active) {
      const float* np = nzp + static_cast<size_t>(Y) * OW;
      if (nz_vec) {
        n4 = __ldg(reinterpret_cast<const float4*>(np));
      } else {
        if (X + 0 < OW) n4.x = __ldg(np + 0);
        if (X + 1 < OW) n4.y = __ldg(np + 1);
        if (X + 2 < OW) n4.z = __ldg(np + 2);
        if (X + 3 < OW) n4.w = __ldg(np + 3);
      }
    }
    return n4;
  };
  auto do_row = [&](auto uc, int j, const uint8_t* rowp) {
    constexpr int u = decltype(uc)::value;
    f32x2 v[7][2];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const uint2 w = *reinterpret_cast<const uint2*>(rowp + i * BC::PIX);
      v[i][0] = f2_from_bf16x2(w.x);
      v[i][1] = f2_from_bf16x2(w.y);
    }
    if (SEP) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
          f32x2 h = f2_mul(kh2[0], v[c][pr]);
          h = f2_fma(kh2[1], v[c + 1][pr], h);
          h = f2_fma(kh2[2], v[c + 2][pr], h);
          h = f2_fma(kh2[3], v[c + 3][pr], h);
          acc[u & 3][c][pr] = f2_mul(kv2[0], h);            // tap 0 opens a new output row: no zeroing pass
#pragma unroll
          for (int a = 1; a < 4; ++a) acc[(u - a) & 3][c][pr] = f2_fma(kv2[a], h, acc[(u - a) & 3][c][pr]);
        }
    } else {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float kk = s_k[a * 4 + q];
          const f32x2 k2 = f2_pack(kk, kk);
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int pr = 0; pr < 2; ++pr)
              acc[(u - a) & 3][c][pr] = (a == 0 && q == 0) ? f2_mul(k2, v[c + q][pr]) : f2_fma(k2, v[c + q][pr], acc[(u - a) & 3][c][pr]);
        }
    }
    const int Y = Y0 + j - 3;
    if (j >= 3 && Y < Y1) {
      // (widths that are not a multiple of 4 -- no generator resolution -- load the row's noise here, unvectorised)
      const float4 npre = nz_smem ? *reinterpret_cast<const float4*>(s_nz + (Y - Y0) * BC::TW + 4 * cg) : load_noise(Y);
      const float nz[4] = {npre.x * nw, npre.y * nw, npre.z * nw, npre.w * nw};
      __nv_bfloat16* orow = orow0 + static_cast<size_t>(Y) * OW * cs;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t o2[2];
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
          f32x2 x = f2_fma(acc[(u + 1) & 3][c][pr], td[pr], f2_add(tb[pr], f2_pack(nz[c], nz[c])));
          const f32x2 xs = f2_mul(x, ts[pr]);
          float xl, xh, sl, sh;
          f2_unpack(x, xl, xh);
          f2_unpack(xs, sl, sh);
          // x > 0 ? x : x*slope
          const f32x2 y = f2_mul(f2_pack(xl > 0.f ? xl : sl, xh > 0.f ? xh : sh), tp[pr]);
          float yl, yh;
          f2_unpack(y, yl, yh);
          o2[pr] = pack_bf16x2(yl, yh);
        }
        if (X + c < OW) *reinterpret_cast<uint2*>(orow + static_cast<size_t>(c) * cs) = make_uint2(o2[0], o2[1]);
      }
    }
  };

  // two stages (4 input rows) per iteration keep the slot renaming static
  for (int s = 0; s < nstages; s += 2) {
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      const int sc = s + h2;
      if (sc >= nstages) break;
      const int slot = sc % BL_NST;
      mbar_wait(&s_full[slot], (sc / BL_NST) & 1);
      if (active) {
        const uint8_t* base = my + slot * BC::STAGE_BYTES;
        const int j0 = sc * BL_SR;
        if (h2 == 0) {
          do_row(std::integral_constant<int, 0>{}, j0, base);
          if (j0 + 1 < nrows_in) do_row(std::integral_constant<int, 1>{}, j0 + 1, base + BC::IW * BC::PIX);
        } else {
          do_row(std::integral_constant<int, 2>{}, j0, base);
          if (j0 + 1 < nrows_in) do_row(std::integral_constant<int, 3>{}, j0 + 1, base + BC::IW * BC::PIX);
        }
      }
      __syncthreads();                                   // the slot is drained
      if (tid == 0 && sc + BL_NST < nstages) issue(sc + BL_NST);
    }
  }
}

// ------------------------------------------------------------------------------------
// ToRGB tail: rgb_out = acc + bias + Upsample(skip)   (stylegan2.py:394-399; Upsample =
// upfirdn2d(up=2, pad=(2,1)) with kernel*4, stylegan2.py:52-63).  One thread per pixel.
// ------------------------------------------------------------------------------------
// A thread owns 4 horizontally adjacent pixels: 4 LDG.128 of the accumulator (re-zeroed for the next
// forward), the 4x2 skip samples per plane its 2x-upsampled footprint needs, 3 STG.128 (one per plane).
__global__ void __launch_bounds__(256) rgb_finalize_kernel(float* __restrict__ rgb_out, float* __restrict__ acc,
                                                           const float* __restrict__ bias3, const float* __restrict__ skip,
                                                           const float* __restrict__ kernel, int H, int W, int total_quads) {
  __shared__ float s_k[16];
  if (threadIdx.x < 16) {
    const int a = threadIdx.x >> 2, b = threadIdx.x & 3;
    s_k[threadIdx.x] = kernel ? kernel[(3 - a) * 4 + (3 - b)] : 0.f;
  }
  __syncthreads();
  const int qw = W >> 2;
  const int h2 = H >> 1, w2 = W >> 1;
  const float bias[3] = {__ldg(bias3 + 0), __ldg(bias3 + 1), __ldg(bias3 + 2)};
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total_quads; idx += gridDim.x * blockDim.x) {
    const int xq = idx % qw;
    const int t = idx / qw;
    const int y = t % H;
    const int b = t / H;
    const int x0 = xq * 4;
    float4* ap = reinterpret_cast<float4*>(acc) + (static_cast<size_t>(t) * W + x0);
    float v[3][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 a = ap[j];
      ap[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      v[0][j] = a.x + bias[0]; v[1][j] = a.y + bias[1]; v[2][j] = a.z + bias[2];
    }
    if (skip) {
      // zero-stuffed row u = y + ta - 2 is live when even: ta has the parity of y -> 2 of the 4 taps per axis;
      // pixel x0+j reads skip columns (x0+j + tb - 2) / 2 with tb = (j & 1) + 2*ib -> columns x0/2 - 1 .. x0/2 + 2
      const size_t plane = static_cast<size_t>(h2) * w2;
      const float* sp = skip + static_cast<size_t>(b) * 3 * plane;
      const int sx0 = (x0 >> 1) - 1;
#pragma unroll
      for (int ia = 0; ia < 2; ++ia) {
        const int ta = (y & 1) + 2 * ia;
        const int sy = (y + ta - 2) >> 1;
        if (y + ta - 2 < 0 || sy >= h2) continue;
        float kx[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) kx[i] = s_k[ta * 4 + i];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float* row = sp + c * plane + static_cast<size_t>(sy) * w2;
          float sv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) sv[i] = (sx0 + i >= 0 && sx0 + i < w2) ? __ldg(row + sx0 + i) : 0.f;
          // j even: tb in {0,2} -> columns (j/2 - 1, j/2) + x0/2 ; j odd: tb in {1,3} -> ((j-1)/2, (j+1)/2) + x0/2
          v[c][0] = fmaf(sv[0], kx[0], fmaf(sv[1], kx[2], v[c][0]));
          v[c][1] = fmaf(sv[1], kx[1], fmaf(sv[2], kx[3], v[c][1]));
          v[c][2] = fmaf(sv[1], kx[0], fmaf(sv[2], kx[2], v[c][2]));
          v[c][3] = fmaf(sv[2], kx[1], fmaf(sv[3], kx[3], v[c][3]));
        }
      }
    }
    const size_t hw = static_cast<size_t>(H) * W;
    float* op = rgb_out + static_cast<size_t>(b) * 3 * hw + static_cast<size_t>(y) * W + x0;
#pragma unroll
    for (int c = 0; c < 3; ++c) *reinterpret_cast<float4*>(op + c * hw) = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
  }
}

// ------------------------------------------------------------------------------------
// weight prep: [cout][cin][kh][kw] fp32 -> bf16 [tap][cout_rows][cin_stride] (scaled), and
// wsq[cout][cin] = sum_taps (scale*W)^2 for the demodulation table.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prep_weight_kernel(__nv_bfloat16* __restrict__ wq, float* __restrict__ wsq,
                                                          const float* __restrict__ w, int cout, int cin, int taps,
                                                          float scale, int cout_rows, int cin_stride, int64_t total) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int i = static_cast<int>(idx % cin_stride);
    const int o = static_cast<int>(idx / cin_stride);   // 0..cout_rows-1
    float sq = 0.f;
    for (int t = 0; t < taps; ++t) {
      float v = 0.f;
      if (o < cout && i < cin) v = w[(static_cast<size_t>(o) * cin + i) * taps + t] * scale;
      sq = fmaf(v, v, sq);
      wq[(static_cast<size_t>(t) * cout_rows + o) * cin_stride + i] = __float2bfloat16_rn(v);
    }
    if (wsq && o < cout && i < cin) wsq[static_cast<size_t>(o) * cin + i] = sq;
  }
}

static inline unsigned grid_for(int64_t total, int per_block = 256, int waves = 16) {
  int64_t want = (total + per_block - 1) / per_block;
  const int64_t cap = static_cast<int64_t>(sm_count()) * waves;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return static_cast<unsigned>(want);
}

}  // namespace fm

using namespace fm;

static int small_gemm_smem(const void* fn, int K) {
  const int bytes = 32 * (((K + 3) & ~3) + 4) * 4;
  if (bytes > 48 * 1024) cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  return bytes;
}

extern "C" int fm_style_affine(const fm_style_layer* layers_dev, int n_layers, int max_cin, const float* latent, int B,
                               int n_latent, int style_dim, void* stream) {
  FM_CHECK_ARG(layers_dev && latent && n_layers > 0 && max_cin > 0 && B > 0 && style_dim > 0, "fm_style_affine: bad args");
  FM_CHECK_ARG(style_dim <= 1700, "fm_style_affine: style_dim %d too large for the shared-memory tile", style_dim);
  const int smem = small_gemm_smem(reinterpret_cast<const void*>(style_affine_kernel), style_dim);
  dim3 grid((max_cin + SG_CH - 1) / SG_CH, n_layers);
  style_affine_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(layers_dev, latent, B, n_latent, style_dim,
                                                                              1.0f / sqrtf(static_cast<float>(style_dim)));
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

extern "C" int fm_build_tables(const fm_table_layer* layers_dev, int n_layers, int max_cout, int max_cin, int B,
                               void* stream) {
  FM_CHECK_ARG(layers_dev && n_layers > 0 && max_cout > 0 && max_cin > 0 && B > 0, "fm_build_tables: bad args");
  FM_CHECK_ARG(max_cin <= 1700, "fm_build_tables: cin %d too large for the shared-memory tile", max_cin);
  const int smem = small_gemm_smem(reinterpret_cast<const void*>(build_tables_kernel), max_cin);
  dim3 grid((max_cout + SG_CH - 1) / SG_CH, n_layers);
  build_tables_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(layers_dev, B);
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

extern "C" int fm_nchw_to_nhwc_bf16(void* out, const float* x, const float* scale_bc, int B, int C, int H, int W,
                                    int out_cstride, void* stream) {
  FM_CHECK_ARG(out && x && B > 0 && C > 0 && H > 0 && W > 0, "fm_nchw_to_nhwc_bf16: bad args");
  FM_CHECK_ARG(out_cstride % 8 == 0 && out_cstride >= C, "fm_nchw_to_nhwc_bf16: cstride must be a multiple of 8 >= C");
  const int HW = H * W;
  FM_CHECK_ARG(B <= 65535 && (out_cstride + NT_CH - 1) / NT_CH <= 65535, "fm_nchw_to_nhwc_bf16: batch / channel count exceeds the grid");
  const dim3 grid(static_cast<unsigned>((HW + NT_PIX - 1) / NT_PIX), static_cast<unsigned>((out_cstride + NT_CH - 1) / NT_CH),
                  static_cast<unsigned>(B));
  nchw_to_nhwc_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<__nv_bfloat16*>(out), x, scale_bc, C, HW,
                                                                        out_cstride);
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

extern "C" int fm_nhwc_bf16_to_nchw(float* out, const void* x, const float* inv_scale_bc, int B, int C, int H, int W,
                                    int x_cstride, void* stream) {
  FM_CHECK_ARG(out && x && B > 0 && C > 0 && H > 0 && W > 0, "fm_nhwc_bf16_to_nchw: bad args");
  FM_CHECK_ARG(x_cstride % 8 == 0 && x_cstride >= C, "fm_nhwc_bf16_to_nchw: cstride must be a multiple of 8 >= C");
  const int64_t total = static_cast<int64_t>(B) * ((C + 7) / 8) * H * W;
  nhwc_to_nchw_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      out, static_cast<const __nv_bfloat16*>(x), inv_scale_bc, C, H * W, x_cstride, total);
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

extern "C" int fm_blur_act_nhwc(void* out, const void* t, const float* kernel4x4, const float* tab, const float* noise,
                                int noise_bstride, const float* noise_w, int B, int OH, int OW, int C, int cstride,
                                int separable, int t_pitch_h, int t_pitch_w, void* stream) {
  FM_CHECK_ARG(out && t && kernel4x4 && tab && B > 0 && OH > 0 && OW > 0 && C > 0, "fm_blur_act_nhwc: bad args");
  FM_CHECK_ARG(cstride % 8 == 0 && cstride >= C, "fm_blur_act_nhwc: cstride must be a multiple of 8 >= C");
  FM_CHECK_ARG((t_pitch_h == 0 || t_pitch_h >= OH + 1) && (t_pitch_w == 0 || t_pitch_w >= OW + 1),
               "fm_blur_act_nhwc: t_pitch_h / t_pitch_w must be >= OH+1 / OW+1");
  FM_CHECK_ARG((reinterpret_cast<uintptr_t>(t) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "fm_blur_act_nhwc: tensors must be 16-byte aligned");
  // strip height: the grid is not persistent, so pick the height whose CTA count wastes the least of the last wave
  // (2 CTAs per SM); shorter strips re-read 3 halo rows more often, but those hit L2 (neighbouring strips run together)
  static const int env_rows = []() { const char* e = getenv("FM3D_BLUR_ROWS"); return e ? atoi(e) : 0; }();
  static const int env_cb = []() { const char* e = getenv("FM3D_BLUR_CB"); return e ? atoi(e) : 0; }();
  const int CBv = env_cb == 64 || env_cb == 128 ? env_cb : (cstride % 128 == 0 ? 128 : 64);
  const int TWv = 4096 / CBv;
  const int tiles_x = (OW + TWv - 1) / TWv, cblocks = (cstride + CBv - 1) / CBv;
  int strip_rows = BL_ROWS;
  if (env_rows > 0) {
    strip_rows = env_rows;
  } else {
    const int64_t slots = static_cast<int64_t>(sm_count()) * 2;
    double best = -1.0;
    for (int r : {128, 64, 32, 16}) {
      const int64_t n = static_cast<int64_t>(B) * tiles_x * ((OH + r - 1) / r) * cblocks;
      const int64_t waves = (n + slots - 1) / slots;
      // useful fraction: wave fill x (1 - halo overhead)
      const double eff = static_cast<double>(n) / (waves * slots) * r / (r + 3.0);
      if (eff > best + 1e-9) { best = eff; strip_rows = r; }
    }
  }
  const int tiles_y = (OH + strip_rows - 1) / strip_rows;
  const int64_t blocks = static_cast<int64_t>(B) * tiles_x * tiles_y * cblocks;
  FM_CHECK_ARG(blocks < 0x7FFFFFFF, "fm_blur_act_nhwc: too many blocks");
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) { set_error("fm_blur_act_nhwc: cuTensorMapEncodeTiled driver entry point unavailable"); return FM_ERR_NO_DEVICE; }
  CUtensorMap tmT;
  {
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(cstride), static_cast<cuuint64_t>(OW + 1), static_cast<cuuint64_t>(OH + 1),
                                static_cast<cuuint64_t>(B)};
    const cuuint64_t pw = t_pitch_w > 0 ? t_pitch_w : OW + 1, ph = t_pitch_h > 0 ? t_pitch_h : OH + 1;
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(cstride) * 2, static_cast<cuuint64_t>(cstride) * 2 * pw,
                                   static_cast<cuuint64_t>(cstride) * 2 * pw * ph};
    const cuuint32_t box[4] = {static_cast<cuuint32_t>(CBv), static_cast<cuuint32_t>(TWv + 3), BL_SR, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tmT, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(t), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("fm_blur_act_nhwc: cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return FM_ERR_CUDA; }
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static SmemOptIn opt_in[4];     // per kernel instantiation, per device
  // + the strip's noise: at most 128 rows x TW columns of fp32
  FM_CUDA_OK(smem_opt_in(opt_in[0], blur_act_nhwc_kernel<true, 64>, BlurCfg<64>::SMEM + 128 * BlurCfg<64>::TW * 4));
  FM_CUDA_OK(smem_opt_in(opt_in[1], blur_act_nhwc_kernel<false, 64>, BlurCfg<64>::SMEM + 128 * BlurCfg<64>::TW * 4));
  FM_CUDA_OK(smem_opt_in(opt_in[2], blur_act_nhwc_kernel<true, 128>, BlurCfg<128>::SMEM + 128 * BlurCfg<128>::TW * 4));
  FM_CUDA_OK(smem_opt_in(opt_in[3], blur_act_nhwc_kernel<false, 128>, BlurCfg<128>::SMEM + 128 * BlurCfg<128>::TW * 4));
  FM_CHECK_ARG(strip_rows <= 128, "fm_blur_act_nhwc: FM3D_BLUR_ROWS must be <= 128");
  auto fn = CBv == 128 ? (separable ? blur_act_nhwc_kernel<true, 128> : blur_act_nhwc_kernel<false, 128>)
                       : (separable ? blur_act_nhwc_kernel<true, 64> : blur_act_nhwc_kernel<false, 64>);
  FM_CUDA_OK(launch_pdl(fn, dim3(static_cast<unsigned>(blocks)), dim3(256), (CBv == 128 ? BlurCfg<128>::SMEM : BlurCfg<64>::SMEM) + (noise ? strip_rows * TWv * 4 : 0), st,
                        tmT, static_cast<__nv_bfloat16*>(out), kernel4x4, tab, noise, noise_bstride, noise_w,
                        OH, OW, C, cstride, tiles_x, tiles_y, cblocks, strip_rows));
  count_launch();
  return FM_OK;
}

extern "C" int fm_rgb_finalize(float* rgb_out, float* acc, const float* bias3, const float* skip, const float* kernel4x4,
                               int B, int H, int W, void* stream) {
  FM_CHECK_ARG(rgb_out && acc && bias3 && B > 0 && H > 0 && W > 0, "fm_rgb_finalize: bad args");
  FM_CHECK_ARG(!skip || (kernel4x4 && H % 2 == 0 && W % 2 == 0), "fm_rgb_finalize: skip needs a kernel and even H, W");
  FM_CHECK_ARG(W % 4 == 0, "fm_rgb_finalize: W must be a multiple of 4");
  const int64_t total = static_cast<int64_t>(B) * H * (W / 4);
  FM_CHECK_ARG(total < 0x7FFFFFFF, "fm_rgb_finalize: too many pixels");
  rgb_finalize_kernel<<<grid_for(total, 256, 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      rgb_out, acc, bias3, skip, kernel4x4, H, W, static_cast<int>(total));
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

extern "C" int fm_prep_weight(void* wq, float* wsq, const float* w_oikk, int cout, int cin, int kh, int kw, float scale,
                              int cout_rows, int cin_stride, void* stream) {
  FM_CHECK_ARG(wq && w_oikk && cout > 0 && cin > 0 && kh > 0 && kw > 0, "fm_prep_weight: bad args");
  FM_CHECK_ARG(cout_rows >= cout && cin_stride >= cin && cin_stride % 8 == 0, "fm_prep_weight: bad padded sizes");
  const int64_t total = static_cast<int64_t>(cout_rows) * cin_stride;
  prep_weight_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<__nv_bfloat16*>(wq), wsq, w_oikk, cout, cin, kh * kw, scale, cout_rows, cin_stride, total);
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}
