// Bandwidth-bound helper kernels of the synthesis path (sm_100a): style affine, epilogue
// tables (demodulation), layout converts, the fused blur+noise+bias+lrelu pass after the
// stride-2 transposed conv, the ToRGB tail and weight preparation.
#include "common.cuh"

namespace fm {

// ------------------------------------------------------------------------------------
// Small "NT" GEMMs of the modulation path:  C[b][o] = sum_k A[b][k] * W[o][k]  with B <= a few
// dozen samples.  lane = sample; the A rows of a 32-sample chunk sit in shared memory (padded, so
// the per-lane float4 reads are conflict-free); W values are warp-uniform broadcast loads.  Each
// warp produces 4 output channels per pass: no shuffles, 1 LDS.128 + 4 LDG.128 per 16 FMAs.
//   style_affine : A = latent[:, idx, :], W = modulation weight  (stylegan2.py:165-175,240,257)
//   build_tables : A = s^2,               W = wsq                 (stylegan2.py:258-262)
// grid (ceil(max_out/64), n_layers), 256 threads, dynamic smem 32*(K+4)*4 bytes.
// ------------------------------------------------------------------------------------
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
#pragma unroll 8
    for (int k = 0; k < K; k += 4) {      // unrolled: 32 independent broadcast loads in flight per thread
      const float4 av = *reinterpret_cast<const float4*>(s_row + k);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + static_cast<size_t>(min(c, o_valid - 1)) * K + k));
        acc[c] = fmaf(av.x, wv.x, fmaf(av.y, wv.y, fmaf(av.z, wv.z, fmaf(av.w, wv.w, acc[c]))));
      }
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
  if (blockIdx.x * 64 >= L.cin) return;
  const int KP = ((D + 3) & ~3) + 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b0 = 0; b0 < B; b0 += 32) {
    __syncthreads();
    stage_rows<false>(s_a, latent + static_cast<int64_t>(L.latent_idx) * D, static_cast<int64_t>(n_latent) * D, b0, B, D, KP);
    __syncthreads();
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      const int o = blockIdx.x * 64 + warp * 8 + pass * 4;
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
  if (blockIdx.x * 64 >= L.cout) return;
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
    for (int pass = 0; pass < 2; ++pass) {
      const int o = blockIdx.x * 64 + warp * 8 + pass * 4;
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
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(__nv_bfloat16* __restrict__ out, const float* __restrict__ x,
                                                           const float* __restrict__ scale, int C, int HW, int cs,
                                                           int64_t total) {
  const int groups = cs / 8;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int pix = static_cast<int>(idx % HW);
    const int64_t t = idx / HW;
    const int g = static_cast<int>(t % groups);
    const int64_t b = t / groups;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      float f = 0.f;
      if (c < C) {
        f = x[(b * C + c) * HW + pix];
        if (scale) f *= __ldg(scale + b * C + c);
      }
      v[j] = f;
    }
    uint4 w;
    w.x = pack_bf16x2(v[0], v[1]); w.y = pack_bf16x2(v[2], v[3]);
    w.z = pack_bf16x2(v[4], v[5]); w.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + (b * HW + pix) * cs + g * 8) = w;
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
// One CTA = 16x16 output pixels x 64 channels: the 19x19x128 B input patch is staged in
// shared memory with 16-byte coalesced loads; a thread owns 8 channels of a vertical strip of
// 8 outputs and walks the 11 input rows it needs (4 LDS.128 per row), so every staged value
// is reused from registers.  Rank-1 kernels (the default [1,3,3,1] outer product) take the
// separable path: half the FMAs.
// Algorithmic bytes per output pixel-channel: 2 B read ((OH+1)(OW+1)/(OH*OW) ~ 1) + 2 B write.
// ------------------------------------------------------------------------------------
constexpr int BL_T = 16;            // output tile edge
constexpr int BL_P = BL_T + 3;      // patch edge

template <bool SEP>
__global__ void __launch_bounds__(512, 2) blur_act_nhwc_kernel(__nv_bfloat16* __restrict__ out,
                                                               const __nv_bfloat16* __restrict__ t,
                                                               const float* __restrict__ kernel, const float* __restrict__ tab,
                                                               const float* __restrict__ noise, int noise_bstride,
                                                               const float* __restrict__ noise_w, int OH, int OW, int C, int cs,
                                                               int tiles_x, int tiles_y, int cblocks) {
  __shared__ __align__(16) uint4 s_patch[BL_P * BL_P * 8];   // [row][px][64 ch]
  __shared__ float4 s_tab[64];
  __shared__ float s_k[16];
  __shared__ float s_kv[4], s_kh[4];
  int bid = blockIdx.x;
  const int cb = bid % cblocks; bid /= cblocks;
  const int tx = bid % tiles_x; bid /= tiles_x;
  const int ty = bid % tiles_y;
  const int b = bid / tiles_y;
  const int X0 = tx * BL_T, Y0 = ty * BL_T, c0 = cb * 64;
  const int IH = OH + 1, IW = OW + 1;
  const int tid = threadIdx.x;
  if (tid < 16) {
    const int a = tid >> 2, bb = tid & 3;
    s_k[tid] = kernel[(3 - a) * 4 + (3 - bb)];   // flipped taps (true convolution)
  }
  if (SEP && tid < 4) {
    // rank-1 kernel: k[a][b] = rowsum[a] * colsum[b] / total
    float tot = 0.f;
    for (int i = 0; i < 16; ++i) tot += kernel[i];
    float rs = 0.f, csum = 0.f;
    for (int j = 0; j < 4; ++j) { rs += kernel[(3 - tid) * 4 + j]; csum += kernel[j * 4 + (3 - tid)]; }
    s_kv[tid] = rs;
    s_kh[tid] = csum / tot;
  }
  if (tid < 64) {
    const int o = c0 + tid;
    s_tab[tid] = o < C ? __ldg(reinterpret_cast<const float4*>(tab + (static_cast<size_t>(b) * C + o) * 8))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // ---- stage the patch (rows Y0-1 .. Y0+17, cols X0-1 .. X0+17), zero outside the image
  const __nv_bfloat16* tb = t + static_cast<size_t>(b) * IH * IW * cs;
  for (int i = tid; i < BL_P * BL_P * 8; i += 512) {
    const int g = i & 7;
    const int px = (i >> 3) % BL_P, py = (i >> 3) / BL_P;
    const int iy = Y0 - 1 + py, ix = X0 - 1 + px;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < IH && ix >= 0 && ix < IW && c0 + g * 8 < cs)
      v = __ldg(reinterpret_cast<const uint4*>(tb + (static_cast<size_t>(iy) * IW + ix) * cs + c0 + g * 8));
    s_patch[i] = v;
  }
  __syncthreads();

  // thread = 4 channels x one column x a strip of 8 rows  (512 threads: 16 ch-quads x 16 cols x 2 strips)
  const int g = tid & 15;
  const int sx = (tid >> 4) & 15;
  const int sy = (tid >> 8) * 8;
  const uint2* patch2 = reinterpret_cast<const uint2*>(s_patch);
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;

#pragma unroll
  for (int r = 0; r < 11; ++r) {        // patch rows sy + r feed outputs sy + r - a, a = 0..3
    float v[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint2 w = patch2[((sy + r) * BL_P + sx + q) * 16 + g];
      v[q][0] = __uint_as_float(w.x << 16); v[q][1] = __uint_as_float(w.x & 0xffff0000u);
      v[q][2] = __uint_as_float(w.y << 16); v[q][3] = __uint_as_float(w.y & 0xffff0000u);
    }
    if (SEP) {
      float hrow[4];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        hrow[c] = fmaf(s_kh[3], v[3][c], fmaf(s_kh[2], v[2][c], fmaf(s_kh[1], v[1][c], s_kh[0] * v[0][c])));
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int i = r - a;
        if (i < 0 || i > 7) continue;
        const float kv = s_kv[a];
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] = fmaf(kv, hrow[c], acc[i][c]);
      }
    } else {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int i = r - a;
        if (i < 0 || i > 7) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float kv = s_k[a * 4 + q];
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[i][c] = fmaf(kv, v[q][c], acc[i][c]);
        }
      }
    }
  }

  if (c0 + g * 4 >= cs) return;
  const int X = X0 + sx;
  if (X >= OW) return;
  const float nw = noise ? (noise_w ? __ldg(noise_w) : 1.f) : 0.f;
  float4 tv[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) tv[c] = s_tab[g * 4 + c];
  float nzv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int Y = Y0 + sy + i;
    nzv[i] = (noise && Y < OH) ? nw * __ldg(noise + (static_cast<size_t>(noise_bstride ? b : 0) * OH + Y) * OW + X) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int Y = Y0 + sy + i;
    if (Y >= OH) break;
    float o[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float x = fmaf(acc[i][c], tv[c].x, tv[c].y + nzv[i]);
      x = x > 0.f ? x : x * tv[c].z;
      o[c] = x * tv[c].w;
    }
    uint2 w;
    w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]);
    *reinterpret_cast<uint2*>(out + ((static_cast<size_t>(b) * OH + Y) * OW + X) * cs + c0 + g * 4) = w;
  }
}

// ------------------------------------------------------------------------------------
// ToRGB tail: rgb_out = acc + bias + Upsample(skip)   (stylegan2.py:394-399; Upsample =
// upfirdn2d(up=2, pad=(2,1)) with kernel*4, stylegan2.py:52-63).  One thread per pixel.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rgb_finalize_kernel(float* __restrict__ rgb_out, float* __restrict__ acc,
                                                           const float* __restrict__ bias3, const float* __restrict__ skip,
                                                           const float* __restrict__ kernel, int H, int W, int64_t total) {
  __shared__ float s_k[16];
  if (threadIdx.x < 16) {
    const int a = threadIdx.x >> 2, b = threadIdx.x & 3;
    s_k[threadIdx.x] = kernel ? kernel[(3 - a) * 4 + (3 - b)] : 0.f;
  }
  __syncthreads();
  const int h2 = H / 2, w2 = W / 2;
  const float b0 = __ldg(bias3 + 0), b1 = __ldg(bias3 + 1), b2 = __ldg(bias3 + 2);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int x = static_cast<int>(idx % W);
    const int y = static_cast<int>((idx / W) % H);
    const int64_t b = idx / (static_cast<int64_t>(W) * H);
    float4* ap = reinterpret_cast<float4*>(acc + idx * 4);
    const float4 a = *ap;
    *ap = make_float4(0.f, 0.f, 0.f, 0.f);
    float v[3] = {a.x + b0, a.y + b1, a.z + b2};
    if (skip) {
      // zero-stuffed row u = y + ta - 2 is live when even: ta has the parity of y -> 2 of the 4 taps per axis
      const size_t plane = static_cast<size_t>(h2) * w2;
      const float* sp = skip + b * 3 * plane;
#pragma unroll
      for (int ia = 0; ia < 2; ++ia) {
        const int ta = (y & 1) + 2 * ia;
        const int sy = (y + ta - 2) >> 1;
        if (y + ta - 2 < 0 || sy >= h2) continue;
#pragma unroll
        for (int ib = 0; ib < 2; ++ib) {
          const int tb = (x & 1) + 2 * ib;
          const int sx = (x + tb - 2) >> 1;
          if (x + tb - 2 < 0 || sx >= w2) continue;
          const float kv = s_k[ta * 4 + tb];
          const size_t off = static_cast<size_t>(sy) * w2 + sx;
          v[0] = fmaf(__ldg(sp + off), kv, v[0]);
          v[1] = fmaf(__ldg(sp + plane + off), kv, v[1]);
          v[2] = fmaf(__ldg(sp + 2 * plane + off), kv, v[2]);
        }
      }
    }
    const size_t o = (b * 3 * H + y) * static_cast<size_t>(W) + x;
    rgb_out[o] = v[0];
    rgb_out[o + static_cast<size_t>(H) * W] = v[1];
    rgb_out[o + 2 * static_cast<size_t>(H) * W] = v[2];
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
  dim3 grid((max_cin + 63) / 64, n_layers);
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
  dim3 grid((max_cout + 63) / 64, n_layers);
  build_tables_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(layers_dev, B);
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

extern "C" int fm_nchw_to_nhwc_bf16(void* out, const float* x, const float* scale_bc, int B, int C, int H, int W,
                                    int out_cstride, void* stream) {
  FM_CHECK_ARG(out && x && B > 0 && C > 0 && H > 0 && W > 0, "fm_nchw_to_nhwc_bf16: bad args");
  FM_CHECK_ARG(out_cstride % 8 == 0 && out_cstride >= C, "fm_nchw_to_nhwc_bf16: cstride must be a multiple of 8 >= C");
  const int64_t total = static_cast<int64_t>(B) * (out_cstride / 8) * H * W;
  nchw_to_nhwc_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<__nv_bfloat16*>(out), x, scale_bc, C, H * W, out_cstride, total);
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
                                int separable, void* stream) {
  FM_CHECK_ARG(out && t && kernel4x4 && tab && B > 0 && OH > 0 && OW > 0 && C > 0, "fm_blur_act_nhwc: bad args");
  FM_CHECK_ARG(cstride % 8 == 0 && cstride >= C, "fm_blur_act_nhwc: cstride must be a multiple of 8 >= C");
  const int tiles_x = (OW + BL_T - 1) / BL_T, tiles_y = (OH + BL_T - 1) / BL_T, cblocks = (cstride + 63) / 64;
  const int64_t blocks = static_cast<int64_t>(B) * tiles_x * tiles_y * cblocks;
  FM_CHECK_ARG(blocks < 0x7FFFFFFF, "fm_blur_act_nhwc: too many blocks");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (separable)
    blur_act_nhwc_kernel<true><<<static_cast<unsigned>(blocks), 512, 0, st>>>(
        static_cast<__nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(t), kernel4x4, tab, noise, noise_bstride, noise_w,
        OH, OW, C, cstride, tiles_x, tiles_y, cblocks);
  else
    blur_act_nhwc_kernel<false><<<static_cast<unsigned>(blocks), 512, 0, st>>>(
        static_cast<__nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(t), kernel4x4, tab, noise, noise_bstride, noise_w,
        OH, OW, C, cstride, tiles_x, tiles_y, cblocks);
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

extern "C" int fm_rgb_finalize(float* rgb_out, float* acc, const float* bias3, const float* skip, const float* kernel4x4,
                               int B, int H, int W, void* stream) {
  FM_CHECK_ARG(rgb_out && acc && bias3 && B > 0 && H > 0 && W > 0, "fm_rgb_finalize: bad args");
  FM_CHECK_ARG(!skip || (kernel4x4 && H % 2 == 0 && W % 2 == 0), "fm_rgb_finalize: skip needs a kernel and even H, W");
  const int64_t total = static_cast<int64_t>(B) * H * W;
  rgb_finalize_kernel<<<grid_for(total, 256, 64), 256, 0, static_cast<cudaStream_t>(stream)>>>(rgb_out, acc, bias3, skip, kernel4x4,
                                                                                     H, W, total);
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
