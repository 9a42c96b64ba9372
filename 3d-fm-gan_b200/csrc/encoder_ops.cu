// Bandwidth-bound NHWC bf16 helpers for the encoder path (ResNet-18 and pSp IR-SE encoders):
// stem input packing, max/avg pooling, squeeze-and-excitation, bilinear FPN upsampling.
// Each replaces a separate ATen kernel of the reference graph (resnet_encoder.py:258-280,
// psp_encoder_model/encoders/helpers.py:76-139, psp_encoders.py:81-98).
#include "common.cuh"

namespace fm {

__device__ __forceinline__ void unpack8(const uint4& w, float (&v)[8]) {
  float2 f;
  f = unpack_bf16x2(w.x); v[0] = f.x; v[1] = f.y;
  f = unpack_bf16x2(w.y); v[2] = f.x; v[3] = f.y;
  f = unpack_bf16x2(w.z); v[4] = f.x; v[5] = f.y;
  f = unpack_bf16x2(w.w); v[6] = f.x; v[7] = f.y;
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 w;
  w.x = pack_bf16x2(v[0], v[1]); w.y = pack_bf16x2(v[2], v[3]);
  w.z = pack_bf16x2(v[4], v[5]); w.w = pack_bf16x2(v[6], v[7]);
  return w;
}

static inline unsigned grid_for2(int64_t total, int waves = 16) {
  int64_t want = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * waves;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return static_cast<unsigned>(want);
}

// fp32 NCHW image (C <= 8) -> zero-padded bf16 [B, Hp, Wp, 8]
__global__ void __launch_bounds__(256) image_pack_kernel(__nv_bfloat16* __restrict__ out, const float* __restrict__ x, int C,
                                                         int H, int W, int pad_t, int pad_l, int Hp, int Wp, int64_t total) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    int xp, yp;
    int64_t b;
    if (total < 0x7fffffffLL) {            // 32-bit index arithmetic: a 64-bit division costs ~70 instructions, this kernel has three
      const uint32_t i32 = static_cast<uint32_t>(idx), r1 = i32 / static_cast<uint32_t>(Wp), r2 = r1 / static_cast<uint32_t>(Hp);
      xp = static_cast<int>(i32 - r1 * Wp);
      yp = static_cast<int>(r1 - r2 * Hp);
      b = r2;
    } else {
      xp = static_cast<int>(idx % Wp);
      yp = static_cast<int>((idx / Wp) % Hp);
      b = idx / (static_cast<int64_t>(Wp) * Hp);
    }
    const int y = yp - pad_t, xx = xp - pad_l;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (y >= 0 && y < H && xx >= 0 && xx < W) {
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c < C) v[c] = x[((b * C + c) * H + y) * static_cast<int64_t>(W) + xx];
    }
    *reinterpret_cast<uint4*>(out + idx * 8) = pack8(v);
  }
}

// 3x3 stride-2 pad-1 max pooling, NHWC bf16
__global__ void __launch_bounds__(256) maxpool_kernel(__nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ x,
                                                      int H, int W, int OH, int OW, int cs, int64_t total) {
  const int groups = cs / 8;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int g = static_cast<int>(idx % groups);
    int64_t r = idx / groups;
    const int ox = static_cast<int>(r % OW); r /= OW;
    const int oy = static_cast<int>(r % OH);
    const int64_t b = r / OH;
    float m[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) m[c] = -INFINITY;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 - 1 + ky;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * 2 - 1 + kx;
        if (ix < 0 || ix >= W) continue;
        float v[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + ((b * H + iy) * W + ix) * cs + g * 8)), v);
#pragma unroll
        for (int c = 0; c < 8; ++c) m[c] = fmaxf(m[c], v[c]);
      }
    }
    *reinterpret_cast<uint4*>(out + ((b * OH + oy) * OW + ox) * cs + g * 8) = pack8(m);
  }
}

// The same on 16 channels (32 bytes) per thread, packed bf16 maxima (max of bf16 values is exact in bf16: no fp32 round
// trip) and 32-bit index arithmetic (the 64-bit divisions of the form above cost more than its nine loads).
__global__ void __launch_bounds__(256) maxpool16_kernel(__nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ x,
                                                        int H, int W, int OH, int OW, int cs, uint32_t total) {
  const uint32_t groups = static_cast<uint32_t>(cs) / 16;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const uint32_t r1 = idx / groups, g = idx - r1 * groups;
    const uint32_t r2 = r1 / OW, ox = r1 - r2 * OW;
    const uint32_t b = r2 / OH, oy = r2 - b * OH;
    __nv_bfloat162 m[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) m[c] = __float2bfloat162_rn(-INFINITY);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = static_cast<int>(oy) * 2 - 1 + ky;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = static_cast<int>(ox) * 2 - 1 + kx;
        if (ix < 0 || ix >= W) continue;
        uint4 a, c4;
        ld_global_nc_256(x + ((static_cast<size_t>(b) * H + iy) * W + ix) * cs + g * 16, a, c4);
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, c4.x, c4.y, c4.z, c4.w};
#pragma unroll
        for (int c = 0; c < 8; ++c) m[c] = __hmax2(m[c], *reinterpret_cast<const __nv_bfloat162*>(&w[c]));
      }
    }
    uint4 o0, o1;
    o0.x = *reinterpret_cast<uint32_t*>(&m[0]); o0.y = *reinterpret_cast<uint32_t*>(&m[1]);
    o0.z = *reinterpret_cast<uint32_t*>(&m[2]); o0.w = *reinterpret_cast<uint32_t*>(&m[3]);
    o1.x = *reinterpret_cast<uint32_t*>(&m[4]); o1.y = *reinterpret_cast<uint32_t*>(&m[5]);
    o1.z = *reinterpret_cast<uint32_t*>(&m[6]); o1.w = *reinterpret_cast<uint32_t*>(&m[7]);
    st_global_256(out + ((static_cast<size_t>(b) * OH + oy) * OW + ox) * cs + g * 16, o0, o1);
  }
}

// non-overlapping ph x pw average pooling, NHWC bf16 -> NCHW fp32
__global__ void __launch_bounds__(256) avgpool_kernel(float* __restrict__ out, const __nv_bfloat16* __restrict__ x, int H, int W,
                                                      int C, int cs, int ph, int pw, int64_t total) {
  const int OH = H / ph, OW = W / pw;
  const float inv = 1.f / (ph * pw);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    // idx enumerates (b, oy, ox, c) so that reads are channel-contiguous
    const int c = static_cast<int>(idx % C);
    int64_t r = idx / C;
    const int ox = static_cast<int>(r % OW); r /= OW;
    const int oy = static_cast<int>(r % OH);
    const int64_t b = r / OH;
    float acc = 0.f;
    for (int dy = 0; dy < ph; ++dy)
      for (int dx = 0; dx < pw; ++dx)
        acc += __bfloat162float(x[((b * H + oy * ph + dy) * W + ox * pw + dx) * cs + c]);
    out[((b * C + c) * OH + oy) * OW + ox] = acc * inv;
  }
}

// sum over pixels per (b, c): grid (chunks, B); fp32 atomics into sum[B][C] (zeroed by the caller)
__global__ void __launch_bounds__(256) channel_sum_kernel(float* __restrict__ sum, const __nv_bfloat16* __restrict__ x, int HW,
                                                          int C, int cs, int chunk) {
  __shared__ float s_part[256 * 8];
  const int groups = cs / 8;
  const int lanes = 256 / groups;               // pixel lanes per block (groups <= 64 -> >= 4)
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * chunk;
  const int p1 = min(p0 + chunk, HW);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (pl < lanes) {
    for (int p = p0 + pl; p < p1; p += lanes) {
      float v[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + (static_cast<int64_t>(b) * HW + p) * cs + g * 8)), v);
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] += v[c];
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) s_part[threadIdx.x * 8 + c] = (pl < lanes) ? acc[c] : 0.f;
  __syncthreads();
  if (threadIdx.x < groups * 8) {
    const int gg = threadIdx.x / 8, c = threadIdx.x % 8;
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += s_part[(l * groups + gg) * 8 + c];
    const int ch = gg * 8 + c;
    if (ch < C) atomicAdd(sum + static_cast<int64_t>(b) * C + ch, t);
  }
}

// gate[b,c] = sigmoid(w2[c,:] . relu(w1 . mean[b,:]));  one block per sample.  Clears sum afterwards.
// Both matrix-vector products issue all their (independent) weight loads before the first FMA: the first version walked
// them in rolled loops of dependent-looking scalar loads and spent 19 us at C = 512 on L2 latency alone.
__global__ void __launch_bounds__(256) se_gate_kernel(float* __restrict__ gate, float* __restrict__ sum, float inv_hw,
                                                      const float* __restrict__ w1, const float* __restrict__ w2, int C, int Cr) {
  __shared__ __align__(16) float s_mean[1024];
  __shared__ __align__(16) float s_hid[64];
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += 256) {
    s_mean[c] = sum[static_cast<int64_t>(b) * C + c] * inv_hw;
    sum[static_cast<int64_t>(b) * C + c] = 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool vec = (C & 127) == 0 && (Cr & 3) == 0;         // float4 rows (C = 64 takes the scalar path)
  for (int j = warp; j < Cr; j += 8) {
    float acc = 0.f;
    if (vec) {
      const float4* wr = reinterpret_cast<const float4*>(w1 + static_cast<int64_t>(j) * C);
      const float4* mr = reinterpret_cast<const float4*>(s_mean);
      float4 wv[8];
      const int n4 = C >> 7;                                  // float4 per lane (C <= 1024 -> <= 8)
#pragma unroll
      for (int i = 0; i < 8; ++i) wv[i] = i < n4 ? __ldg(wr + i * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i < n4) {
          const float4 m = mr[i * 32 + lane];
          acc = fmaf(wv[i].x, m.x, fmaf(wv[i].y, m.y, fmaf(wv[i].z, m.z, fmaf(wv[i].w, m.w, acc))));
        }
    } else {
      for (int c = lane; c < C; c += 32) acc = fmaf(__ldg(w1 + static_cast<int64_t>(j) * C + c), s_mean[c], acc);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) s_hid[j] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float acc = 0.f;
    if (vec) {
      const float4* wr = reinterpret_cast<const float4*>(w2 + static_cast<int64_t>(c) * Cr);
      const float4* hr = reinterpret_cast<const float4*>(s_hid);
      float4 wv[16];
      const int n4 = Cr >> 2;                                 // Cr <= 64 -> <= 16
#pragma unroll
      for (int i = 0; i < 16; ++i) wv[i] = i < n4 ? __ldg(wr + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i < n4) {
          const float4 h = hr[i];
          acc = fmaf(wv[i].x, h.x, fmaf(wv[i].y, h.y, fmaf(wv[i].z, h.z, fmaf(wv[i].w, h.w, acc))));
        }
    } else {
      for (int j = 0; j < Cr; ++j) acc = fmaf(__ldg(w2 + static_cast<int64_t>(c) * Cr + j), s_hid[j], acc);
    }
    gate[static_cast<int64_t>(b) * C + c] = 1.f / (1.f + __expf(-acc));
  }
}

// out = r * gate[b,c] + shortcut[b, y*ss, x*ss, c]
__global__ void __launch_bounds__(256) se_combine_kernel(__nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ r,
                                                         const float* __restrict__ gate, const __nv_bfloat16* __restrict__ sc,
                                                         int H, int W, int C, int cs, int sc_H, int sc_W, int sc_cs, int ss,
                                                         int64_t total) {
  const int groups = cs / 8;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int g = static_cast<int>(idx % groups);
    int64_t t = idx / groups;
    const int x = static_cast<int>(t % W); t /= W;
    const int y = static_cast<int>(t % H);
    const int64_t b = t / H;
    float rv[8], sv[8], o[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(r + ((b * H + y) * W + x) * cs + g * 8)), rv);
    unpack8(__ldg(reinterpret_cast<const uint4*>(sc + ((b * sc_H + y * ss) * sc_W + x * ss) * sc_cs + g * 8)), sv);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int ch = g * 8 + c;
      o[c] = ch < C ? fmaf(rv[c], __ldg(gate + b * C + ch), sv[c]) : 0.f;
    }
    *reinterpret_cast<uint4*>(out + ((b * H + y) * W + x) * cs + g * 8) = pack8(o);
  }
}

// The same with 16 channels (32 bytes) per thread: one 256-bit load of r and of the shortcut, one 256-bit store -- half
// the memory instructions of the 8-channel form for the same bytes (cs, sc_cs multiples of 16; 32-byte aligned tensors).
__global__ void __launch_bounds__(256) se_combine16_kernel(__nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ r,
                                                           const float* __restrict__ gate, const __nv_bfloat16* __restrict__ sc,
                                                           int H, int W, int C, int cs, int sc_H, int sc_W, int sc_cs, int ss,
                                                           int64_t total) {
  const int groups = cs / 16;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int g = static_cast<int>(idx % groups);
    int64_t t = idx / groups;
    const int x = static_cast<int>(t % W); t /= W;
    const int y = static_cast<int>(t % H);
    const int64_t b = t / H;
    uint4 ra, rb, sa, sb;
    ld_global_nc_256(r + ((b * H + y) * W + x) * cs + g * 16, ra, rb);
    ld_global_nc_256(sc + ((b * sc_H + y * ss) * sc_W + x * ss) * sc_cs + g * 16, sa, sb);
    const float4* gp = reinterpret_cast<const float4*>(gate + b * C + g * 16);
    float gv[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 v = (g * 16 + 4 * q + 3 < C) ? __ldg(gp + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      gv[4 * q] = v.x; gv[4 * q + 1] = v.y; gv[4 * q + 2] = v.z; gv[4 * q + 3] = v.w;
    }
    float r0[8], r1[8], s0[8], s1[8], o0[8], o1[8];
    unpack8(ra, r0); unpack8(rb, r1);
    unpack8(sa, s0); unpack8(sb, s1);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      o0[c] = (g * 16 + c < C) ? fmaf(r0[c], gv[c], s0[c]) : 0.f;
      o1[c] = (g * 16 + 8 + c < C) ? fmaf(r1[c], gv[8 + c], s1[c]) : 0.f;
    }
    st_global_256(out + ((b * H + y) * W + x) * cs + g * 16, pack8(o0), pack8(o1));
  }
}

// bilinear resize with align_corners=True (F.interpolate, psp_encoders.py:98), NHWC bf16
__global__ void __launch_bounds__(256) bilinear_kernel(__nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ x,
                                                       int IH, int IW, int OH, int OW, int cs, float ry, float rx,
                                                       int64_t total) {
  const int groups = cs / 8;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int g = static_cast<int>(idx % groups);
    int64_t t = idx / groups;
    const int ox = static_cast<int>(t % OW); t /= OW;
    const int oy = static_cast<int>(t % OH);
    const int64_t b = t / OH;
    const float fy = ry * oy, fx = rx * ox;
    const int y0 = min(static_cast<int>(fy), IH - 1), x0 = min(static_cast<int>(fx), IW - 1);
    const int y1 = min(y0 + 1, IH - 1), x1 = min(x0 + 1, IW - 1);
    const float ly = fy - y0, lx = fx - x0;
    float a[8], bq[8], c[8], d[8], o[8];
    const __nv_bfloat16* base = x + b * IH * IW * static_cast<int64_t>(cs) + g * 8;
    unpack8(__ldg(reinterpret_cast<const uint4*>(base + (static_cast<int64_t>(y0) * IW + x0) * cs)), a);
    unpack8(__ldg(reinterpret_cast<const uint4*>(base + (static_cast<int64_t>(y0) * IW + x1) * cs)), bq);
    unpack8(__ldg(reinterpret_cast<const uint4*>(base + (static_cast<int64_t>(y1) * IW + x0) * cs)), c);
    unpack8(__ldg(reinterpret_cast<const uint4*>(base + (static_cast<int64_t>(y1) * IW + x1) * cs)), d);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      o[k] = (1.f - ly) * ((1.f - lx) * a[k] + lx * bq[k]) + ly * ((1.f - lx) * c[k] + lx * d[k]);
    *reinterpret_cast<uint4*>(out + ((b * OH + oy) * OW + ox) * cs + g * 8) = pack8(o);
  }
}

}  // namespace fm

using namespace fm;
#define ST static_cast<cudaStream_t>(stream)

extern "C" int fm_image_to_nhwc8_padded(void* out, const float* x, int B, int C, int H, int W, int pad_t, int pad_l, int Hp,
                                        int Wp, void* stream) {
  FM_CHECK_ARG(out && x && B > 0 && C > 0 && C <= 8 && H > 0 && W > 0, "fm_image_to_nhwc8_padded: bad args");
  FM_CHECK_ARG(pad_t >= 0 && pad_l >= 0 && Hp >= H + pad_t && Wp >= W + pad_l, "fm_image_to_nhwc8_padded: bad padding");
  const int64_t total = static_cast<int64_t>(B) * Hp * Wp;
  image_pack_kernel<<<grid_for2(total), 256, 0, ST>>>(static_cast<__nv_bfloat16*>(out), x, C, H, W, pad_t, pad_l, Hp, Wp, total);
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

extern "C" int fm_maxpool3x3s2_nhwc(void* out, const void* x, int B, int H, int W, int cs, void* stream) {
  FM_CHECK_ARG(out && x && B > 0 && H > 0 && W > 0 && cs > 0 && cs % 8 == 0, "fm_maxpool3x3s2_nhwc: bad args");
  const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  if (cs % 16 == 0 && ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(x)) & 31) == 0 &&
      static_cast<int64_t>(B) * OH * OW * (cs / 16) < 0x7fffffffLL) {
    const int64_t total16 = static_cast<int64_t>(B) * OH * OW * (cs / 16);
    maxpool16_kernel<<<grid_for2(total16), 256, 0, ST>>>(static_cast<__nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(x), H, W,
                                                         OH, OW, cs, static_cast<uint32_t>(total16));
    count_launch();
    FM_LAUNCH_OK();
    return FM_OK;
  }
  const int64_t total = static_cast<int64_t>(B) * OH * OW * (cs / 8);
  maxpool_kernel<<<grid_for2(total), 256, 0, ST>>>(static_cast<__nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(x), H, W,
                                                   OH, OW, cs, total);
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

extern "C" int fm_avgpool_nhwc_to_nchw(float* out, const void* x, int B, int H, int W, int C, int cs, int ph, int pw,
                                       void* stream) {
  FM_CHECK_ARG(out && x && B > 0 && H > 0 && W > 0 && C > 0 && cs >= C && ph > 0 && pw > 0 && H % ph == 0 && W % pw == 0,
               "fm_avgpool_nhwc_to_nchw: bad args");
  const int64_t total = static_cast<int64_t>(B) * (H / ph) * (W / pw) * C;
  avgpool_kernel<<<grid_for2(total), 256, 0, ST>>>(out, static_cast<const __nv_bfloat16*>(x), H, W, C, cs, ph, pw, total);
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

// ------------------------------------------------------------------------------------
// tensor2im (Evaluation/visual_eval.py:24-38 of the reference): fp32 NCHW [B,3,H,W] in [-1,1] ->
// uint8 NHWC [B,H,W,3] = trunc((clip(x,-1,1) + cent) * factor), for the whole batch on the device.
// The reference does this per image on the host (.cpu().float().numpy(), np.clip, transpose, astype):
// 4x the D2H bytes plus three numpy passes.  A thread owns 4 adjacent pixels: 3 LDG.128, 12 bytes out.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tensor2im_kernel(uint8_t* __restrict__ out, const float* __restrict__ img, int HW,
                                                        float cent, float factor, int total_quads) {
  const int qpi = HW >> 2;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total_quads; idx += gridDim.x * blockDim.x) {
    const int b = idx / qpi, q = idx - b * qpi;
    const float* ip = img + static_cast<size_t>(b) * 3 * HW + 4 * q;
    float v[3][4];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float4 f = __ldg(reinterpret_cast<const float4*>(ip + static_cast<size_t>(c) * HW));
      v[c][0] = f.x; v[c][1] = f.y; v[c][2] = f.z; v[c][3] = f.w;
    }
    uint32_t w[3] = {0u, 0u, 0u};
#pragma unroll
    for (int i = 0; i < 12; ++i) {            // byte i of the 12-byte group = pixel i/3, channel i%3
      const float x = fminf(fmaxf(v[i % 3][i / 3], -1.f), 1.f);
      // separate rounded add and multiply (never contracted to an FMA): bit-identical to numpy's float32 passes
      const uint32_t u = static_cast<uint32_t>(static_cast<int>(__fmul_rn(__fadd_rn(x, cent), factor))) & 0xffu;   // astype: truncation
      w[i >> 2] |= u << (8 * (i & 3));
    }
    uint32_t* op = reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(b) * HW + 4 * q) * 3);
    op[0] = w[0]; op[1] = w[1]; op[2] = w[2];
  }
}

extern "C" int fm_tensor2im_u8(void* out_u8, const float* img, int B, int H, int W, float cent, float factor, void* stream) {
  FM_CHECK_ARG(out_u8 && img && B > 0 && H > 0 && W > 0, "fm_tensor2im_u8: bad args");
  FM_CHECK_ARG((static_cast<int64_t>(H) * W) % 4 == 0, "fm_tensor2im_u8: H*W must be a multiple of 4");
  FM_CHECK_ARG((reinterpret_cast<uintptr_t>(img) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_u8) & 3) == 0,
               "fm_tensor2im_u8: image must be 16-byte aligned, output 4-byte aligned");
  const int64_t total = static_cast<int64_t>(B) * H * W / 4;
  FM_CHECK_ARG(total < 0x7FFFFFFF, "fm_tensor2im_u8: too many pixels");
  tensor2im_kernel<<<grid_for2(total), 256, 0, ST>>>(static_cast<uint8_t*>(out_u8), img, H * W, cent, factor,
                                                     static_cast<int>(total));
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

// ------------------------------------------------------------------------------------
// Input stage (the inverse of tensor2im): uint8 NHWC [B,H,W,3] -> fp32 NCHW [B,3,H,W]
//   y = (x / 255 - mean) / std      == torchvision ToTensor() + Normalize(mean, std) of the reference's transform
// (train_3_encoder.py:231-237), bit for bit: a rounded fp32 division by 255, a rounded subtraction, a rounded division
// (the library is built with --use_fast_math, so the IEEE intrinsics are spelled out).  The decoded uint8 batch is what
// crosses PCIe (1/4 of the fp32 bytes); a thread owns 4 adjacent pixels: 12 bytes in, 3 STG.128 out.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) im2tensor_kernel(float* __restrict__ out, const uint8_t* __restrict__ img, int HW,
                                                        float mean, float stdv, int total_quads) {
  const int qpi = HW >> 2;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total_quads; idx += gridDim.x * blockDim.x) {
    const int b = idx / qpi, q = idx - b * qpi;
    const uint32_t* ip = reinterpret_cast<const uint32_t*>(img + (static_cast<size_t>(b) * HW + 4 * q) * 3);
    const uint32_t w[3] = {__ldg(ip), __ldg(ip + 1), __ldg(ip + 2)};
    float v[3][4];
#pragma unroll
    for (int i = 0; i < 12; ++i) {            // byte i of the 12-byte group = pixel i/3, channel i%3
      const float x = static_cast<float>((w[i >> 2] >> (8 * (i & 3))) & 0xffu);
      v[i % 3][i / 3] = __fdiv_rn(__fsub_rn(__fdiv_rn(x, 255.f), mean), stdv);
    }
    float* op = out + static_cast<size_t>(b) * 3 * HW + 4 * q;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      *reinterpret_cast<float4*>(op + static_cast<size_t>(c) * HW) = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
  }
}

extern "C" int fm_im2tensor_f32(float* out, const void* img_u8, int B, int H, int W, float mean, float stdv, void* stream) {
  FM_CHECK_ARG(out && img_u8 && B > 0 && H > 0 && W > 0 && stdv != 0.f, "fm_im2tensor_f32: bad args");
  FM_CHECK_ARG((static_cast<int64_t>(H) * W) % 4 == 0, "fm_im2tensor_f32: H*W must be a multiple of 4");
  FM_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(img_u8) & 3) == 0,
               "fm_im2tensor_f32: output must be 16-byte aligned, image 4-byte aligned");
  const int64_t total = static_cast<int64_t>(B) * H * W / 4;
  FM_CHECK_ARG(total < 0x7FFFFFFF, "fm_im2tensor_f32: too many pixels");
  im2tensor_kernel<<<grid_for2(total), 256, 0, ST>>>(out, static_cast<const uint8_t*>(img_u8), H * W, mean, stdv,
                                                     static_cast<int>(total));
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

extern "C" int fm_channel_sum_nhwc(float* sum_bc, const void* x, int B, int HW, int C, int cs, void* stream) {
  FM_CHECK_ARG(sum_bc && x && B > 0 && HW > 0 && C > 0 && cs >= C && cs % 8 == 0 && cs <= 512 && 256 % (cs / 8) == 0,
               "fm_channel_sum_nhwc: bad args (cs must be 8*2^k <= 512)");
  // enough CTAs to fill the chip even for small maps (16x16: one CTA per sample would idle 116 SMs)
  int per_sample = (4 * sm_count() + B - 1) / B;
  if (per_sample < 1) per_sample = 1;
  int chunk = (HW + per_sample - 1) / per_sample;
  if (chunk < 32) chunk = 32;
  dim3 grid((HW + chunk - 1) / chunk, B);
  channel_sum_kernel<<<grid, 256, 0, ST>>>(sum_bc, static_cast<const __nv_bfloat16*>(x), HW, C, cs, chunk);
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

extern "C" int fm_se_gate(float* gate_bc, float* sum_bc, float inv_hw, const float* w1, const float* w2, int B, int C, int Cr,
                          void* stream) {
  FM_CHECK_ARG(gate_bc && sum_bc && w1 && w2 && B > 0 && C > 0 && C <= 1024 && Cr > 0 && Cr <= 64, "fm_se_gate: bad args");
  se_gate_kernel<<<B, 256, 0, ST>>>(gate_bc, sum_bc, inv_hw, w1, w2, C, Cr);
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

extern "C" int fm_se_combine_nhwc(void* out, const void* r, const float* gate_bc, const void* sc, int B, int H, int W, int C,
                                  int cs, int sc_H, int sc_W, int sc_cs, int sc_stride, void* stream) {
  FM_CHECK_ARG(out && r && gate_bc && sc && B > 0 && H > 0 && W > 0 && C > 0 && cs >= C && cs % 8 == 0 && sc_cs >= cs - 7 &&
                   sc_cs % 8 == 0 && sc_stride >= 1, "fm_se_combine_nhwc: bad args");
  FM_CHECK_ARG((H - 1) * sc_stride < sc_H && (W - 1) * sc_stride < sc_W, "fm_se_combine_nhwc: shortcut too small");
  if (cs % 16 == 0 && sc_cs % 16 == 0 && C % 4 == 0 && ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(r) |
                                                          reinterpret_cast<uintptr_t>(sc)) & 31) == 0 &&
      (reinterpret_cast<uintptr_t>(gate_bc) & 15) == 0) {
    const int64_t total16 = static_cast<int64_t>(B) * H * W * (cs / 16);
    se_combine16_kernel<<<grid_for2(total16), 256, 0, ST>>>(static_cast<__nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(r),
                                                            gate_bc, static_cast<const __nv_bfloat16*>(sc), H, W, C, cs, sc_H, sc_W,
                                                            sc_cs, sc_stride, total16);
    count_launch();
    FM_LAUNCH_OK();
    return FM_OK;
  }
  const int64_t total = static_cast<int64_t>(B) * H * W * (cs / 8);
  se_combine_kernel<<<grid_for2(total), 256, 0, ST>>>(static_cast<__nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(r),
                                                      gate_bc, static_cast<const __nv_bfloat16*>(sc), H, W, C, cs, sc_H, sc_W,
                                                      sc_cs, sc_stride, total);
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

extern "C" int fm_bilinear_up_nhwc(void* out, const void* x, int B, int IH, int IW, int OH, int OW, int cs, void* stream) {
  FM_CHECK_ARG(out && x && B > 0 && IH > 0 && IW > 0 && OH > 0 && OW > 0 && cs > 0 && cs % 8 == 0, "fm_bilinear_up_nhwc: bad args");
  const float ry = OH > 1 ? static_cast<float>(IH - 1) / (OH - 1) : 0.f;
  const float rx = OW > 1 ? static_cast<float>(IW - 1) / (OW - 1) : 0.f;
  const int64_t total = static_cast<int64_t>(B) * OH * OW * (cs / 8);
  bilinear_kernel<<<grid_for2(total), 256, 0, ST>>>(static_cast<__nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(x), IH,
                                                    IW, OH, OW, cs, ry, rx, total);
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}
