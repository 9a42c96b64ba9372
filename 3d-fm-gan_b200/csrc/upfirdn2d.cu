// upfirdn2d for sm_100a: upsample (zero-stuff) -> pad/crop -> FIR -> decimate on [planes, H, W].
//
// Replaces op/upfirdn2d_kernel.cu (reference): definition
//   out[p, oy, ox] = sum_{ky,kx} xpad_up[p, oy*down_y + ky, ox*down_x + kx] * k[kh-1-ky, kw-1-kx]
// Design: one CTA stages the input footprint of a TILE_H x TILE_W output tile in shared
// memory with coalesced row loads (zero-filled outside the image), taps live in registers
// (compile-time KH x KW, already flipped), every thread produces a strip of 4 horizontally
// adjacent outputs so each staged input is reused from registers, accumulation is fp32
// for every dtype, and the strip is written with one 16-byte store when aligned.
// Algorithmic bytes: (planes*in_h*in_w + planes*out_h*out_w) * sizeof(T); roofline = HBM.
#include "common.cuh"

namespace fm {

__host__ __device__ __forceinline__ int floor_div(int a, int b) {
  int q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

struct UpfirdnParams {
  int in_h, in_w, out_h, out_w;
  int pad_x0, pad_y0;
  int up_x, up_y, down_x, down_y, kh, kw;
  int tiles_x, tiles_y;
};

constexpr int UF_TILE_W = 128;  // 32 threads x 4 outputs
constexpr int UF_ROWS_PER_PASS = 8;

// Tiled kernel: UP/DOWN/KH/KW are compile-time (square factors), kernel taps in smem.
template <typename T, int UP, int DOWN, int KH, int KW, int TILE_H>
__global__ void __launch_bounds__(256) upfirdn2d_tile_kernel(T* __restrict__ out, const T* __restrict__ x,
                                                             const float* __restrict__ kernel, UpfirdnParams p,
                                                             int64_t planes) {
  // input footprint of the tile (in source samples)
  constexpr int IN_H = ((TILE_H - 1) * DOWN + KH - 1) / UP + 2;
  constexpr int IN_W = ((UF_TILE_W - 1) * DOWN + KW - 1) / UP + 2;
  // strips are read as aligned float4s: NV vectors starting at column 4*lane*DOWN
  constexpr int NV = (3 * DOWN + KW + 3) / 4;
  constexpr int IN_WV = (UF_TILE_W - 4) * DOWN + 4 * NV;
  constexpr int IN_WP = ((IN_W > IN_WV ? IN_W : IN_WV) + 3) & ~3;
  __shared__ __align__(16) float s_in[IN_H][IN_WP];
  __shared__ float s_k[KH][KW];

  const int tile_x = blockIdx.x % p.tiles_x;
  const int tile_y = blockIdx.x / p.tiles_x;
  const int oy0 = tile_y * TILE_H, ox0 = tile_x * UF_TILE_W;
  const int in_y0 = floor_div(oy0 * DOWN - p.pad_y0, UP);
  const int in_x0 = floor_div(ox0 * DOWN - p.pad_x0, UP);
  const int tid = threadIdx.x;

  if (tid < KH * KW) {
    const int ky = tid / KW, kx = tid % KW;
    // flipped: tap (ky,kx) of the correlation = k[kh-1-ky][kw-1-kx]  (op/upfirdn2d_kernel.cu:137)
    s_k[ky][kx] = (ky < p.kh && kx < p.kw) ? kernel[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)] : 0.f;
  }

  for (int64_t plane = blockIdx.y; plane < planes; plane += gridDim.y) {
    const T* xp = x + plane * static_cast<int64_t>(p.in_h) * p.in_w;
    __syncthreads();  // previous plane's readers are done; s_k is visible
    for (int i = tid; i < IN_H * IN_W; i += 256) {
      const int ry = i / IN_W, rx = i - ry * IN_W;
      const int iy = in_y0 + ry, ix = in_x0 + rx;
      float v = 0.f;
      if (iy >= 0 && iy < p.in_h && ix >= 0 && ix < p.in_w) v = to_f32<T>(xp[static_cast<int64_t>(iy) * p.in_w + ix]);
      s_in[ry][rx] = v;
    }
    __syncthreads();

    T* op = out + plane * static_cast<int64_t>(p.out_h) * p.out_w;
    const int lx = (tid & 31) * 4;
#pragma unroll
    for (int pass = 0; pass < TILE_H / UF_ROWS_PER_PASS; ++pass) {
      const int ly = pass * UF_ROWS_PER_PASS + (tid >> 5);
      const int oy = oy0 + ly, ox = ox0 + lx;
      if (oy >= p.out_h || ox >= p.out_w) continue;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      if (UP == 1) {
        // rows: Y = oy*DOWN + ky - pad ; source row = Y (UP==1)
        const int ry0 = oy * DOWN - p.pad_y0 - in_y0;
        const int rx0 = ox * DOWN - p.pad_x0 - in_x0;
#pragma unroll
        for (int ky = 0; ky < KH; ++ky) {
          float row[4 * NV];
#pragma unroll
          for (int j = 0; j < NV; ++j)
            *reinterpret_cast<float4*>(&row[4 * j]) = *reinterpret_cast<const float4*>(&s_in[ry0 + ky][rx0 + 4 * j]);
#pragma unroll
          for (int kx = 0; kx < KW; ++kx) {
            const float kv = s_k[ky][kx];
#pragma unroll
            for (int o = 0; o < 4; ++o) acc[o] = fmaf(row[o * DOWN + kx], kv, acc[o]);
          }
        }
      } else {
        // zero-stuffed: only taps with (Y % UP == 0) hit a sample
#pragma unroll
        for (int ky = 0; ky < KH; ++ky) {
          const int Y = oy * DOWN + ky - p.pad_y0;
          if (((Y % UP) + UP) % UP != 0) continue;
          const int ry = floor_div(Y, UP) - in_y0;
#pragma unroll
          for (int o = 0; o < 4; ++o) {
#pragma unroll
            for (int kx = 0; kx < KW; ++kx) {
              const int X = (ox + o) * DOWN + kx - p.pad_x0;
              if (((X % UP) + UP) % UP != 0) continue;
              acc[o] = fmaf(s_in[ry][floor_div(X, UP) - in_x0], s_k[ky][kx], acc[o]);
            }
          }
        }
      }
      T* orow = op + static_cast<int64_t>(oy) * p.out_w + ox;
      if (sizeof(T) == 4 && ox + 3 < p.out_w && (reinterpret_cast<uintptr_t>(orow) & 15) == 0) {
        *reinterpret_cast<float4*>(orow) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      } else if (sizeof(T) == 2 && ox + 3 < p.out_w && (reinterpret_cast<uintptr_t>(orow) & 7) == 0) {
        T q[4] = {from_f32<T>(acc[0]), from_f32<T>(acc[1]), from_f32<T>(acc[2]), from_f32<T>(acc[3])};
        *reinterpret_cast<uint2*>(orow) = *reinterpret_cast<uint2*>(q);
      } else {
#pragma unroll
        for (int o = 0; o < 4; ++o)
          if (ox + o < p.out_w) orow[o] = from_f32<T>(acc[o]);
      }
    }
  }
}

// Generic kernel: any up/down (per axis), any kernel size, one thread per output.
template <typename T>
__global__ void __launch_bounds__(256) upfirdn2d_generic_kernel(T* __restrict__ out, const T* __restrict__ x,
                                                                const float* __restrict__ kernel, UpfirdnParams p,
                                                                int64_t total) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int ox = static_cast<int>(i % p.out_w);
    const int64_t t = i / p.out_w;
    const int oy = static_cast<int>(t % p.out_h);
    const int64_t plane = t / p.out_h;
    const T* xp = x + plane * static_cast<int64_t>(p.in_h) * p.in_w;
    float acc = 0.f;
    for (int ky = 0; ky < p.kh; ++ky) {
      const int Y = oy * p.down_y + ky - p.pad_y0;
      if (((Y % p.up_y) + p.up_y) % p.up_y != 0) continue;
      const int iy = floor_div(Y, p.up_y);
      if (iy < 0 || iy >= p.in_h) continue;
      for (int kx = 0; kx < p.kw; ++kx) {
        const int X = ox * p.down_x + kx - p.pad_x0;
        if (((X % p.up_x) + p.up_x) % p.up_x != 0) continue;
        const int ix = floor_div(X, p.up_x);
        if (ix < 0 || ix >= p.in_w) continue;
        acc = fmaf(to_f32<T>(xp[static_cast<int64_t>(iy) * p.in_w + ix]),
                   __ldg(kernel + (p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)), acc);
      }
    }
    out[i] = from_f32<T>(acc);
  }
}

template <typename T, int UP, int DOWN, int K, int TILE_H>
static int launch_tile(void* out, const void* x, const float* kernel, UpfirdnParams p, int64_t planes, cudaStream_t st) {
  p.tiles_x = (p.out_w + UF_TILE_W - 1) / UF_TILE_W;
  p.tiles_y = (p.out_h + TILE_H - 1) / TILE_H;
  const int64_t tiles = static_cast<int64_t>(p.tiles_x) * p.tiles_y;
  FM_CHECK_ARG(tiles < 0x7FFFFFFF, "fm_upfirdn2d: image too large");
  const unsigned gy = static_cast<unsigned>(planes < 65535 ? planes : 65535);
  dim3 grid(static_cast<unsigned>(tiles), gy);
  upfirdn2d_tile_kernel<T, UP, DOWN, K, K, TILE_H><<<grid, 256, 0, st>>>(static_cast<T*>(out), static_cast<const T*>(x),
                                                                        kernel, p, planes);
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

template <typename T>
static int upfirdn2d_dispatch(void* out, const void* x, const float* kernel, int64_t planes, UpfirdnParams p,
                              cudaStream_t st) {
  const int64_t total = planes * p.out_h * p.out_w;
  if (total == 0) return FM_OK;
  const bool sq = p.up_x == p.up_y && p.down_x == p.down_y;
  const int kmax = p.kh > p.kw ? p.kh : p.kw;
  // small images: shorter tiles keep more CTAs busy
  const bool small = p.out_h <= 16;
  if (sq && kmax <= 4) {
    const int up = p.up_x, down = p.down_x;
    if (up == 1 && down == 1)
      return small ? launch_tile<T, 1, 1, 4, 8>(out, x, kernel, p, planes, st)
                   : launch_tile<T, 1, 1, 4, 16>(out, x, kernel, p, planes, st);
    if (up == 2 && down == 1)
      return small ? launch_tile<T, 2, 1, 4, 8>(out, x, kernel, p, planes, st)
                   : launch_tile<T, 2, 1, 4, 16>(out, x, kernel, p, planes, st);
    if (up == 1 && down == 2) return launch_tile<T, 1, 2, 4, 8>(out, x, kernel, p, planes, st);
  }
  const int64_t want = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 32;
  upfirdn2d_generic_kernel<T><<<static_cast<unsigned>(want < cap ? want : cap), 256, 0, st>>>(
      static_cast<T*>(out), static_cast<const T*>(x), kernel, p, total);
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

}  // namespace fm

extern "C" int fm_upfirdn2d(void* out, const void* x, const float* kernel, int64_t planes, int in_h, int in_w, int kh,
                            int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0,
                            int pad_y1, int dtype, void* stream) {
  FM_CHECK_ARG(planes >= 0 && in_h >= 0 && in_w >= 0, "fm_upfirdn2d: negative size");
  FM_CHECK_ARG(kh >= 1 && kw >= 1, "fm_upfirdn2d: empty kernel");
  FM_CHECK_ARG(up_x >= 1 && up_y >= 1 && down_x >= 1 && down_y >= 1, "fm_upfirdn2d: up/down must be >= 1");
  fm::UpfirdnParams p{};
  p.in_h = in_h; p.in_w = in_w;
  p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
  p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y; p.kh = kh; p.kw = kw;
  // op/upfirdn2d_kernel.cu:237-240
  const int num_h = in_h * up_y + pad_y0 + pad_y1 - kh + down_y;
  const int num_w = in_w * up_x + pad_x0 + pad_x1 - kw + down_x;
  p.out_h = num_h > 0 ? num_h / down_y : 0;
  p.out_w = num_w > 0 ? num_w / down_x : 0;
  if (planes == 0 || p.out_h <= 0 || p.out_w <= 0) return FM_OK;
  FM_CHECK_ARG(out && x && kernel, "fm_upfirdn2d: null tensor");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case FM_F32: return fm::upfirdn2d_dispatch<float>(out, x, kernel, planes, p, st);
    case FM_F16: return fm::upfirdn2d_dispatch<__half>(out, x, kernel, planes, p, st);
    case FM_BF16: return fm::upfirdn2d_dispatch<__nv_bfloat16>(out, x, kernel, planes, p, st);
    default: fm::set_error("fm_upfirdn2d: bad dtype %d", dtype); return FM_ERR_INVALID;
  }
}
