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
#include <stdlib.h>

#include <type_traits>

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

// ------------------------------------------------------------------------------------
// Streaming kernel for up = down = 1 (every Blur of the generator and the discriminator:
// stylegan2.py:95-105 -> op/upfirdn2d.py:154 with pad (1,1) / (2,2) / (2,1)), kernel <= 4x4.
// A work item is a full-width strip of R output rows of one plane (or, for small images, a group of
// P whole planes): its input footprint is ONE contiguous byte range of the [planes, H, W] tensor, so a
// single thread fetches it with 1-D bulk async copies (cp.async.bulk, 16-byte aligned superset of the
// range) into a double-buffered shared-memory ring while the other threads filter the previous item.
// This is the only way to get wide loads here: rows are W*sizeof(T) bytes with W odd (2h+1), so
// neither vector loads nor tensor-map TMA (16-byte strides) apply.
// Compute: a thread owns one output column and walks down the rows with a rolling window of 4 partial
// sums (statically renamed, 4 rows per unrolled step): 4 conflict-free LDS, 8 FMAs (rank-1 taps) or 16,
// one coalesced store per row.  In strip mode the rows above / below the image are three zeroed rows in
// front of / behind the staged range, so the steady-state loop has no bounds test at all (the kernel is
// instruction-issue bound before it is HBM bound).  fp32 accumulation for every dtype.
// Round 2: the variant of the walk (column-tested or not) is chosen per WARP, whole planes are staged one bulk copy each
// between rows of zeros (so they take the untested walk too), the 2-byte types walk column pairs on aligned 32-bit
// words (ufs_rows_pk), and the per-item setup holds no division -- DESIGN.md 4.3 has the instruction counts.
// ------------------------------------------------------------------------------------
constexpr int UFS_THREADS = 256;
constexpr int UFS_MAX_SLOTS = 4;
constexpr int UFS_CHUNK = 16 * 1024;            // bytes per bulk copy
// ring geometry (slots x bytes per slot), tunable for experiments: FM3D_UFS_SLOTS / FM3D_UFS_SLOT_KB.
struct UfsRing { int slots, bytes; };
template <typename T>
static UfsRing ufs_ring(const UpfirdnParams& p) {
  static const int env_slots = []() { const char* e = getenv("FM3D_UFS_SLOTS"); return e ? atoi(e) : 0; }();
  static const int env_kb = []() { const char* e = getenv("FM3D_UFS_SLOT_KB"); return e ? atoi(e) : 0; }();
  // Measured on the largest generator blur (fp32 [4096,257,257], B200): 2 x 36 KB 534 us, 3 x 24 KB 698, 4 x 18 KB 924,
  // 3 x 36 KB (2 CTAs/SM) 670: a strip costs ~3.5k cycles of fixed work (halo rows, hand-over) on top of ~290 per
  // row, so fewer, taller strips beat a deeper ring.
  // Round 2 (after the walks were cut to 1/2 - 1/3 of their instructions): whole planes still do best with 2 x 36 KB
  // (3 CTAs per SM: bf16 129^2 3.89 TB/s vs 3.11 with 72 KB slots), but strips want to be tall -- 2 x 50 KB with up to 96
  // rows per strip (2 CTAs per SM): fp32 257^2 5.60 -> 6.19 TB/s, 129^2 4.68 -> 5.62, bf16 257^2 3.44 -> 4.08.
  const int64_t row_bytes = static_cast<int64_t>(p.in_w) * sizeof(T), plane_bytes = row_bytes * p.in_h;
  const int64_t zgap = (3 * row_bytes + 15) / 16 * 16 + 16;
  const bool whole = plane_bytes + 30 + 2 * zgap <= 36 * 1024 || plane_bytes + 32 <= 36 * 1024;
  UfsRing r{2, (whole ? 36 : 50) * 1024};
  if (env_slots >= 2 && env_slots <= UFS_MAX_SLOTS) r.slots = env_slots;
  if (env_kb >= 8 && env_kb <= 100) r.bytes = env_kb * 1024;
  return r;
}

struct UfsItem {
  int64_t plane0;      // first plane
  int nplanes;         // planes in the item (1 in strip mode)
  int oy0, oy1;        // output rows [oy0, oy1)
  int r_lo, r_hi;      // staged input rows [r_lo, r_hi] of each plane
};

__device__ __forceinline__ UfsItem ufs_item(int64_t i, const UpfirdnParams& p, int64_t planes, int R, int P, int strips) {
  UfsItem it;
  if (strips > 1) {
    // (a 64-bit division is ~100 instructions on every thread of every item)
    it.plane0 = (i >> 31) == 0 ? static_cast<int64_t>(static_cast<uint32_t>(i) / static_cast<uint32_t>(strips)) : i / strips;
    it.nplanes = 1;
    const int s = static_cast<int>(i - it.plane0 * strips);
    it.oy0 = s * R;
    it.oy1 = min(p.out_h, it.oy0 + R);
    it.r_lo = max(0, it.oy0 - p.pad_y0);
    it.r_hi = min(p.in_h - 1, it.oy1 - 1 - p.pad_y0 + 3);
  } else {
    it.plane0 = i * P;
    it.nplanes = static_cast<int>(min(static_cast<int64_t>(P), planes - it.plane0));
    it.oy0 = 0; it.oy1 = p.out_h;
    it.r_lo = 0; it.r_hi = p.in_h - 1;
  }
  return it;
}

// Row walk of one thread: COLS adjacent output columns, output rows [ty0, ty1).  Relative input row j
// (iy = ty0 - pad_y0 + j) feeds output row ty0 + j - ky with tap row ky; its partial sum sits in slot
// (j - ky) & 3 (statically renamed: 4 rows per unrolled step) and tap row 3 closes output row ty0 + j - 3.
template <typename T, int COLS, bool SEP, bool EDGE, bool ROWCHK>
__device__ __forceinline__ void ufs_rows(const T* __restrict__ sp, T* __restrict__ op, const float (&w)[4][4],
                                         const float (&kh1)[4], const float (&kv1)[4], const UpfirdnParams& p, int r_lo, int r_hi,
                                         int ix0, int ox, int ty0, int ty1) {
  bool cok[COLS + 3];
#pragma unroll
  for (int q = 0; q < COLS + 3; ++q) cok[q] = !EDGE || (ix0 + q >= 0 && ix0 + q < p.in_w);
  float acc[4][COLS];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < COLS; ++c) acc[j][c] = 0.f;
  const int jn = (ty1 - ty0) + 3;
  const T* rp = sp + static_cast<int64_t>(ty0 - p.pad_y0) * p.in_w + ix0;
  T* orow = op + static_cast<int64_t>(ty0 - 3) * p.out_w;
  int iy = ty0 - p.pad_y0;
  // one input row: horizontal taps, then its contribution to the 4 open output rows; u = j & 3 is static
  auto row_math = [&](auto uc, const float (&v)[COLS + 3]) {
    constexpr int u = decltype(uc)::value;
    if (SEP) {
#pragma unroll
      for (int c = 0; c < COLS; ++c) {
        const float h = fmaf(v[c + 3], kh1[3], fmaf(v[c + 2], kh1[2], fmaf(v[c + 1], kh1[1], v[c] * kh1[0])));
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) acc[(u - ky) & 3][c] = fmaf(h, kv1[ky], acc[(u - ky) & 3][c]);
      }
    } else {
#pragma unroll
      for (int ky = 0; ky < 4; ++ky)
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
          float a = acc[(u - ky) & 3][c];
#pragma unroll
          for (int kx = 0; kx < 4; ++kx) a = fmaf(v[c + kx], w[ky][kx], a);
          acc[(u - ky) & 3][c] = a;
        }
    }
  };
  // tap row 3 closed output row j - 3 (slot (u + 1) & 3): store it and re-open the slot
  auto row_store = [&](auto uc, T* o) {
    constexpr int u = decltype(uc)::value;
    float (&done)[COLS] = acc[(u + 1) & 3];
    if (COLS == 2 && sizeof(T) == 2 && (!EDGE || ox + 1 < p.out_w) && (reinterpret_cast<uintptr_t>(o) & 3) == 0) {
      T q2[2] = {from_f32<T>(done[0]), from_f32<T>(done[COLS - 1])};
      *reinterpret_cast<uint32_t*>(o) = *reinterpret_cast<uint32_t*>(q2);
    } else if (COLS == 2 && sizeof(T) == 4 && (!EDGE || ox + 1 < p.out_w) && (reinterpret_cast<uintptr_t>(o) & 7) == 0) {
      *reinterpret_cast<float2*>(o) = make_float2(done[0], done[COLS - 1]);
    } else {
#pragma unroll
      for (int c = 0; c < COLS; ++c)
        if (!EDGE || ox + c < p.out_w) o[c] = from_f32<T>(done[c]);
    }
  };
  auto row_load = [&](const T* r, int y, float (&v)[COLS + 3]) {
    const bool rok = !ROWCHK || (y >= r_lo && y <= r_hi);   // rows outside the image contribute nothing (or read zeroed rows)
#pragma unroll
    for (int q = 0; q < COLS + 3; ++q) v[q] = (rok && (!EDGE || cok[q])) ? to_f32<T>(r[q]) : 0.f;
  };
  int jb = 0;
  // steady state: 4 input rows per step with all their loads issued up front (no loop-carried branch between rows, so
  // the LDS latency of rows 1-3 hides behind the arithmetic of the rows before them)
  for (; jb + 4 <= jn; jb += 4) {
    float v0[COLS + 3], v1[COLS + 3], v2[COLS + 3], v3[COLS + 3];
    row_load(rp, iy, v0);
    row_load(rp + p.in_w, iy + 1, v1);
    row_load(rp + 2 * p.in_w, iy + 2, v2);
    row_load(rp + 3 * static_cast<int64_t>(p.in_w), iy + 3, v3);
    const bool st012 = jb >= 4;                               // the first step only closes a row at u = 3
    row_math(std::integral_constant<int, 0>{}, v0);
    if (st012) row_store(std::integral_constant<int, 0>{}, orow);
#pragma unroll
    for (int c = 0; c < COLS; ++c) acc[1][c] = 0.f;
    row_math(std::integral_constant<int, 1>{}, v1);
    if (st012) row_store(std::integral_constant<int, 1>{}, orow + p.out_w);
#pragma unroll
    for (int c = 0; c < COLS; ++c) acc[2][c] = 0.f;
    row_math(std::integral_constant<int, 2>{}, v2);
    if (st012) row_store(std::integral_constant<int, 2>{}, orow + 2 * p.out_w);
#pragma unroll
    for (int c = 0; c < COLS; ++c) acc[3][c] = 0.f;
    row_math(std::integral_constant<int, 3>{}, v3);
    row_store(std::integral_constant<int, 3>{}, orow + 3 * static_cast<int64_t>(p.out_w));
#pragma unroll
    for (int c = 0; c < COLS; ++c) acc[0][c] = 0.f;
    rp += 4 * static_cast<int64_t>(p.in_w);
    orow += 4 * static_cast<int64_t>(p.out_w);
    iy += 4;
  }
  // tail: up to 3 rows (u = 0, 1, 2)
  {
    float v[COLS + 3];
    if (jb < jn) {
      row_load(rp, iy, v);
      row_math(std::integral_constant<int, 0>{}, v);
      if (jb >= 3) row_store(std::integral_constant<int, 0>{}, orow);
#pragma unroll
      for (int c = 0; c < COLS; ++c) acc[1][c] = 0.f;
    }
    if (jb + 1 < jn) {
      row_load(rp + p.in_w, iy + 1, v);
      row_math(std::integral_constant<int, 1>{}, v);
      if (jb + 1 >= 3) row_store(std::integral_constant<int, 1>{}, orow + p.out_w);
#pragma unroll
      for (int c = 0; c < COLS; ++c) acc[2][c] = 0.f;
    }
    if (jb + 2 < jn) {
      row_load(rp + 2 * p.in_w, iy + 2, v);
      row_math(std::integral_constant<int, 2>{}, v);
      if (jb + 2 >= 3) row_store(std::integral_constant<int, 2>{}, orow + 2 * p.out_w);
#pragma unroll
      for (int c = 0; c < COLS; ++c) acc[3][c] = 0.f;
    }
  }
}

// Packed row walk for the 2-byte types (rank-1 taps, zero-row layouts): a thread owns TWO adjacent output columns.
// The kernel is bound by instruction issue, not by HBM: the walk above spends ~23 instructions per output on bf16
// (4 LDS.U16 + 4 conversions + 8 FMAs + a store per 2 bytes written), twice what the HBM rate leaves room for.  Here a
// row costs 3 aligned LDS.32 (the 5 inputs of the pair lie in 3 words whatever the parity of their address: rows are
// in_w * 2 bytes apart, so odd widths flip the parity every row -- the words are re-aligned with two funnel shifts by a
// per-thread 0 / 16, which rows j and j + 2 share), 5 unpacks, 8 FMAs for the two horizontal sums, 4 packed FFMA2 for
// the vertical window of the pair, one conversion and one 4-byte store: ~13 instructions per output.
// Edge columns AND the staged words with per-thread masks instead of testing loads.
__device__ __forceinline__ uint32_t ufs_lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
template <typename T> __device__ __forceinline__ void ufs_unpack2(uint32_t w, float& lo, float& hi);
template <> __device__ __forceinline__ void ufs_unpack2<__nv_bfloat16>(uint32_t w, float& lo, float& hi) {
  lo = __uint_as_float(w << 16);
  hi = __uint_as_float(w & 0xffff0000u);
}
template <> __device__ __forceinline__ void ufs_unpack2<__half>(uint32_t w, float& lo, float& hi) {
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
  lo = f.x; hi = f.y;
}
template <> __device__ __forceinline__ void ufs_unpack2<float>(uint32_t, float&, float&) {}
template <typename T> __device__ __forceinline__ uint32_t ufs_pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t ufs_pack2<__nv_bfloat16>(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
template <> __device__ __forceinline__ uint32_t ufs_pack2<__half>(float lo, float hi) {
  const __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
template <> __device__ __forceinline__ uint32_t ufs_pack2<float>(float, float) { return 0u; }

template <typename T, bool EDGE, bool VEC>
__device__ __forceinline__ void ufs_rows_pk(const T* __restrict__ sp, T* __restrict__ op, const float (&kh1)[4],
                                            const float (&kv1)[4], const UpfirdnParams& p, int ix0, int ox, int ty0, int ty1) {
  static_assert(sizeof(T) == 2, "packed walk: 2-byte types");
  uint32_t m0 = 0xffffffffu, m1 = 0xffffffffu, m2 = 0xffffffffu;
  if (EDGE) {
    auto ok = [&](int q) { return ix0 + q >= 0 && ix0 + q < p.in_w; };
    m0 = (ok(0) ? 0xffffu : 0u) | (ok(1) ? 0xffff0000u : 0u);
    m1 = (ok(2) ? 0xffffu : 0u) | (ok(3) ? 0xffff0000u : 0u);
    m2 = ok(4) ? 0xffffu : 0u;
  }
  f32x2 kv2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) kv2[i] = f2_pack(kv1[i], kv1[i]);
  f32x2 acc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = 0ull;
  const int jn = (ty1 - ty0) + 3;
  const uint32_t rb = static_cast<uint32_t>(p.in_w) * 2u;
  const uint32_t a0 = smem_u32(sp + static_cast<int64_t>(ty0 - p.pad_y0) * p.in_w + ix0);
  // rows j and j + 2 are 4 * in_w bytes apart: same parity
  uint32_t pA = a0 & ~3u, pB = (a0 + rb) & ~3u;
  const uint32_t sA = (a0 & 2u) << 3, sB = ((a0 + rb) & 2u) << 3;
  T* orow = op + static_cast<int64_t>(ty0 - 3) * p.out_w;
  // VEC (even output pitch, 4-byte aligned plane: warp-uniform, chosen by the caller): one 4-byte store per row.  A
  // run-time flag here left both store sequences in the loop as predicated instructions, which issue either way.
  const bool two = VEC || !EDGE || ox + 1 < p.out_w;
  struct Row { uint32_t w0, w1, w2; };
  auto row_load = [&](uint32_t a) { Row r; r.w0 = ufs_lds32(a); r.w1 = ufs_lds32(a + 4); r.w2 = ufs_lds32(a + 8); return r; };
  auto row_math = [&](auto uc, const Row& r, uint32_t sh) {
    constexpr int u = decltype(uc)::value;
    uint32_t n0 = __funnelshift_r(r.w0, r.w1, sh), n1 = __funnelshift_r(r.w1, r.w2, sh), n2 = r.w2 >> sh;
    if (EDGE) { n0 &= m0; n1 &= m1; n2 &= m2; }
    float e0, e1, e2, e3, e4, e5;
    ufs_unpack2<T>(n0, e0, e1);
    ufs_unpack2<T>(n1, e2, e3);
    ufs_unpack2<T>(n2, e4, e5);
    (void)e5;
    const float h0 = fmaf(e3, kh1[3], fmaf(e2, kh1[2], fmaf(e1, kh1[1], e0 * kh1[0])));
    const float h1 = fmaf(e4, kh1[3], fmaf(e3, kh1[2], fmaf(e2, kh1[1], e1 * kh1[0])));
    const f32x2 h = f2_pack(h0, h1);
    acc[u & 3] = f2_mul(kv2[0], h);                          // tap row 0 opens output row j: no zeroing pass
#pragma unroll
    for (int a = 1; a < 4; ++a) acc[(u - a) & 3] = f2_fma(kv2[a], h, acc[(u - a) & 3]);
  };
  auto row_store = [&](auto uc, T* o) {                      // tap row 3 closed output row j - 3 (slot (u + 1) & 3)
    constexpr int u = decltype(uc)::value;
    float lo, hi;
    f2_unpack(acc[(u + 1) & 3], lo, hi);
    if (VEC) {
      *reinterpret_cast<uint32_t*>(o) = ufs_pack2<T>(lo, hi);
    } else {
      o[0] = from_f32<T>(lo);
      if (two) o[1] = from_f32<T>(hi);
    }
  };
  const uint32_t rb2 = 2u * rb, rb4 = 4u * rb;
  int jb = 0;
  for (; jb + 4 <= jn; jb += 4) {
    const Row r0 = row_load(pA), r1 = row_load(pB), r2 = row_load(pA + rb2), r3 = row_load(pB + rb2);
    const bool st012 = jb >= 4;                               // the first step only closes a row at u = 3
    row_math(std::integral_constant<int, 0>{}, r0, sA);
    if (st012) row_store(std::integral_constant<int, 0>{}, orow);
    row_math(std::integral_constant<int, 1>{}, r1, sB);
    if (st012) row_store(std::integral_constant<int, 1>{}, orow + p.out_w);
    row_math(std::integral_constant<int, 2>{}, r2, sA);
    if (st012) row_store(std::integral_constant<int, 2>{}, orow + 2 * p.out_w);
    row_math(std::integral_constant<int, 3>{}, r3, sB);
    row_store(std::integral_constant<int, 3>{}, orow + 3 * static_cast<int64_t>(p.out_w));
    pA += rb4; pB += rb4;
    orow += 4 * static_cast<int64_t>(p.out_w);
  }
  if (jb < jn) {                                              // tail: up to 3 rows (u = 0, 1, 2)
    row_math(std::integral_constant<int, 0>{}, row_load(pA), sA);
    if (jb >= 3) row_store(std::integral_constant<int, 0>{}, orow);
  }
  if (jb + 1 < jn) {
    row_math(std::integral_constant<int, 1>{}, row_load(pB), sB);
    if (jb + 1 >= 3) row_store(std::integral_constant<int, 1>{}, orow + p.out_w);
  }
  if (jb + 2 < jn) {
    row_math(std::integral_constant<int, 2>{}, row_load(pA + rb2), sA);
    if (jb + 2 >= 3) row_store(std::integral_constant<int, 2>{}, orow + 2 * p.out_w);
  }
}

template <typename T, int COLS>
__global__ void __launch_bounds__(UFS_THREADS) upfirdn2d_stream_kernel(T* __restrict__ out, const T* __restrict__ x,
                                                                       const float* __restrict__ kernel, UpfirdnParams p,
                                                                       int64_t planes, int R, int P, int strips,
                                                                       int64_t n_items, int head, int nslots, int buf_bytes,
                                                                       int ppitch, int rot, int rs_log2, uint32_t cg_magic) {
  // head > 0: zero-row layouts -- rows above / below the image read zeros instead of being tested for.
  //  * strip mode (strips > 1): the staged range starts `head` bytes into the slot, preceded (top of the image) and
  //    followed (bottom) by three zeroed rows;
  //  * whole planes (ppitch > 0): every plane of the item is copied on its own into a region of `ppitch` bytes
  //    (first region `head` bytes into the slot), the gaps between regions hold >= 3 rows of zeros.  The ring is zeroed
  //    once; after a copy lands only the few bytes around each plane that the 16-byte granules of the copy (or the
  //    previous item, whose planes sat up to 15 bytes further along) dirtied are cleared again.
  // rot: column groups are handed out rotated by `rot` so that the groups touching the right edge share a warp with the
  // ones touching the left edge (a warp with any edge column runs the column-tested variant)
  extern __shared__ __align__(128) uint8_t ufs_smem[];
  __shared__ __align__(8) uint64_t s_bar[UFS_MAX_SLOTS];
  __shared__ float s_k[16];
  const int tid = threadIdx.x;
  if (tid < 16) {
    const int ky = tid >> 2, kx = tid & 3;
    // flipped: tap (ky,kx) of the correlation = k[kh-1-ky][kw-1-kx]  (op/upfirdn2d_kernel.cu:137)
    s_k[tid] = (ky < p.kh && kx < p.kw) ? kernel[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)] : 0.f;
  }
  if (tid == 0) {
    for (int i = 0; i < nslots; ++i) mbar_init(&s_bar[i], 1);
    fence_barrier_init();
  }
  if (ppitch > 0) {     // whole planes between zero rows: the ring starts out as zeros
    uint4* z = reinterpret_cast<uint4*>(ufs_smem);
    const int n16 = nslots * buf_bytes / 16;
    for (int i = tid; i < n16; i += UFS_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  float w[4][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i >> 2][i & 3] = s_k[i];
  // rank-1 test of the (flipped, zero-padded) taps: w == rowsum (x) colsum / total  ->  separable passes
  float kv1[4], kh1[4], tot = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    kv1[i] = w[i][0] + w[i][1] + w[i][2] + w[i][3];
    kh1[i] = w[0][i] + w[1][i] + w[2][i] + w[3][i];
    tot += kv1[i];
  }
  bool sep = fabsf(tot) > 1e-20f;
  const float inv_tot = sep ? 1.f / tot : 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) kh1[i] *= inv_tot;
#pragma unroll
  for (int i = 0; i < 16; ++i)
    sep = sep && fabsf(kv1[i >> 2] * kh1[i & 3] - w[i >> 2][i & 3]) <= 1e-6f * fabsf(tot);

  const int64_t plane_elems = static_cast<int64_t>(p.in_h) * p.in_w;
  const uintptr_t xbase = reinterpret_cast<uintptr_t>(x);
  const uint32_t plane_bytes = static_cast<uint32_t>(plane_elems * sizeof(T));     // (whole-plane modes: < one ring slot)
  const uint32_t plane_cap = (plane_bytes + 15u + 15u) & ~15u;                      // most bytes one plane's copy can write

  // byte range of an item, widened to 16-byte boundaries (the extra bytes stay inside the 16-byte
  // granules that hold the first / last wanted byte, so they are always mapped)
  auto issue = [&](const UfsItem& it, int slot) {
    if (ppitch > 0) {
      fence_proxy_async();                     // earlier generic accesses of this slot precede the async writes
      uint32_t total = 0;
      for (int q = 0; q < it.nplanes; ++q) {
        const uintptr_t lo = xbase + static_cast<uintptr_t>(it.plane0 + q) * plane_bytes;
        total += static_cast<uint32_t>(((lo + plane_bytes + 15) & ~static_cast<uintptr_t>(15)) - (lo & ~static_cast<uintptr_t>(15)));
      }
      mbar_arrive_expect_tx(&s_bar[slot], total);
      for (int q = 0; q < it.nplanes; ++q) {
        const uintptr_t lo = xbase + static_cast<uintptr_t>(it.plane0 + q) * plane_bytes;
        const uintptr_t lo_a = lo & ~static_cast<uintptr_t>(15);
        const uint32_t bytes = static_cast<uint32_t>(((lo + plane_bytes + 15) & ~static_cast<uintptr_t>(15)) - lo_a);
        uint8_t* dst = ufs_smem + slot * buf_bytes + head + q * ppitch;
        for (uint32_t off = 0; off < bytes; off += UFS_CHUNK)
          bulk_load_1d(dst + off, reinterpret_cast<const void*>(lo_a + off), min(static_cast<uint32_t>(UFS_CHUNK), bytes - off),
                       &s_bar[slot]);
      }
      return;
    }
    const uintptr_t lo = xbase + static_cast<uintptr_t>((it.plane0 * plane_elems + static_cast<int64_t>(it.r_lo) * p.in_w) * sizeof(T));
    const uintptr_t hi = xbase + static_cast<uintptr_t>(((it.plane0 + it.nplanes - 1) * plane_elems +
                                                         static_cast<int64_t>(it.r_hi + 1) * p.in_w) * sizeof(T));
    const uintptr_t lo_a = lo & ~static_cast<uintptr_t>(15);
    const uint32_t bytes = static_cast<uint32_t>(((hi + 15) & ~static_cast<uintptr_t>(15)) - lo_a);
    uint8_t* dst = ufs_smem + slot * buf_bytes + head;
    fence_proxy_async();                       // earlier generic reads of this slot precede the async writes
    mbar_arrive_expect_tx(&s_bar[slot], bytes);
    for (uint32_t off = 0; off < bytes; off += UFS_CHUNK)
      bulk_load_1d(dst + off, reinterpret_cast<const void*>(lo_a + off), min(static_cast<uint32_t>(UFS_CHUNK), bytes - off),
                   &s_bar[slot]);
  };

  const int colgroups = (p.out_w + COLS - 1) / COLS;
  int64_t item = blockIdx.x;
  if (tid == 0)
    for (int a = 0; a < nslots - 1; ++a)
      if (item + static_cast<int64_t>(a) * gridDim.x < n_items)
        issue(ufs_item(item + static_cast<int64_t>(a) * gridDim.x, p, planes, R, P, strips), a);
  int slot = 0, ahead_slot = nslots - 1;
  uint32_t parity = 0;
  // strip mode: (plane, strip) of this CTA's items advance by a fixed step -- no division per item
  const uint32_t step_q = strips > 1 ? gridDim.x / static_cast<uint32_t>(strips) : 0u;
  const int step_r = strips > 1 ? static_cast<int>(gridDim.x - step_q * strips) : 0;
  int64_t cur_plane = 0;
  int cur_strip = 0;
  if (strips > 1) {
    const UfsItem f = ufs_item(item, p, planes, R, P, strips);
    cur_plane = f.plane0;
    cur_strip = f.oy0 / R;
  }
  for (int k = 0; item < n_items; item += gridDim.x, ++k) {
    UfsItem it;
    if (strips > 1) {
      it.plane0 = cur_plane; it.nplanes = 1;
      it.oy0 = cur_strip * R;
      it.oy1 = min(p.out_h, it.oy0 + R);
      it.r_lo = max(0, it.oy0 - p.pad_y0);
      it.r_hi = min(p.in_h - 1, it.oy1 - 1 - p.pad_y0 + 3);
      cur_plane += step_q; cur_strip += step_r;
      if (cur_strip >= strips) { cur_strip -= strips; ++cur_plane; }
    } else {
      it = ufs_item(item, p, planes, R, P, strips);
    }
    // the slot refilled now was drained by the previous iteration (its closing __syncthreads)
    const int64_t ahead = item + static_cast<int64_t>(nslots - 1) * gridDim.x;
    if (tid == 0 && ahead < n_items) issue(ufs_item(ahead, p, planes, R, P, strips), ahead_slot);
    if (tid < 32) {                                           // one warp polls, the rest sleep in the barrier
      mbar_wait(&s_bar[slot], parity);
      if (ppitch > 0) {
        // clear what the copies (16-byte granules) and the previous item left outside [plane start, plane end)
        for (int q = 0; q < it.nplanes; ++q) {
          const uint32_t off = static_cast<uint32_t>((xbase + static_cast<uintptr_t>(it.plane0 + q) * plane_bytes) & 15);
          uint8_t* rb = ufs_smem + slot * buf_bytes + head + q * ppitch;
          T* a = reinterpret_cast<T*>(rb);
          T* b = reinterpret_cast<T*>(rb + off + plane_bytes);
          const int na = static_cast<int>(off / sizeof(T)), nb = static_cast<int>((plane_cap - off - plane_bytes) / sizeof(T));
          if (tid < na) a[tid] = from_f32<T>(0.f);
          if (tid < nb) b[tid] = from_f32<T>(0.f);
        }
      }
    }
    __syncthreads();

    const uintptr_t lo = xbase + static_cast<uintptr_t>((it.plane0 * plane_elems + static_cast<int64_t>(it.r_lo) * p.in_w) * sizeof(T));
    const T* sbuf = reinterpret_cast<const T*>(ufs_smem + slot * buf_bytes + head + (lo & 15));
    if (head > 0 && ppitch == 0) {
      const bool top = it.oy0 - p.pad_y0 < it.r_lo, bottom = it.oy1 - 1 - p.pad_y0 + 3 > it.r_hi;
      if (top || bottom) {                                    // uniform: only the first / last strip of a plane
        T* zb = const_cast<T*>(sbuf);
        const int n3 = 3 * p.in_w;
        if (top) for (int i = tid; i < n3; i += UFS_THREADS) zb[i - n3] = from_f32<T>(0.f);
        if (bottom) {
          T* zt = zb + static_cast<int64_t>(it.r_hi - it.r_lo + 1) * p.in_w;
          for (int i = tid; i < n3; i += UFS_THREADS) zt[i] = from_f32<T>(0.f);
        }
        __syncthreads();
      }
    }
    // thread tasks: (plane of the item, column group, row split), handed out warp by warp so that the choice of the
    // row-walk variant is warp-uniform: a warp whose lanes disagreed (column 0 is an edge column, 1..31 are not) used
    // to run BOTH variants one after the other -- every warp of a 64-wide image, at 2-3x the instructions per output
    // (the row split is a power of two chosen by the host and column-group indices come from a multiply-high with
    // ceil(2^32 / colgroups), exact for task < 2^16: the five integer divisions this block used to hold were a
    // quarter of all instructions on the 2-byte types)
    const int nrows = it.oy1 - it.oy0;
    const int ncg = it.nplanes * colgroups;
    const int rs = 1 << rs_log2;
    const int rows_per = (nrows + rs - 1) >> rs_log2;
    const int ntask = ncg * rs;
    for (int tb = tid & ~31; tb < ntask; tb += UFS_THREADS) {
      const int task = tb + (tid & 31);
      const int t2 = cg_magic ? static_cast<int>(__umulhi(static_cast<uint32_t>(task), cg_magic)) : task;   // task / colgroups
      int part = 0, pl = t2;                                                                 // t2 = part * nplanes + pl
      if (rs_log2 > 0) {
        if (pl >= 2 * it.nplanes) { pl -= 2 * it.nplanes; part = 2; }
        if (pl >= it.nplanes) { pl -= it.nplanes; part += 1; }
      }
      int cgx = task - t2 * colgroups - rot;
      cgx += cgx < 0 ? colgroups : 0;
      const int ox = cgx * COLS;
      const int ty0 = it.oy0 + part * rows_per;
      const int ty1 = min(it.oy1, ty0 + rows_per);
      const bool valid = task < ntask && ty0 < ty1;
      const T* sp;                                                                                   // row iy at sp + iy*in_w
      if (ppitch > 0) {
        const uint32_t off = static_cast<uint32_t>((xbase + static_cast<uintptr_t>(it.plane0 + pl) * plane_bytes) & 15);
        sp = reinterpret_cast<const T*>(ufs_smem + slot * buf_bytes + head + pl * ppitch + off);
      } else {
        sp = sbuf + static_cast<int64_t>(pl) * plane_elems - static_cast<int64_t>(it.r_lo) * p.in_w;
      }
      const int ix0 = ox - p.pad_x0;
      T* op = out + ((it.plane0 + pl) * p.out_h) * static_cast<int64_t>(p.out_w) + ox;
      const bool interior = ix0 >= 0 && ix0 + COLS + 3 <= p.in_w && ox + COLS <= p.out_w;
      const bool warp_interior = __all_sync(0xffffffffu, interior || !valid);
      if (!valid) continue;
      // instantiations of the row walk: rank-1 taps take 8 instead of 16 FMAs per output, interior warps
      // skip every column test, the zero-row layouts skip every row test
#define UFS_CALL(SEP_, EDGE_, RC_) \
  ufs_rows<T, COLS, SEP_, EDGE_, RC_>(sp, op, w, kh1, kv1, p, it.r_lo, it.r_hi, ix0, ox, ty0, ty1)
      if (sizeof(T) == 2 && COLS == 2 && head > 0 && sep) {
        if constexpr (sizeof(T) == 2 && COLS == 2) {
          // even output pitch and an even number of outputs per plane: every column pair of every row is 4-byte aligned
          // when the tensor is (kernel-uniform)
          const bool vec = (p.out_w & 1) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0;
          if (vec) {
            if (warp_interior) ufs_rows_pk<T, false, true>(sp, op, kh1, kv1, p, ix0, ox, ty0, ty1);
            else ufs_rows_pk<T, true, true>(sp, op, kh1, kv1, p, ix0, ox, ty0, ty1);
          } else {
            if (warp_interior) ufs_rows_pk<T, false, false>(sp, op, kh1, kv1, p, ix0, ox, ty0, ty1);
            else ufs_rows_pk<T, true, false>(sp, op, kh1, kv1, p, ix0, ox, ty0, ty1);
          }
        }
      } else if (head > 0) {
        if (sep) { if (warp_interior) UFS_CALL(true, false, false); else UFS_CALL(true, true, false); }
        else     { if (warp_interior) UFS_CALL(false, false, false); else UFS_CALL(false, true, false); }
      } else {
        if (sep) { if (warp_interior) UFS_CALL(true, false, true); else UFS_CALL(true, true, true); }
        else     { if (warp_interior) UFS_CALL(false, false, true); else UFS_CALL(false, true, true); }
      }
#undef UFS_CALL
    }
    __syncthreads();      // everyone is done with this slot before it is refilled (next iteration)
    ahead_slot = slot;
    if (++slot == nslots) { slot = 0; parity ^= 1; }
  }
}

template <typename T>
static bool stream_eligible(const UpfirdnParams& p, int pad_x1, int pad_y1) {
  if (p.up_x != 1 || p.up_y != 1 || p.down_x != 1 || p.down_y != 1) return false;
  if (p.kh > 4 || p.kw > 4) return false;
  if (p.pad_x0 < 0 || p.pad_y0 < 0 || pad_x1 < 0 || pad_y1 < 0 || p.pad_x0 > 16 || p.pad_y0 > 16) return false;
  // at least 4 output rows (+3 halo rows, +6 zero rows, alignment slack) of a full-width strip must fit one ring slot
  return static_cast<int64_t>(14) * p.in_w * static_cast<int64_t>(sizeof(T)) + 64 <= ufs_ring<T>(p).bytes;
}

template <typename T, int COLS>
static int launch_stream_c(void* out, const void* x, const float* kernel, const UpfirdnParams& p, int pad_y1, int64_t planes,
                           cudaStream_t st) {
  const int64_t row_bytes = static_cast<int64_t>(p.in_w) * sizeof(T);
  const int64_t plane_bytes = row_bytes * p.in_h;
  int R, P, strips, head = 0;
  const UfsRing ring = ufs_ring<T>(p);
  const int UFS_BUF_BYTES = ring.bytes, nslots = ring.slots;
  int ppitch = 0;
  // the walk always spans 4 tap rows: with pad_y0 <= 3 and pad_y1 < kh it never leaves three rows of zeros above / below
  const bool zrows = p.pad_y0 <= 3 && pad_y1 < p.kh;
  static const int env_zplanes = []() { const char* e = getenv("FM3D_UFS_ZPLANES"); return e ? atoi(e) : 1; }();
  const int64_t zgap = (3 * row_bytes + 15) / 16 * 16 + 16;                      // >= 3 rows after the furthest plane end
  const int64_t zpitch = (plane_bytes + 15 + 15) / 16 * 16 + zgap;               // copy capacity + gap
  if (zrows && env_zplanes && plane_bytes >= 2048 && zgap + zpitch <= UFS_BUF_BYTES) {
    // whole planes between zero rows (see the kernel): [gap | plane 0 | gap | plane 1 | gap ...]
    strips = 1; R = p.out_h;
    head = static_cast<int>(zgap);
    ppitch = static_cast<int>(zpitch);
    P = static_cast<int>((UFS_BUF_BYTES - zgap) / zpitch);
    const int64_t want_items = static_cast<int64_t>(sm_count()) * 6;
    while (P > 1 && (planes + P - 1) / P < want_items) P >>= 1;
    while (P > 1 && static_cast<int64_t>(P) * ((p.out_w + COLS - 1) / COLS) * 4 >= 65536) P >>= 1;   // task index < 2^16
  } else if (plane_bytes + 32 <= UFS_BUF_BYTES) {
    strips = 1; R = p.out_h;
    P = static_cast<int>((UFS_BUF_BYTES - 32) / plane_bytes);
    // keep enough items to fill the machine
    const int64_t want_items = static_cast<int64_t>(sm_count()) * 6;
    while (P > 1 && (planes + P - 1) / P < want_items) P >>= 1;
    while (P > 1 && static_cast<int64_t>(P) * ((p.out_w + COLS - 1) / COLS) * 4 >= 65536) P >>= 1;   // task index < 2^16
  } else {
    P = 1;
    // zero-row layout when the vertical padding fits three rows: [3 zero rows | staged rows | 3 zero rows]
    // (the last output row reads pad_y1 - kh + 4 rows past the image)
    head = zrows ? static_cast<int>((3 * row_bytes + 15) / 16 * 16 + 16) : 0;
    R = static_cast<int>((UFS_BUF_BYTES - head - 32) / row_bytes) - 3 - (zrows ? 3 : 0);
    static const int env_rmax = []() { const char* e = getenv("FM3D_UFS_RMAX"); return e ? atoi(e) : 0; }();
    const int rmax = env_rmax > 0 ? env_rmax : 96;
    if (R > rmax) R = rmax;
    strips = (p.out_h + R - 1) / R;
    R = (p.out_h + strips - 1) / strips;          // balance the strips
  }
  const int64_t n_items = strips > 1 ? planes * strips : (planes + P - 1) / P;
  static SmemOptIn opt_in;        // per instantiation (T, COLS), per device
  auto fn = upfirdn2d_stream_kernel<T, COLS>;
  const int smem = nslots * UFS_BUF_BYTES;
  FM_CUDA_OK(smem_opt_in(opt_in, fn, 200 * 1024));
  int cps = (227 * 1024) / (smem + 2048);          // CTAs per SM that fit (persistent grid)
  cps = cps < 1 ? 1 : (cps > 4 ? 4 : cps);
  const int64_t cap = static_cast<int64_t>(sm_count()) * cps;
  const unsigned grid = static_cast<unsigned>(n_items < cap ? n_items : cap);
  // column groups that touch the right edge (their taps read past the row, or the group is ragged): handed out first,
  // next to the left-edge groups
  const int colgroups = (p.out_w + COLS - 1) / COLS;
  int last_int = (p.in_w + p.pad_x0 - COLS - 3) / COLS;             // last group with ix0 + COLS + 3 <= in_w
  if (p.in_w + p.pad_x0 - COLS - 3 < 0) last_int = -1;
  if (last_int > p.out_w / COLS - 1) last_int = p.out_w / COLS - 1;  // ... and ox + COLS <= out_w
  int rot = colgroups - 1 - last_int;
  rot = rot < 0 ? 0 : (rot >= colgroups ? 0 : rot);
  // rows of an item split between 1, 2 or 4 thread groups when its column groups leave threads idle
  static const int env_rs = []() { const char* e = getenv("FM3D_UFS_RS"); return e ? atoi(e) : 4; }();
  int rs_log2 = 0;
  // (each part re-walks 3 warm-up rows: measured, a split pays only while fewer than half of the threads have a task --
  // fp32 65^2 5.33 -> 5.67 TB/s, 129^2 5.39 -> 5.68 without the needless split, bf16 129^2 2.87 -> 3.89 with the needed one)
  static const int env_busy = []() { const char* e = getenv("FM3D_UFS_BUSY"); return e ? atoi(e) : 0; }();
  // ... except 4-byte strips (2 CTAs per SM, a cheap walk): 129^2 fp32 5.23 TB/s with 128 busy threads, 5.62 with 256
  const int busy_min = env_busy > 0 ? env_busy : (strips > 1 && sizeof(T) == 4 ? 192 : UFS_THREADS / 2);
  while (rs_log2 < 2 && (2 << rs_log2) <= env_rs && static_cast<int64_t>(P) * colgroups * (1 << rs_log2) < busy_min &&
         static_cast<int64_t>(P) * colgroups * (2 << rs_log2) <= UFS_THREADS) ++rs_log2;
  FM_CHECK_ARG(static_cast<int64_t>(P) * colgroups * 4 < 65536, "fm_upfirdn2d: too many thread tasks per item");
  const uint32_t cg_magic = colgroups > 1 ? static_cast<uint32_t>((0x100000000ull + colgroups - 1) / colgroups) : 0u;   // 0: identity
  fn<<<grid, UFS_THREADS, smem, st>>>(static_cast<T*>(out), static_cast<const T*>(x), kernel, p, planes, R, P, strips,
                                      n_items, head, nslots, UFS_BUF_BYTES, ppitch, rot, rs_log2, cg_magic);
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

// columns per thread: 1 = one thread per output column; 2 = column pairs (5 LDS + 16 FMA + one 8-byte store per two
// outputs) with the strip's rows split between two thread groups (FM3D_UFS_COLS)
template <typename T>
static int launch_stream(void* out, const void* x, const float* kernel, const UpfirdnParams& p, int pad_y1, int64_t planes,
                         cudaStream_t st) {
  // 2-byte types: column pairs, so that rank-1 taps in a zero-row layout take the packed walk (ufs_rows_pk)
  static const int env_cols = []() { const char* e = getenv("FM3D_UFS_COLS"); return e ? atoi(e) : 0; }();
  const int cols = env_cols ? env_cols : (sizeof(T) == 2 ? 2 : 1);
  if (cols == 2) return launch_stream_c<T, 2>(out, x, kernel, p, pad_y1, planes, st);
  return launch_stream_c<T, 1>(out, x, kernel, p, pad_y1, planes, st);
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

// ------------------------------------------------------------------------------------
// up = 2, down = 1 (Upsample of the ToRGB skip path, stylegan2.py:65-80 -> op/upfirdn2d.py with up=2, pad (2,1)):
// polyphase form.  Of the 4 x 4 taps only the 2 x 2 whose row / column parity matches the zero-stuffed grid meet a
// sample, so an output is 4 FMAs on a 2 x 2 source patch.  A thread owns 2 output rows x 4 output columns = a 3 x 4
// source patch read straight through L1 (the tensors are 3-channel images: the whole input is a few MB and every
// source sample is used by 4 threads of neighbouring lanes), no shared-memory tile, no barrier; the parities of the two
// pads are template parameters so every tap / patch index is static.  The tile kernel above spent its time on the
// per-tap modulo tests and on staging with two barriers per 2048 outputs: 35 us for [96,128,128] -> [96,256,256].
// ------------------------------------------------------------------------------------
template <typename T, int PY, int PX>
__global__ void __launch_bounds__(256) upfirdn2d_up2_kernel(T* __restrict__ out, const T* __restrict__ x,
                                                            const float* __restrict__ kernel, UpfirdnParams p,
                                                            uint32_t total, FastDivU32 fd_cq, FastDivU32 fd_rp) {
  __shared__ float s_k[16];
  if (threadIdx.x < 16) {
    const int ky = threadIdx.x >> 2, kx = threadIdx.x & 3;
    // flipped, zero-padded to 4 x 4  (op/upfirdn2d_kernel.cu:137)
    s_k[threadIdx.x] = (ky < p.kh && kx < p.kw) ? kernel[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)] : 0.f;
  }
  __syncthreads();
  float kf[4][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) kf[i >> 2][i & 3] = s_k[i];
  const int yoff = (PY - p.pad_y0) >> 1, xoff = (PX - p.pad_x0) >> 1;      // PY - pad_y0, PX - pad_x0 are even
  const bool vec = (p.out_w & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & (4 * sizeof(T) - 1)) == 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t t = fastdiv(i, fd_cq);
    const int cq = static_cast<int>(i - t * fd_cq.d);
    const uint32_t plane = fastdiv(t, fd_rp);
    const int rp = static_cast<int>(t - plane * fd_rp.d);
    const int iy0 = rp + yoff, ix0 = 2 * cq + xoff;
    const T* xp = x + static_cast<int64_t>(plane) * p.in_h * p.in_w;
    float v[3][4];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int iy = iy0 + r, ix = ix0 + c;
        v[r][c] = (iy >= 0 && iy < p.in_h && ix >= 0 && ix < p.in_w) ? to_f32<T>(__ldg(xp + static_cast<int64_t>(iy) * p.in_w + ix)) : 0.f;
      }
    T* op = out + (static_cast<int64_t>(plane) * p.out_h + 2 * rp) * p.out_w + 4 * cq;
#pragma unroll
    for (int tr = 0; tr < 2; ++tr) {
      if (2 * rp + tr >= p.out_h) break;
      const int ay = tr == 0 ? PY : 1 - PY;            // first matching tap row; the other one is ay + 2
      const int rb = tr == 0 ? 0 : 1 - PY;             // patch row of tap ay
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ax = (PX - j) & 1;
        const int cb = (j + ax - PX) >> 1;
        o[j] = v[rb][cb] * kf[ay][ax];
        o[j] = fmaf(v[rb][cb + 1], kf[ay][ax + 2], o[j]);
        o[j] = fmaf(v[rb + 1][cb], kf[ay + 2][ax], o[j]);
        o[j] = fmaf(v[rb + 1][cb + 1], kf[ay + 2][ax + 2], o[j]);
      }
      T* orow = op + static_cast<int64_t>(tr) * p.out_w;
      if (vec) {
        if (sizeof(T) == 4) {
          *reinterpret_cast<float4*>(orow) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
          T q[4] = {from_f32<T>(o[0]), from_f32<T>(o[1]), from_f32<T>(o[2]), from_f32<T>(o[3])};
          *reinterpret_cast<uint2*>(orow) = *reinterpret_cast<uint2*>(q);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (4 * cq + j < p.out_w) orow[j] = from_f32<T>(o[j]);
      }
    }
  }
}

template <typename T>
static int launch_up2(void* out, const void* x, const float* kernel, const UpfirdnParams& p, int64_t planes, cudaStream_t st) {
  const uint32_t cqs = static_cast<uint32_t>((p.out_w + 3) / 4), rps = static_cast<uint32_t>((p.out_h + 1) / 2);
  const uint32_t total = static_cast<uint32_t>(planes) * cqs * rps;
  const uint32_t want = (total + 255) / 256;
  const uint32_t cap = static_cast<uint32_t>(sm_count()) * 64;
  const unsigned grid = want < cap ? want : cap;
  const FastDivU32 fc = fastdiv_make(cqs), fr = fastdiv_make(rps);
  const int py = p.pad_y0 & 1, px = p.pad_x0 & 1;
#define FM_UP2(PY_, PX_) upfirdn2d_up2_kernel<T, PY_, PX_><<<grid, 256, 0, st>>>(static_cast<T*>(out), static_cast<const T*>(x), kernel, p, total, fc, fr)
  if (py == 0 && px == 0) FM_UP2(0, 0);
  else if (py == 0) FM_UP2(0, 1);
  else if (px == 0) FM_UP2(1, 0);
  else FM_UP2(1, 1);
#undef FM_UP2
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

template <typename T>
static int upfirdn2d_dispatch(void* out, const void* x, const float* kernel, int64_t planes, UpfirdnParams p, int pad_x1,
                              int pad_y1, cudaStream_t st) {
  const int64_t total = planes * p.out_h * p.out_w;
  if (total == 0) return FM_OK;
  static const int env_stream = []() { const char* e = getenv("FM3D_UPFIRDN_STREAM"); return e ? atoi(e) : 1; }();
  if (env_stream && stream_eligible<T>(p, pad_x1, pad_y1)) return launch_stream<T>(out, x, kernel, p, pad_y1, planes, st);
  const bool sq = p.up_x == p.up_y && p.down_x == p.down_y;
  const int kmax = p.kh > p.kw ? p.kh : p.kw;
  static const int env_up2 = []() { const char* e = getenv("FM3D_UPFIRDN_UP2"); return e ? atoi(e) : 1; }();
  if (env_up2 && sq && p.up_x == 2 && p.down_x == 1 && kmax <= 4 &&
      planes * ((p.out_w + 3) / 4) * static_cast<int64_t>((p.out_h + 1) / 2) < 0x7FFFFFFF)
    return launch_up2<T>(out, x, kernel, p, planes, st);
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
    case FM_F32: return fm::upfirdn2d_dispatch<float>(out, x, kernel, planes, p, pad_x1, pad_y1, st);
    case FM_F16: return fm::upfirdn2d_dispatch<__half>(out, x, kernel, planes, p, pad_x1, pad_y1, st);
    case FM_BF16: return fm::upfirdn2d_dispatch<__nv_bfloat16>(out, x, kernel, planes, p, pad_x1, pad_y1, st);
    default: fm::set_error("fm_upfirdn2d: bad dtype %d", dtype); return FM_ERR_INVALID;
  }
}
