// fused bias + activation (+ gradient modes), HBM-bound elementwise kernel for sm_100a.
//
// Replaces op/fused_bias_act_kernel.cu:18-49 of the reference: same arithmetic
//   y = act(x + b[c]) * scale          (grad 0)
//   y = (ref > 0 ? x : alpha * x) * scale   (grad 1, x may carry a bias: double-backward)
//   y = 0                                (grad 2)
// but with 16-byte vector accesses, several independent loads in flight per thread,
// 64-bit-safe indexing and one integer divide per vector instead of div+mod per element.
//
// Algorithmic bytes per element: 2*sizeof(T) forward, 3*sizeof(T) in gradient mode
// (SURVEY.md 8d); roofline = HBM.
#include "common.cuh"

namespace fm {

template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); };

template <typename T, int N>
struct alignas(16) Pack { T v[N]; };

template <typename T>
__device__ __forceinline__ Pack<T, Vec16<T>::N> ld16(const T* p) {
  Pack<T, Vec16<T>::N> r;
  *reinterpret_cast<uint4*>(&r) = __ldg(reinterpret_cast<const uint4*>(p));
  return r;
}
template <typename T>
__device__ __forceinline__ void st16(T* p, const Pack<T, Vec16<T>::N>& v) {
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&v);
}

__device__ __forceinline__ float act_apply(float x, float ref, int mode, float alpha) {
  // mode = act*10+grad, op/fused_bias_act_kernel.cu:36-45
  switch (mode) {
    case 30: return x > 0.f ? x : x * alpha;
    case 31: return ref > 0.f ? x : x * alpha;
    case 12:
    case 32: return 0.f;
    default: return x;  // 10, 11 and the reference's "default:" label
  }
}

// Vector kernel: `inner` is a multiple of the vector width, so a 16-byte vector never
// straddles a channel.  IdxT is uint32_t when the vector count fits, else uint64_t.
template <typename T, typename IdxT, int MODE, bool HAS_BIAS, bool HAS_REF, int UNROLL>
__global__ void __launch_bounds__(256) bias_act_vec_kernel(T* __restrict__ out, const T* __restrict__ x,
                                                           const T* __restrict__ bias, const T* __restrict__ ref,
                                                           IdxT n_vec, IdxT vec_per_row, uint32_t channels,
                                                           float alpha, float scale, FastDivU32 fd_row, FastDivU32 fd_ch) {
  constexpr int N = Vec16<T>::N;
  const bool alpha01 = alpha >= 0.f && alpha <= 1.f;
  const f32x2 aa = f2_pack(alpha, alpha), ss = f2_pack(scale, scale);
  // A warp owns chunks of 32 * UNROLL consecutive vectors (UNROLL coalesced 512-byte requests per warp and tensor).  When
  // a row (one channel of one sample) is a whole number of chunks, the channel -- and with it the bias -- is one
  // computation per chunk instead of one per vector; per-vector index arithmetic was as many instructions as the math.
  constexpr uint32_t CHUNK = 32u * UNROLL;
  const uint32_t lane = threadIdx.x & 31u, wpb = blockDim.x >> 5;
  const bool row_uniform = vec_per_row % CHUNK == 0;
  auto channel_of = [&](IdxT v) -> uint32_t {
    if (sizeof(IdxT) == 4) {
      const uint32_t row = fastdiv(static_cast<uint32_t>(v), fd_row);
      return row - fastdiv(row, fd_ch) * channels;
    }
    return static_cast<uint32_t>((v / vec_per_row) % channels);
  };
  const IdxT nchunk = (n_vec + CHUNK - 1) / CHUNK;
  for (IdxT g = static_cast<IdxT>(blockIdx.x) * wpb + (threadIdx.x >> 5); g < nchunk; g += static_cast<IdxT>(gridDim.x) * wpb) {
    const IdxT v0 = g * CHUNK + lane;
    Pack<T, N> xv[UNROLL], rv[UNROLL];
    float bv[UNROLL];
    float b_chunk = 0.f;
    if (HAS_BIAS && row_uniform) b_chunk = to_f32<T>(__ldg(bias + channel_of(g * CHUNK)));
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const IdxT v = v0 + 32u * u;
      if (v < n_vec) {
        xv[u] = ld16<T>(x + static_cast<size_t>(v) * N);
        if (HAS_REF) rv[u] = ld16<T>(ref + static_cast<size_t>(v) * N);
        if (HAS_BIAS) {
          if (row_uniform) { bv[u] = b_chunk; continue; }
          const uint32_t c = channel_of(v);
          bv[u] = to_f32<T>(__ldg(bias + c));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const IdxT v = v0 + 32u * u;
      if (v < n_vec) {
        Pack<T, N> o;
        if (MODE == 30 && alpha01) {
          // forward leaky ReLU with 0 <= alpha <= 1: x > 0 ? x : alpha * x  ==  max(x, alpha * x) bit for bit (including
          // -0 and NaN), so element pairs run on packed fp32 (FADD2 / FMUL2): ~4 instead of ~12 instructions per element
          // for the 2-byte types, which were bound by instruction issue (bf16 forward 4.7 TB/s against 6.4 for the
          // bias-free gradient pass with its select-only arithmetic)
          const f32x2 bb = f2_pack(HAS_BIAS ? bv[u] : 0.f, HAS_BIAS ? bv[u] : 0.f);
#pragma unroll
          for (int j = 0; j < N; j += 2) {
            f32x2 xx = f2_pack(to_f32<T>(xv[u].v[j]), to_f32<T>(xv[u].v[j + 1]));
            if (HAS_BIAS) xx = f2_add(xx, bb);
            const f32x2 ax = f2_mul(xx, aa);
            float xl, xh, al, ah;
            f2_unpack(xx, xl, xh);
            f2_unpack(ax, al, ah);
            const f32x2 y = f2_mul(f2_pack(fmaxf(xl, al), fmaxf(xh, ah)), ss);
            float yl, yh;
            f2_unpack(y, yl, yh);
            o.v[j] = from_f32<T>(yl);
            o.v[j + 1] = from_f32<T>(yh);
          }
        } else
#pragma unroll
        for (int j = 0; j < N; ++j) {
          float xf = to_f32<T>(xv[u].v[j]);
          if (MODE == 40) { o.v[j] = from_f32<T>(xf * bv[u]); continue; }     // fm_channel_scale: x * s[row]
          if (HAS_BIAS) xf += bv[u];
          const float rf = HAS_REF ? to_f32<T>(rv[u].v[j]) : 0.f;
          o.v[j] = from_f32<T>(act_apply(xf, rf, MODE, alpha) * scale);
        }
        st16<T>(out + static_cast<size_t>(v) * N, o);
      }
    }
  }
}

// Scalar fallback (inner not a multiple of the vector width, or unaligned pointers).
template <typename T, int MODE, bool HAS_BIAS, bool HAS_REF>
__global__ void __launch_bounds__(256) bias_act_scalar_kernel(T* __restrict__ out, const T* __restrict__ x,
                                                              const T* __restrict__ bias, const T* __restrict__ ref,
                                                              uint64_t n, uint64_t inner, uint32_t channels, float alpha,
                                                              float scale) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float xf = to_f32<T>(x[i]);
    if (MODE == 40) { out[i] = from_f32<T>(xf * to_f32<T>(__ldg(bias + static_cast<uint32_t>((i / inner) % channels)))); continue; }
    if (HAS_BIAS) xf += to_f32<T>(__ldg(bias + static_cast<uint32_t>((i / inner) % channels)));
    const float rf = HAS_REF ? to_f32<T>(ref[i]) : 0.f;
    out[i] = from_f32<T>(act_apply(xf, rf, MODE, alpha) * scale);
  }
}

template <typename T, int MODE, bool HAS_BIAS, bool HAS_REF>
static int launch_bias_act(void* out, const void* x, const void* bias, const void* ref, int64_t n_outer,
                           int64_t channels, int64_t inner, float alpha, float scale, cudaStream_t st) {
  constexpr int N = Vec16<T>::N;
  const uint64_t n = static_cast<uint64_t>(n_outer) * channels * inner;
  if (n == 0) return FM_OK;
  const bool aligned = ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(x) |
                         (HAS_REF ? reinterpret_cast<uintptr_t>(ref) : 0)) & 15) == 0;
  const int sms = sm_count();
  if (aligned && inner % N == 0) {
    constexpr int UNROLL = 4;
    const uint64_t n_vec = n / N;
    const uint64_t want = (n_vec + 256ull * UNROLL - 1) / (256ull * UNROLL);
    const unsigned grid = static_cast<unsigned>(want < static_cast<uint64_t>(sms) * 16 ? (want ? want : 1) : sms * 16);
    if (n_vec < 0xFFFFFFFFull - 256ull * UNROLL * grid) {
      bias_act_vec_kernel<T, uint32_t, MODE, HAS_BIAS, HAS_REF, UNROLL><<<grid, 256, 0, st>>>(
          static_cast<T*>(out), static_cast<const T*>(x), static_cast<const T*>(bias), static_cast<const T*>(ref),
          static_cast<uint32_t>(n_vec), static_cast<uint32_t>(inner / N), static_cast<uint32_t>(channels), alpha, scale,
          fastdiv_make(static_cast<uint32_t>(inner / N)), fastdiv_make(static_cast<uint32_t>(channels ? channels : 1)));
    } else {
      bias_act_vec_kernel<T, uint64_t, MODE, HAS_BIAS, HAS_REF, UNROLL><<<grid, 256, 0, st>>>(
          static_cast<T*>(out), static_cast<const T*>(x), static_cast<const T*>(bias), static_cast<const T*>(ref), n_vec,
          static_cast<uint64_t>(inner / N), static_cast<uint32_t>(channels), alpha, scale, FastDivU32{}, FastDivU32{});
    }
  } else {
    const uint64_t want = (n + 255) / 256;
    const unsigned grid = static_cast<unsigned>(want < static_cast<uint64_t>(sms) * 32 ? want : sms * 32);
    bias_act_scalar_kernel<T, MODE, HAS_BIAS, HAS_REF><<<grid, 256, 0, st>>>(
        static_cast<T*>(out), static_cast<const T*>(x), static_cast<const T*>(bias), static_cast<const T*>(ref), n,
        static_cast<uint64_t>(inner), static_cast<uint32_t>(channels), alpha, scale);
  }
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

template <typename T, int MODE>
static int dispatch_flags(void* out, const void* x, const void* bias, const void* ref, int64_t n_outer, int64_t channels,
                          int64_t inner, float alpha, float scale, cudaStream_t st) {
  const bool hb = bias != nullptr, hr = ref != nullptr;
  if (hb && hr) return launch_bias_act<T, MODE, true, true>(out, x, bias, ref, n_outer, channels, inner, alpha, scale, st);
  if (hb) return launch_bias_act<T, MODE, true, false>(out, x, bias, ref, n_outer, channels, inner, alpha, scale, st);
  if (hr) return launch_bias_act<T, MODE, false, true>(out, x, bias, ref, n_outer, channels, inner, alpha, scale, st);
  return launch_bias_act<T, MODE, false, false>(out, x, bias, ref, n_outer, channels, inner, alpha, scale, st);
}

template <typename T>
static int dispatch_mode(int mode, void* out, const void* x, const void* bias, const void* ref, int64_t n_outer,
                         int64_t channels, int64_t inner, float alpha, float scale, cudaStream_t st) {
  switch (mode) {
    case 30: return dispatch_flags<T, 30>(out, x, bias, ref, n_outer, channels, inner, alpha, scale, st);
    case 31: return dispatch_flags<T, 31>(out, x, bias, ref, n_outer, channels, inner, alpha, scale, st);
    case 12:
    case 32: return dispatch_flags<T, 32>(out, x, bias, ref, n_outer, channels, inner, alpha, scale, st);
    default: return dispatch_flags<T, 10>(out, x, bias, ref, n_outer, channels, inner, alpha, scale, st);
  }
}

// ------------------------------------------------------------------------------------
// gradient + fused bias-gradient reduction.
// One block handles a slab of one (n, c) row; block-reduces gx in fp32 and issues a single
// atomicAdd per block into grad_bias[c]  (the reference runs a separate torch sum over the
// whole tensor, op/fused_act.py:42-48: one extra full read).
// ------------------------------------------------------------------------------------
template <typename T, int MODE, bool VEC>
__global__ void __launch_bounds__(256) bias_act_grad_bias_kernel(T* __restrict__ gin, float* __restrict__ gbias,
                                                                 const T* __restrict__ gout, const T* __restrict__ ref,
                                                                 uint64_t inner, uint32_t channels, uint32_t slabs_per_row,
                                                                 uint64_t slab, float alpha, float scale) {
  const uint64_t row = blockIdx.x / slabs_per_row;
  const uint32_t sl = blockIdx.x % slabs_per_row;
  const uint32_t c = static_cast<uint32_t>(row % channels);
  const uint64_t lo = static_cast<uint64_t>(sl) * slab;
  const uint64_t hi = lo + slab < inner ? lo + slab : inner;
  const T* g = gout + row * inner;
  const T* r = ref + row * inner;
  T* o = gin + row * inner;
  float acc = 0.f;
  if (VEC) {
    // 16-byte accesses (the host checked alignment and that rows and slabs are whole vectors): the scalar loop moved 2
    // or 4 bytes per lane and request -- 2.6 TB/s on bf16, 4.5 on fp32
    constexpr int N = Vec16<T>::N;
    const uint64_t v_hi = hi / N;
    uint64_t iv = lo / N + threadIdx.x;
    for (; iv + blockDim.x < v_hi; iv += 2 * blockDim.x) {          // two vectors of each tensor in flight
      const Pack<T, N> g0 = ld16<T>(g + iv * N), r0 = ld16<T>(r + iv * N);
      const Pack<T, N> g1 = ld16<T>(g + (iv + blockDim.x) * N), r1 = ld16<T>(r + (iv + blockDim.x) * N);
      Pack<T, N> o0, o1;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        o0.v[j] = from_f32<T>(act_apply(to_f32<T>(g0.v[j]), to_f32<T>(r0.v[j]), MODE, alpha) * scale);
        o1.v[j] = from_f32<T>(act_apply(to_f32<T>(g1.v[j]), to_f32<T>(r1.v[j]), MODE, alpha) * scale);
        acc += to_f32<T>(o0.v[j]);
        acc += to_f32<T>(o1.v[j]);
      }
      st16<T>(o + iv * N, o0);
      st16<T>(o + (iv + blockDim.x) * N, o1);
    }
    if (iv < v_hi) {
      const Pack<T, N> g0 = ld16<T>(g + iv * N), r0 = ld16<T>(r + iv * N);
      Pack<T, N> o0;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        o0.v[j] = from_f32<T>(act_apply(to_f32<T>(g0.v[j]), to_f32<T>(r0.v[j]), MODE, alpha) * scale);
        acc += to_f32<T>(o0.v[j]);
      }
      st16<T>(o + iv * N, o0);
    }
  } else
  for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const float y = act_apply(to_f32<T>(g[i]), to_f32<T>(r[i]), MODE, alpha) * scale;
    const T yq = from_f32<T>(y);
    o[i] = yq;
    acc += to_f32<T>(yq);   // the reference sums the rounded grad_input tensor
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  __shared__ float wsum[8];
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = wsum[threadIdx.x];
#pragma unroll
    for (int s = 4; s > 0; s >>= 1) v += __shfl_xor_sync(0xffu, v, s);
    if (threadIdx.x == 0) atomicAdd(gbias + c, v);
  }
}

template <typename T>
static int launch_grad_bias(void* gin, float* gbias, const void* gout, const void* ref, int64_t n_outer, int64_t channels,
                            int64_t inner, int act, float alpha, float scale, cudaStream_t st) {
  const uint64_t rows = static_cast<uint64_t>(n_outer) * channels;
  if (rows == 0 || inner == 0) return FM_OK;
  // Aim for >= 4 waves of blocks; slab >= 2048 elements when the row is long.
  uint64_t slab = static_cast<uint64_t>(inner);
  const uint64_t target_blocks = static_cast<uint64_t>(sm_count()) * 32;
  while (slab > 2048 && rows * ((inner + slab - 1) / slab) < target_blocks) slab = (slab + 1) / 2;
  constexpr int N = Vec16<T>::N;
  const bool vec = inner % N == 0 && ((reinterpret_cast<uintptr_t>(gin) | reinterpret_cast<uintptr_t>(gout) |
                                       reinterpret_cast<uintptr_t>(ref)) & 15) == 0;
  if (vec) slab = (slab + N - 1) / N * N;          // slabs of whole vectors (rows are: inner % N == 0)
  const uint64_t slabs = (inner + slab - 1) / slab;
  const uint64_t blocks = rows * slabs;
  FM_CHECK_ARG(blocks < 0x7FFFFFFFull, "bias_act_grad_bias: too many blocks (%llu)", (unsigned long long)blocks);
#define FM_GB(MODE_, VEC_) bias_act_grad_bias_kernel<T, MODE_, VEC_><<<static_cast<unsigned>(blocks), 256, 0, st>>>( \
        static_cast<T*>(gin), gbias, static_cast<const T*>(gout), static_cast<const T*>(ref), static_cast<uint64_t>(inner), \
        static_cast<uint32_t>(channels), static_cast<uint32_t>(slabs), slab, alpha, scale)
  if (act == 3) { if (vec) FM_GB(31, true); else FM_GB(31, false); }
  else          { if (vec) FM_GB(11, true); else FM_GB(11, false); }
#undef FM_GB
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

// ------------------------------------------------------------------------------------
// per-row dot product: out[r] += sum_i a[r, i] * b[r, i].  Same decomposition as the fused bias gradient: one block per slab
// of a row, fp32 block reduction, one atomicAdd per block.
// ------------------------------------------------------------------------------------
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) channel_dot_kernel(float* __restrict__ out, const T* __restrict__ a, const T* __restrict__ b,
                                                          uint64_t inner, uint32_t slabs_per_row, uint64_t slab) {
  const uint64_t row = blockIdx.x / slabs_per_row;
  const uint32_t sl = blockIdx.x % slabs_per_row;
  const uint64_t lo = static_cast<uint64_t>(sl) * slab;
  const uint64_t hi = lo + slab < inner ? lo + slab : inner;
  const T* pa = a + row * inner;
  const T* pb = b + row * inner;
  float acc = 0.f;
  if (VEC) {
    constexpr int N = Vec16<T>::N;
    const uint64_t v_hi = hi / N;
    uint64_t iv = lo / N + threadIdx.x;
    for (; iv + blockDim.x < v_hi; iv += 2 * blockDim.x) {          // two vectors of each tensor in flight
      const Pack<T, N> a0 = ld16<T>(pa + iv * N), b0 = ld16<T>(pb + iv * N);
      const Pack<T, N> a1 = ld16<T>(pa + (iv + blockDim.x) * N), b1 = ld16<T>(pb + (iv + blockDim.x) * N);
#pragma unroll
      for (int j = 0; j < N; ++j) acc = fmaf(to_f32<T>(a0.v[j]), to_f32<T>(b0.v[j]), fmaf(to_f32<T>(a1.v[j]), to_f32<T>(b1.v[j]), acc));
    }
    if (iv < v_hi) {
      const Pack<T, N> a0 = ld16<T>(pa + iv * N), b0 = ld16<T>(pb + iv * N);
#pragma unroll
      for (int j = 0; j < N; ++j) acc = fmaf(to_f32<T>(a0.v[j]), to_f32<T>(b0.v[j]), acc);
    }
  } else {
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) acc = fmaf(to_f32<T>(pa[i]), to_f32<T>(pb[i]), acc);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  __shared__ float wsum[8];
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = wsum[threadIdx.x];
#pragma unroll
    for (int s = 4; s > 0; s >>= 1) v += __shfl_xor_sync(0xffu, v, s);
    if (threadIdx.x == 0) atomicAdd(out + row, v);
  }
}

template <typename T>
static int launch_channel_dot(float* out, const void* a, const void* b, int64_t rows, int64_t inner, cudaStream_t st) {
  if (rows == 0 || inner == 0) return FM_OK;
  constexpr int N = Vec16<T>::N;
  uint64_t slab = static_cast<uint64_t>(inner);
  const uint64_t target_blocks = static_cast<uint64_t>(sm_count()) * 32;
  while (slab > 2048 && static_cast<uint64_t>(rows) * ((inner + slab - 1) / slab) < target_blocks) slab = (slab + 1) / 2;
  const bool vec = inner % N == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
  if (vec) slab = (slab + N - 1) / N * N;
  const uint64_t slabs = (inner + slab - 1) / slab;
  const uint64_t blocks = static_cast<uint64_t>(rows) * slabs;
  FM_CHECK_ARG(blocks < 0x7FFFFFFFull, "fm_channel_dot: too many blocks (%llu)", (unsigned long long)blocks);
  if (vec)
    channel_dot_kernel<T, true><<<static_cast<unsigned>(blocks), 256, 0, st>>>(out, static_cast<const T*>(a), static_cast<const T*>(b),
                                                                               static_cast<uint64_t>(inner), static_cast<uint32_t>(slabs), slab);
  else
    channel_dot_kernel<T, false><<<static_cast<unsigned>(blocks), 256, 0, st>>>(out, static_cast<const T*>(a), static_cast<const T*>(b),
                                                                                static_cast<uint64_t>(inner), static_cast<uint32_t>(slabs), slab);
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

// ------------------------------------------------------------------------------------
// out[b, c, i] = x[b, c, i] + n[b * n_bstride + i]: a warp owns chunks of 128 consecutive 16-byte vectors as above; the plane
// is read through L1 / L2 (it is reused by every channel).
// ------------------------------------------------------------------------------------
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) plane_add_kernel(T* __restrict__ out, const T* __restrict__ x, const T* __restrict__ n,
                                                        uint64_t total, uint64_t inner, uint64_t chan_inner, uint64_t n_bstride) {
  constexpr int N = VEC ? Vec16<T>::N : 1;
  constexpr int UNROLL = 4;
  const uint64_t n_vec = total / N;
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (uint64_t v0 = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v0 < n_vec; v0 += stride * UNROLL) {
    Pack<T, N> xv[UNROLL], nv[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const uint64_t v = v0 + stride * u;
      if (v < n_vec) {
        const uint64_t e = v * N, b = e / chan_inner, i = e % inner;       // (sample, offset inside the plane)
        if constexpr (VEC) { xv[u] = ld16<T>(x + e); nv[u] = ld16<T>(n + b * n_bstride + i); }
        else { xv[u].v[0] = x[e]; nv[u].v[0] = __ldg(n + b * n_bstride + i); }
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const uint64_t v = v0 + stride * u;
      if (v < n_vec) {
        Pack<T, N> o;
#pragma unroll
        for (int j = 0; j < N; ++j) o.v[j] = from_f32<T>(to_f32<T>(xv[u].v[j]) + to_f32<T>(nv[u].v[j]));
        if constexpr (VEC) st16<T>(out + v * N, o); else out[v] = o.v[0];
      }
    }
  }
}

template <typename T>
static int launch_plane_add(void* out, const void* x, const void* n, int64_t B, int64_t C, int64_t inner, int64_t n_bstride,
                            cudaStream_t st) {
  const uint64_t total = static_cast<uint64_t>(B) * C * inner;
  if (total == 0) return FM_OK;
  constexpr int N = Vec16<T>::N;
  const bool vec = inner % N == 0 && n_bstride % N == 0 &&
                   ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(n)) & 15) == 0;
  const uint64_t n_vec = vec ? total / N : total;
  const uint64_t want = (n_vec + 1023) / 1024;
  const uint64_t cap = static_cast<uint64_t>(sm_count()) * 16;
  const unsigned grid = static_cast<unsigned>(want < cap ? (want ? want : 1) : cap);
  if (vec)
    plane_add_kernel<T, true><<<grid, 256, 0, st>>>(static_cast<T*>(out), static_cast<const T*>(x), static_cast<const T*>(n), total,
                                                    static_cast<uint64_t>(inner), static_cast<uint64_t>(C) * inner,
                                                    static_cast<uint64_t>(n_bstride));
  else
    plane_add_kernel<T, false><<<grid, 256, 0, st>>>(static_cast<T*>(out), static_cast<const T*>(x), static_cast<const T*>(n), total,
                                                     static_cast<uint64_t>(inner), static_cast<uint64_t>(C) * inner,
                                                     static_cast<uint64_t>(n_bstride));
  count_launch();
  FM_LAUNCH_OK();
  return FM_OK;
}

}  // namespace fm

extern "C" int fm_plane_add(void* out, const void* x, const void* n, int64_t B, int64_t C, int64_t inner, int64_t n_bstride,
                            int dtype, void* stream) {
  FM_CHECK_ARG(B >= 0 && C >= 0 && inner >= 0 && (n_bstride == 0 || n_bstride == inner), "fm_plane_add: bad sizes");
  FM_CHECK_ARG(out && x && n, "fm_plane_add: null tensor");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case FM_F32: return fm::launch_plane_add<float>(out, x, n, B, C, inner, n_bstride, st);
    case FM_F16: return fm::launch_plane_add<__half>(out, x, n, B, C, inner, n_bstride, st);
    case FM_BF16: return fm::launch_plane_add<__nv_bfloat16>(out, x, n, B, C, inner, n_bstride, st);
    default: fm::set_error("fm_plane_add: bad dtype %d", dtype); return FM_ERR_INVALID;
  }
}

extern "C" int fm_channel_scale(void* out, const void* x, const void* s, int64_t rows, int64_t inner, int dtype, void* stream) {
  FM_CHECK_ARG(rows >= 0 && inner >= 0, "fm_channel_scale: negative size");
  FM_CHECK_ARG(out && x && s, "fm_channel_scale: null tensor");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (dtype) {     // bias_act_vec_kernel in mode 40: the per-channel "bias" is the per-row factor (n_outer = 1, channels = rows)
    case FM_F32: return fm::launch_bias_act<float, 40, true, false>(out, x, s, nullptr, 1, rows, inner, 0.f, 1.f, st);
    case FM_F16: return fm::launch_bias_act<__half, 40, true, false>(out, x, s, nullptr, 1, rows, inner, 0.f, 1.f, st);
    case FM_BF16: return fm::launch_bias_act<__nv_bfloat16, 40, true, false>(out, x, s, nullptr, 1, rows, inner, 0.f, 1.f, st);
    default: fm::set_error("fm_channel_scale: bad dtype %d", dtype); return FM_ERR_INVALID;
  }
}

extern "C" int fm_channel_dot(float* out_f32, const void* a, const void* b, int64_t rows, int64_t inner, int dtype, void* stream) {
  FM_CHECK_ARG(rows >= 0 && inner >= 0, "fm_channel_dot: negative size");
  FM_CHECK_ARG(out_f32 && a && b, "fm_channel_dot: null tensor");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case FM_F32: return fm::launch_channel_dot<float>(out_f32, a, b, rows, inner, st);
    case FM_F16: return fm::launch_channel_dot<__half>(out_f32, a, b, rows, inner, st);
    case FM_BF16: return fm::launch_channel_dot<__nv_bfloat16>(out_f32, a, b, rows, inner, st);
    default: fm::set_error("fm_channel_dot: bad dtype %d", dtype); return FM_ERR_INVALID;
  }
}

extern "C" int fm_bias_act(void* out, const void* x, const void* bias, const void* ref, int64_t n_outer, int64_t channels,
                           int64_t inner, int act, int grad, float alpha, float scale, int dtype, void* stream) {
  FM_CHECK_ARG(n_outer >= 0 && channels >= 0 && inner >= 0, "fm_bias_act: negative size");
  FM_CHECK_ARG(out && x, "fm_bias_act: null tensor");
  FM_CHECK_ARG(act == 1 || act == 3, "fm_bias_act: act must be 1 (linear) or 3 (lrelu), got %d", act);
  FM_CHECK_ARG(grad >= 0 && grad <= 2, "fm_bias_act: grad must be 0..2, got %d", grad);
  FM_CHECK_ARG(grad != 1 || act != 3 || ref != nullptr, "fm_bias_act: grad=1 with lrelu needs ref");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int mode = act * 10 + grad;
  if (mode != 31) ref = nullptr;  // ref is only read by mode 31 (.cu:42)
  switch (dtype) {
    case FM_F32: return fm::dispatch_mode<float>(mode, out, x, bias, ref, n_outer, channels, inner, alpha, scale, st);
    case FM_F16: return fm::dispatch_mode<__half>(mode, out, x, bias, ref, n_outer, channels, inner, alpha, scale, st);
    case FM_BF16: return fm::dispatch_mode<__nv_bfloat16>(mode, out, x, bias, ref, n_outer, channels, inner, alpha, scale, st);
    default: fm::set_error("fm_bias_act: bad dtype %d", dtype); return FM_ERR_INVALID;
  }
}

extern "C" int fm_bias_act_grad_bias(void* grad_in, float* grad_bias_f32, const void* grad_out, const void* ref,
                                     int64_t n_outer, int64_t channels, int64_t inner, int act, float alpha, float scale,
                                     int dtype, void* stream) {
  FM_CHECK_ARG(n_outer >= 0 && channels >= 0 && inner >= 0, "fm_bias_act_grad_bias: negative size");
  FM_CHECK_ARG(grad_in && grad_bias_f32 && grad_out && ref, "fm_bias_act_grad_bias: null tensor");
  FM_CHECK_ARG(act == 1 || act == 3, "fm_bias_act_grad_bias: act must be 1 or 3, got %d", act);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case FM_F32: return fm::launch_grad_bias<float>(grad_in, grad_bias_f32, grad_out, ref, n_outer, channels, inner, act, alpha, scale, st);
    case FM_F16: return fm::launch_grad_bias<__half>(grad_in, grad_bias_f32, grad_out, ref, n_outer, channels, inner, act, alpha, scale, st);
    case FM_BF16: return fm::launch_grad_bias<__nv_bfloat16>(grad_in, grad_bias_f32, grad_out, ref, n_outer, channels, inner, act, alpha, scale, st);
    default: fm::set_error("fm_bias_act_grad_bias: bad dtype %d", dtype); return FM_ERR_INVALID;
  }
}
