/*
 * fm3d.h -- C ABI of libfm3d.so: the B200 (sm_100a) kernels behind 3D-FM GAN's
 * StyleGAN2 synthesis hot path.
 *
 * Conventions (all entry points):
 *   - plain C: raw device pointers, sizes, a cudaStream_t passed as void*; no torch types;
 *   - return 0 on success, non-zero fm_status on error; fm_last_error() gives the text
 *     (thread-local).  No exceptions cross the boundary;
 *   - the callee never allocates device memory, never synchronises, and enqueues all
 *     work on the stream it is given (the reference ops use the current torch stream,
 *     op/fused_bias_act_kernel.cu:54-56, op/upfirdn2d_kernel.cu:213-215);
 *   - the caller owns every buffer; outputs must be pre-allocated and contiguous;
 *   - re-entrant and thread-safe (the reference is called from nn.DataParallel worker
 *     threads, train_3_encoder.py:355-362).
 *
 * Each declaration cites the reference interface it replaces (file:line in
 * adobe/3D-FM-GAN).
 */
#ifndef FM3D_H_
#define FM3D_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { FM_F32 = 0, FM_F16 = 1, FM_BF16 = 2 } fm_dtype;

typedef enum {
  FM_OK = 0,
  FM_ERR_INVALID = 1,      /* bad argument (shape, alignment, enum) */
  FM_ERR_UNSUPPORTED = 2,  /* valid but outside what the kernels cover */
  FM_ERR_CUDA = 3,         /* a CUDA runtime / driver call failed */
  FM_ERR_NO_DEVICE = 4     /* no sm_100 device / driver entry point missing */
} fm_status;

/* Library identification. */
int fm_version(void);
const char* fm_last_error(void);
/* Number of kernel launches issued by this library since load (all threads). */
int64_t fm_launch_count(void);
/* Credit n launches replayed by a captured CUDA graph (they bypass the launch wrappers). */
void fm_add_launches(int64_t n);

/* ------------------------------------------------------------------------------------
 * fused bias + activation.
 * Replaces: fused_bias_act(input, bias, refer, act, grad, alpha, scale) -> Tensor
 *           op/fused_bias_act.cpp:11-17, kernel op/fused_bias_act_kernel.cu:18-49.
 * x is viewed as [n_outer, channels, inner] (bias broadcasts along dim 1; inner =
 * prod(dims[2:]), .cu:67-71).  bias / ref may be NULL (= the reference's empty tensor,
 * .cu:62-63).  act: 1 linear, 3 leaky-ReLU; grad: 0 forward, 1 first derivative w.r.t.
 * ref, 2 second derivative (zero).  Indexing is 64-bit (the reference overflows at 2^31
 * elements, .cu:65).  bias has the same dtype as x.
 * ---------------------------------------------------------------------------------- */
int fm_bias_act(void* out, const void* x, const void* bias, const void* ref,
                int64_t n_outer, int64_t channels, int64_t inner,
                int act, int grad, float alpha, float scale, int dtype, void* stream);

/* Gradient mode (act, grad=1, no bias) fused with the bias-gradient reduction that the
 * reference runs as a separate torch sum (op/fused_act.py:42-48).
 * grad_bias_f32[channels] must be zero-initialised by the caller; it is accumulated in fp32. */
int fm_bias_act_grad_bias(void* grad_in, float* grad_bias_f32, const void* grad_out,
                          const void* ref, int64_t n_outer, int64_t channels, int64_t inner,
                          int act, float alpha, float scale, int dtype, void* stream);

/* Per-row scale and per-row dot product: the two maps of the modulation / demodulation scaling of ModulatedConv2d's
 * shared-weight composition (stylegan2.py:250-298: weight * style, demod), closed under differentiation --
 *   fm_channel_scale: out[r, i] = x[r, i] * s[r]                 (x * s[:, :, None, None] with r = b * C + c)
 *   fm_channel_dot:   out_f32[r] += sum_i a[r, i] * b[r, i]      (its gradient w.r.t. s; out_f32 zero-initialised by the caller)
 * The gradient of each is built from the other.  s has the dtype of x. */
int fm_channel_scale(void* out, const void* x, const void* s, int64_t rows, int64_t inner, int dtype, void* stream);
int fm_channel_dot(float* out_f32, const void* a, const void* b, int64_t rows, int64_t inner, int dtype, void* stream);

/* out[b, c, i] = x[b, c, i] + n[b * n_bstride + i]  (n_bstride = inner for a per-sample plane, 0 for one shared plane):
 * NoiseInjection's broadcast add over the channels (stylegan2.py:312, image + weight * noise) under autograd. */
int fm_plane_add(void* out, const void* x, const void* n, int64_t B, int64_t C, int64_t inner, int64_t n_bstride, int dtype,
                 void* stream);

/* ------------------------------------------------------------------------------------
 * upfirdn2d: zero-stuff x up, pad/crop, correlate with the flipped kernel, decimate.
 * Replaces: upfirdn2d(input[major,H,W,minor], kernel[kh,kw], up_x, up_y, down_x, down_y,
 *                     pad_x0, pad_x1, pad_y0, pad_y1) -> Tensor[major,out_h,out_w,minor]
 *           op/upfirdn2d.cpp:12-19, kernels op/upfirdn2d_kernel.cu:49-105,107-207.
 * The Python wrapper always passes minor = 1 and major = N*C (op/upfirdn2d.py:108), so
 * the ABI takes `planes` = N*C contiguous [in_h, in_w] images.  The kernel taps are fp32
 * (host or device pointer is NOT accepted: device pointer, kh*kw floats).
 * out must hold planes*out_h*out_w elements with
 *   out_h = (in_h*up_y + pad_y0 + pad_y1 - kh) / down_y + 1   (op/upfirdn2d_kernel.cu:237-240).
 * ---------------------------------------------------------------------------------- */
int fm_upfirdn2d(void* out, const void* x, const float* kernel,
                 int64_t planes, int in_h, int in_w, int kh, int kw,
                 int up_x, int up_y, int down_x, int down_y,
                 int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                 int dtype, void* stream);

/* ------------------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05 tensor cores (bf16 in, fp32 accumulate in TMEM).
 * Replaces the reference's ATen calls F.conv2d / F.conv_transpose2d(groups=batch) on the
 * materialised per-sample weights (stylegan2.py:276,285,291), the EqualConv2d call
 * (stylegan2.py:129) and the encoders' nn.Conv2d (resnet_encoder.py:36,42,193;
 * psp_encoder_model/encoders/helpers.py:80-82,124-131; psp_encoders.py:27-31).
 *
 * Data layout: activations NHWC bf16 with a physical channel stride (multiple of 8);
 * weights bf16 [ntaps][cout_rows][cin_stride] (K-major).  Style modulation is applied to
 * the activations by the producer (x * s[b,i]) and demodulation d[b,o] in the epilogue, so
 * the weight operand is shared by the whole batch and B*H*W is the GEMM M dimension
 * (SURVEY.md 7.3 algebra; identical to stylegan2.py:257-262 up to rounding order).
 *
 * Epilogue per output element (b, y, x, o), with table row T = tab[(b*tab_bstride + o)*8]:
 *   v = acc * T[0] + T[1] + noise[b*noise_bstride + y*OWfull + x] * (*noise_w)  [+ residual]
 *   v = v > 0 ? v : v * T[2]
 *   rgb[b,y,x,0..2] += v * T[4..6]            (optional fused ToRGB, stylegan2.py:393)
 *   out[b,y,x,o]     = v * T[3]
 * ---------------------------------------------------------------------------------- */
#define FM_MAX_TAPS 49

typedef struct {
  /* input activations */
  const void* x;            /* bf16 NHWC */
  int32_t B, H, W, Cin;     /* logical sizes */
  int32_t x_cstride;        /* physical channel stride in elements (multiple of 8) */
  /* weights */
  const void* w;            /* bf16 [ntaps][w_rows][w_cstride] */
  int32_t ntaps, Cout, w_rows, w_cstride;
  int8_t tap_dy[FM_MAX_TAPS];   /* input row offset of each tap relative to oy*stride */
  int8_t tap_dx[FM_MAX_TAPS];
  int8_t tap_widx[FM_MAX_TAPS]; /* which [w_rows][w_cstride] slab of w each tap multiplies */
  int32_t stride;           /* convolution stride (1 or 2), both axes unless stride_x/stride_y are set */
  int32_t stride_x, stride_y;   /* 0 = use `stride` */
  /* optional custom input strides in elements (0 = dense NHWC): lets a "pixel" be an overlapping
   * window of a padded row (3-channel stems: 8 px x 8 ch = one 64-wide K chunk) */
  int64_t x_pixstride, x_rowstride, x_imgstride;
  /* groups: x holds groups*Bg images (group-major), w holds groups*[slabs][w_rows][w_cstride],
   * a shared table holds groups*[Cout][8].  0/1 = ungrouped. */
  int32_t groups;
  /* logical output grid computed by this launch */
  int32_t OH, OW;
  /* placement of that grid inside the output tensor: pixel (oy,ox) is written to
   * (oy*out_ys + out_y0, ox*out_xs + out_x0) of a [B,out_H,out_W,out_cstride] tensor */
  void* out;                /* bf16 NHWC, or fp32 NCHW when out_nchw_f32 != 0; NULL (with rgb != NULL) = only the
                             * fused ToRGB sums are produced: the last synthesis layer's activations feed nothing else */
  int32_t out_H, out_W, out_cstride, out_y0, out_x0, out_ys, out_xs;
  int32_t out_nchw_f32;
  /* epilogue */
  const float* tab;         /* [tab_rows][Cout][8] fp32 (see above); tab_bstride = Cout or 0.  NULL = identity epilogue
                             * (out = acc): the parity phases of the stride-2 transposed conv, whose epilogue runs in
                             * the blur pass */
  int32_t tab_bstride;      /* 0: one table shared by the batch; 1: per-sample tables */
  const float* noise;       /* fp32 [B or 1][out_H][out_W] or NULL */
  int32_t noise_bstride;    /* 0 shared / 1 per-sample */
  const float* noise_w;     /* device scalar or NULL (=> 1.0) */
  const void* residual;     /* bf16 NHWC like out, or NULL */
  float* rgb;               /* fp32 [B][out_H][out_W][4] accumulated with atomics, or NULL */
  /* 3x3/pad-1 convs fed by a folded input BatchNorm: additive correction per border class
   * ((y==0?1:y==OH-1?2:0)*3 + (x==0?1:x==OW-1?2:0)), fp32 [9][Cout], or NULL */
  const float* border_tab;
  /* concatenated-N output: channels [g*out_cgroup,(g+1)*out_cgroup) go to out + g*out_gstride
   * elements (a [G*B,H,W,out_cstride] group-major tensor); 0 = off */
  int32_t out_cgroup;
  int64_t out_gstride;
  /* split-K workspace: fp32 zeros, at least B*OH*OW*roundup(Cout,16)*4 bytes when used.  Small-M
   * layers split the K loop over several CTAs (fp32 atomics into the workspace) and a second
   * tiny kernel applies the epilogue and re-zeroes it.  NULL disables split-K. */
  float* splitk_ws;
  int64_t splitk_ws_bytes;
  int32_t ksplit;           /* 0 = choose automatically, 1 = off, k > 1 = force */
  /* upmode != 0: fused stride-2 transposed 3x3 conv (stylegan2.py:276).  w holds the 9 slabs of the
   * 3x3 kernel; OH = H+1, OW = W+1 is the grid of output 2x2 blocks; the four output parities are
   * accumulated from four shared shifted input views and written to out[b, 2y+py, 2x+px, :]
   * (out_H = 2H+1, out_W = 2W+1, out_ys = out_xs = 2).  taps are ignored. */
  int32_t upmode;
  /* tiling hints (0 = choose) */
  int32_t block_n;          /* 64, 128 or 256 */
  int32_t tile_w, tile_h;   /* tile_w*tile_h*tile_b = 128 output pixels */
  /* concatenated-N output: group g only writes columns ox < OW - g*out_cgroup_ow_shrink (the odd-column
   * parity of a stride-2 transposed conv has one column less than the even one) */
  int32_t out_cgroup_ow_shrink;
  /* residual_up_h/w > 0: `residual` is a LOW-resolution bf16 NHWC tensor [B, residual_up_h, residual_up_w, out_cstride]
   * that the epilogue samples bilinearly with align_corners=True (the FPN top-down add of the pSp encoder,
   * psp_encoders.py:81-98, without materialising the upsampled map) */
  int32_t residual_up_h, residual_up_w;
  /* physical pitch of the bf16 NHWC output tensor when it is larger than its logical size: rows per image and pixels
   * per row (0 = out_H / out_W).  noise, residual and rgb stay indexed on the logical [out_H][out_W] grid.  The
   * synthesis engine stores every activation that feeds a stride-2 transposed conv with one zero row and column of
   * padding, so that the batch is ONE tall image whose separator rows are the conv's zero padding. */
  int32_t out_pitch_h, out_pitch_w;
  /* nphases > 1: ONE launch computes several output phases of a stride-2 transposed conv (stylegan2.py:276) that
   * share the input tensor and the tile grid: the tap list is the concatenation of the phases' tap lists
   * (phase_ntaps[i] taps each, 4/2/2/1 for the 3x3 kernel), phase i is written to
   * (oy*out_ys + phase_out_y0[i], ox*out_xs + phase_out_x0[i]); out_y0 / out_x0 are ignored.  The phase is a tile
   * coordinate of the persistent grid, so the weight-bound 1- and 2-tap tiles run next to the tensor-bound 4-tap tiles
   * instead of in launches of their own.  Stride 1, ungrouped, halo-patch eligible shapes only (OH >= 12, OW >= 8). */
  int32_t nphases;
  int32_t phase_ntaps[4], phase_out_y0[4], phase_out_x0[4];
  /* max_ctas > 0: the persistent grid takes at most this many CTAs (= SMs: one CTA owns an SM) instead of every SM of
   * the device.  A latency-bound network of small layers (the two ResNet-18 encoders) confined to a few SMs runs next
   * to the large layers of the other networks instead of time-slicing the whole chip with them; such launches do not
   * use programmatic dependent launch (the next kernel's CTAs would wait on SMs outside the partition). */
  int32_t max_ctas;
  /* tile_counter != NULL: dynamic tile schedule.  Two int32 in device memory, zero before the first launch; the CTAs
   * draw their tiles from [0] instead of taking them round robin, and the last cluster to finish re-zeroes both, so
   * launches that are stream-ordered (or programmatically dependent) may share one pair.  Launches that can run
   * concurrently must not.  A CTA that got its SM late -- another stream's kernel still held it -- then takes fewer
   * tiles instead of delaying the grid. */
  int32_t* tile_counter;
  /* colsum != NULL (plain epilogue only: tab, no rgb / residual / border_tab / split-K / groups, bf16 NHWC output, tiles
   * inside one image): colsum[b][o] += sum over the pixels of out[b,:,:,o] (the values before bf16 rounding), fp32
   * [B][Cout], accumulated with atomics -- the squeeze (global average pool) of the SE block that follows the second
   * conv of a bottleneck_IR_SE unit (psp_encoder_model/encoders/helpers.py), without a pass over the output. */
  float* colsum;
} fm_conv_desc;

int fm_conv_igemm(const fm_conv_desc* desc, void* stream);
/* Debug aid: with FM3D_TRACE=1 every igemm launch records per-CTA event clocks (kernel start, per tile: loads
 * issued / accumulator free / first operand landed / accumulator complete); this copies the last launch's
 * [n_ctas][*slots_out] int64 table to the host (synchronises the device). */
int fm_igemm_trace(long long* host_out, int n_ctas, int* slots_out);

/* ------------------------------------------------------------------------------------
 * Synthesis-path helper kernels (all bandwidth-bound).
 * ---------------------------------------------------------------------------------- */

/* s[b,i] = latent[b, latent_idx, :] . wmod[i,:] / sqrt(D) + bmod[i]     (EqualLinear with
 * bias_init=1, stylegan2.py:165-175,240,257) for a batch of layers in one launch. */
typedef struct {
  const float* wmod;   /* [cin][style_dim] */
  const float* bmod;   /* [cin] */
  float* s;            /* out [B][cin] */
  int32_t cin, latent_idx;
} fm_style_layer;
int fm_style_affine(const fm_style_layer* layers_dev, int n_layers, int max_cin,
                    const float* latent, int B, int n_latent, int style_dim, void* stream);

/* Epilogue tables for one modulated conv (stylegan2.py:258-262,312,371,393-394):
 *   T[0] = demod ? rsqrt(sum_i s[b,i]^2 wsq[o,i] + 1e-8) : 1
 *   T[1] = act_bias[o] (or 0), T[2] = slope, T[3] = gain * (s_next ? s_next[b,o] : 1)
 *   T[4..6] = gain * wrgb[j,o] * s_rgb[b,o] / sqrt(cout)   (when wrgb != NULL) */
typedef struct {
  const float* s;        /* [B][cin] this layer's style */
  const float* wsq;      /* [cout][cin] sum_k (scale*W)^2, or NULL (no demod) */
  const float* act_bias; /* [cout] or NULL */
  const float* s_next;   /* [B][cout] style of the 3x3 consumer or NULL */
  const float* wrgb;     /* [3][cout] ToRGB weight or NULL */
  const float* s_rgb;    /* [B][cout] */
  float* tab;            /* out [B][cout][8] */
  int32_t cin, cout;
  float slope, gain;
} fm_table_layer;
int fm_build_tables(const fm_table_layer* layers_dev, int n_layers, int max_cout, int max_cin, int B,
                    void* stream);

/* NCHW fp32 -> NHWC bf16 with optional per-(b,c) scale; pad channels are zero-filled. */
int fm_nchw_to_nhwc_bf16(void* out, const float* x, const float* scale_bc,
                         int B, int C, int H, int W, int out_cstride, void* stream);
/* NHWC bf16 -> NCHW fp32 with optional per-(b,c) inverse scale. */
int fm_nhwc_bf16_to_nchw(float* out, const void* x, const float* inv_scale_bc,
                         int B, int C, int H, int W, int x_cstride, void* stream);

/* Blur after the stride-2 transposed modulated conv, fused with demod, noise, bias,
 * leaky-ReLU and the next layer's style (stylegan2.py:279,312,371):
 *   t[b,2h+1,2w+1,C] bf16 NHWC  ->  out[b,2h,2w,C] bf16 NHWC
 *   v = (sum_{a,b} k[a,b] t[..]) * T[0] + T[1] + noise*noise_w ; lrelu(T[2]) ; * T[3]
 * separable != 0 asserts that kernel4x4 is rank-1 (outer product): half the FMAs. */
int fm_blur_act_nhwc(void* out, const void* t, const float* kernel4x4,
                     const float* tab, const float* noise, int noise_bstride, const float* noise_w,
                     int B, int OH, int OW, int C, int cstride, int separable,
                     int t_pitch_h, int t_pitch_w, void* stream);
/* t_pitch_h / t_pitch_w: physical rows per image / pixels per row of t when larger than OH+1 / OW+1 (0 = dense). */

/* ToRGB tail (stylegan2.py:394-399): rgb_out[b,:,y,x] = acc[b,y,x,:] + bias + up2(skip).
 * acc fp32 [B][H][W][4] (from the fused epilogue), skip fp32 NCHW [B,3,H/2,W/2] or NULL,
 * kernel4x4 = Upsample.kernel (k*4).  Output fp32 NCHW [B,3,H,W].  acc is re-zeroed. */
int fm_rgb_finalize(float* rgb_out, float* acc, const float* bias3, const float* skip,
                    const float* kernel4x4, int B, int H, int W, void* stream);

/* Weight preparation (derived caches; never state):
 *   wq [kh*kw][cout_rows][cin_stride] bf16 = scale * W[o,i,ky,kx] (slab index ky*kw+kx),
 *   wsq[cout][cin] fp32 = sum_taps (scale*W)^2. */
int fm_prep_weight(void* wq, float* wsq, const float* w_oikk, int cout, int cin, int kh, int kw,
                   float scale, int cout_rows, int cin_stride, void* stream);

/* ------------------------------------------------------------------------------------
 * Weight gradient of a convolution (training, BASELINE config 4).
 * Replaces the wgrad half of ATen autograd over F.conv2d / F.conv_transpose2d
 * (stylegan2.py:129,276,285,291; cuDNN in the reference).  dL/dW is one GEMM over pixels on the
 * tcgen05 tensor cores, reading both operands as NHWC bf16 (the layout of fm_conv_igemm):
 *
 *   dw[t*dw_tap_stride + ca*dw_row_stride + cb] +=
 *       sum_{n < B, gy < GH, gx < GW}  a[n, gy*a.stride + tap_dy_a[t], gx*a.stride + tap_dx_a[t], ca]
 *                                    * b[n, gy*b.stride + tap_dy_b[t], gx*b.stride + tap_dx_b[t], cb]
 *
 * (fp32 atomics: dw must be zero-initialised by the caller; pixels outside an operand read as 0 --
 * the conv's zero padding).  One operand must BE the grid (H = GH, W = GW, stride 1, zero tap offsets): its
 * out-of-bounds zero fill masks the pixels of the last 64-pixel K chunk that lie beyond the grid.  For y = conv(x, w, stride s, padding p): the grid is y's, a = dL/dy
 * (stride 1, no shift), b = x (stride s, shift (ky - p, kx - p)) or the two swapped (the operand with
 * fewer channels should be b).  dL/dW of the modulated conv in its shared-weight form (SURVEY
 * Appendix D: one wgrad GEMM over M = B*HW on d*g and s*x) is this call on the scaled tensors.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  const void* ptr;          /* bf16 NHWC [B][H][W][cstride] */
  int32_t C, cstride;       /* logical channels, physical channel stride (multiple of 8) */
  int32_t H, W;
  int32_t stride;           /* 1 or 2: step of this operand per grid pixel */
} fm_wgrad_operand;

typedef struct {
  fm_wgrad_operand a, b;
  int32_t B, GH, GW;        /* contraction grid */
  int32_t ntaps;
  int8_t tap_dy_a[FM_MAX_TAPS], tap_dx_a[FM_MAX_TAPS], tap_dy_b[FM_MAX_TAPS], tap_dx_b[FM_MAX_TAPS];
  float* dw;                /* fp32, accumulated into */
  int64_t dw_tap_stride;
  int32_t dw_row_stride;
  int32_t ksplit;           /* 0 = choose (split-K over pixels so that every SM has work) */
} fm_wgrad_desc;
int fm_conv_wgrad(const fm_wgrad_desc* desc, void* stream);

/* ------------------------------------------------------------------------------------
 * Encoder-path helpers (NHWC bf16, bandwidth-bound).  They replace the ATen kernels behind
 * resnet_encoder.py:258-280 (stem, MaxPool2d, AvgPool2d / AdaptiveAvgPool2d),
 * psp_encoder_model/encoders/helpers.py:76-139 (SEModule, residual add, MaxPool2d(1,s)
 * shortcut) and psp_encoders.py:81-98 (bilinear FPN upsample).
 * ---------------------------------------------------------------------------------- */
/* fp32 NCHW image with C <= 8 channels -> zero-padded bf16 [B,Hp,Wp,8] (image at (pad_t,pad_l)). */
int fm_image_to_nhwc8_padded(void* out, const float* x, int B, int C, int H, int W,
                             int pad_t, int pad_l, int Hp, int Wp, void* stream);
/* tensor2im of the whole batch (replaces Evaluation/visual_eval.py:24-38, which converts one image at a time on
 * the host): img fp32 NCHW [B,3,H,W] -> out uint8 NHWC [B,H,W,3] = trunc((clip(img,-1,1) + cent) * factor). */
int fm_tensor2im_u8(void* out_u8, const float* img, int B, int H, int W, float cent, float factor, void* stream);
/* Input stage, the inverse of fm_tensor2im_u8 (replaces torchvision ToTensor() + Normalize(mean, std) of the reference's
 * transform, train_3_encoder.py:231-237, run per image on the loader workers, and the fp32 H2D copy that follows it,
 * dataset.py:401): img uint8 NHWC [B,H,W,3] -> out fp32 NCHW [B,3,H,W] = (img / 255 - mean) / std, bit-exact. */
int fm_im2tensor_f32(float* out, const void* img_u8, int B, int H, int W, float mean, float stdv, void* stream);
/* MaxPool2d(3, stride 2, padding 1): [B,H,W,cs] -> [B,(H-1)/2+1,(W-1)/2+1,cs]. */
int fm_maxpool3x3s2_nhwc(void* out, const void* x, int B, int H, int W, int cs, void* stream);
/* Non-overlapping ph x pw average pooling, output fp32 NCHW [B,C,H/ph,W/pw]. */
int fm_avgpool_nhwc_to_nchw(float* out, const void* x, int B, int H, int W, int C, int cs,
                            int ph, int pw, void* stream);
/* sum_bc[b,c] += sum over pixels (fp32 atomics; caller zero-initialises once, fm_se_gate re-zeroes). */
int fm_channel_sum_nhwc(float* sum_bc, const void* x, int B, int HW, int C, int cs, void* stream);
/* gate[b,c] = sigmoid(w2[C,Cr] relu(w1[Cr,C] (sum[b,:]*inv_hw))). */
int fm_se_gate(float* gate_bc, float* sum_bc, float inv_hw, const float* w1, const float* w2,
               int B, int C, int Cr, void* stream);
/* out = r * gate[b,c] + shortcut[b, y*sc_stride, x*sc_stride, c]. */
int fm_se_combine_nhwc(void* out, const void* r, const float* gate_bc, const void* sc,
                       int B, int H, int W, int C, int cs, int sc_H, int sc_W, int sc_cs,
                       int sc_stride, void* stream);
/* F.interpolate(mode='bilinear', align_corners=True). */
int fm_bilinear_up_nhwc(void* out, const void* x, int B, int IH, int IW, int OH, int OW, int cs,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FM3D_H_ */
