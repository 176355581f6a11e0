/*
 * sup3r_b200 C ABI  --  libsup3r_b200.so
 *
 * Drop-in boundary for the sup3r GAN hot path (Sup3rGan generator / discriminator forward and
 * backward, losses, optimiser step).  The reference (NREL/sup3r @ dd96e798) has NO native FFI:
 * its seam is the Python object protocol of phygnn.CustomNetwork + TensorFlow autodiff that
 * sup3r.models consumes (sup3r/models/abstract.py:96-101, 1081-1092, 1157-1165, 1230-1238;
 * sup3r/models/base.py:283-313, 505-549).  Each entry point below names the reference call it
 * stands in for.  Conventions:
 *   - every pointer is a DEVICE pointer owned by the caller unless stated otherwise; the library
 *     allocates nothing after s3_init (one small internal scratch per device excepted);
 *   - tensors are channels-last and dense: 4-D (n, y, x, c), 5-D (n, z, y, x, c); a 2-D op is a
 *     3-D op with z extent 1 and kernel/stride 1 on z.  sup3r's (n, s1, s2, t, c) maps to
 *     z = s1, y = s2, x = t; (n, s1, s2, c) maps to y = s1, x = s2;
 *   - all calls are asynchronous on the given CUDA stream (a cudaStream_t passed as void*);
 *   - return value 0 = ok, < 0 = error (message from s3_last_error(), thread-local);
 *   - no C++ exceptions, no exit(), no implicit device synchronisation.
 */
#ifndef SUP3R_B200_H_
#define SUP3R_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S3_OK 0
#define S3_ERR_INVALID (-1)
#define S3_ERR_CUDA (-2)
#define S3_ERR_UNSUPPORTED (-3)

#define S3_PAD_ZERO 0      /* keras padding='same' / tf.pad CONSTANT */
#define S3_PAD_REFLECT 1   /* tf.pad REFLECT (phygnn FlexiblePadding) */
#define S3_PAD_SYMMETRIC 2 /* tf.pad SYMMETRIC */

#define S3_ACT_NONE 0
#define S3_ACT_RELU 1
#define S3_ACT_LEAKY 2
#define S3_ACT_SIGMOID 3
#define S3_ACT_TANH 4

typedef void* s3_stream; /* cudaStream_t */

/* ------------------------------------------------------------------------------------------
 * Convolution descriptor.  One fused op replaces the reference's
 *   FlexiblePadding -> Conv2D/Conv3D/Conv2DTranspose -> Cropping -> [LeakyReLU|relu] ->
 *   [SpatialExpansion|SpatioTemporalExpansion] -> [SkipConnection add]
 * layer run (sup3r/configs/ ** /gen_*.json driven by abstract.py:1081-1092).
 * ---------------------------------------------------------------------------------------- */
typedef struct s3_conv_desc {
  int32_t ndim;        /* 2 or 3 convolved dims (2: z extent must be 1) */
  int32_t n;           /* batch */
  int32_t in_dims[3];  /* input extents (z, y, x) */
  int32_t cin, cout;
  int32_t ksize[3];    /* kernel extents, 1 on z for ndim == 2 */
  int32_t stride[3];
  int32_t pad_lo[3];   /* implicit padding in input voxels */
  int32_t pad_hi[3];
  int32_t pad_mode;    /* S3_PAD_ZERO | S3_PAD_REFLECT */
  int32_t act;         /* S3_ACT_*, applied to conv + bias */
  float alpha;         /* LeakyReLU slope */
  /* epilogue scatter: conv-output voxel (z, y, x), channel c -> destination element */
  int32_t d2s;         /* depth_to_space factor r over the spatial dims (3-D: z,y; 2-D: y,x) */
  int32_t d2t;         /* depth_to_time factor m over x (3-D only), 1 = none */
  int32_t t_roll;      /* tf.roll shift applied after depth_to_time */
  int32_t out_repeat[3]; /* nearest-neighbour repeat of the mapped output per dim (>= 1) */
  int32_t out_cstride; /* channels per voxel of the destination buffer (0 = mapped channels) */
  int32_t out_coffset; /* first destination channel (concat-by-stride) */
  /* channel slice of a wider convolution (both 0 = the whole convolution): this call computes
   * output channels [cout_base, cout_base + cout) of a convolution with cout_total channels --
   * w / bias are the slice, the scatter map (d2s / d2t) is that of the full convolution.  Lets a
   * head with more than 256 channels (e.g. 64 -> 1600 with 5x depth_to_space) run as several
   * tcgen05 launches into one destination. */
  int32_t cout_total;
  int32_t cout_base;
  /* 1: the f32 `residual` is added BEFORE the activation (partial sums of a convolution split
   * over input-channel groups, e.g. 64 tensor-core channels + the Sup3rConcat exo channel);
   * 0: after it (SkipConnection). */
  int32_t res_pre_act;
} s3_conv_desc;

/* conv_dims: conv output extents before the scatter; out_dims/out_channels: after it. */
int s3_conv_out_dims(const s3_conv_desc* d, int32_t conv_dims[3], int32_t out_dims[3],
                     int32_t* out_channels);

/* Operand format and tuning knobs of the tcgen05 kernels (all 0 = library default). */
#define S3_FMT_BF16 0  /* bf16 operands (x_lo / w_lo given: three-pass bf16x3 split) */
#define S3_FMT_FP16 1  /* fp16 operands */
#define S3_FMT_FP16C 2 /* fp16 operands + e4m3 correction rows: x_lo / w_lo are the "corr" tensors
                          written by s3_pack_act_pad16 / s3_pack_weights_umma_c / the kernels'
                          own epilogues; one fp16 MMA pass + one e4m3 MMA pass (same time as an
                          fp16 pass) give ~2^-15 relative operand precision */
typedef struct s3_umma_tuning {
  int32_t tiles;            /* M tiles (128 voxels) per CTA work item, 1..8 */
  int32_t w_stages;         /* weight ring depth */
  int32_t box_x;            /* smem x extent of the activation box (>= 10) */
  int32_t box_y;            /* ring kernel: bit flags for A/B measurements, 0 = product path --
                               2 / 4 no plane / weight TMA loads after the first item, 8 no
                               epilogue work, 16 generic MMA role + thread-per-row epilogue,
                               64 no TMA stores, 1024 y-halo rows from registers (all but 16 and
                               1024 produce wrong results: timing experiments only) */
  int32_t max_ctas;         /* 0 = SM count */
  int32_t fmt;              /* S3_FMT_* */
  void* trace;              /* optional device buffer of 16 int64: role timings of CTA 0 */
  int32_t scheme;           /* reserved (0) */
  int32_t ring_slots;       /* ring kernel: activation plane slots in shared memory (0 = as many
                               as fit); tile kernel: experiment flags (8 = no epilogue work) */
  float acc_scale;          /* accumulator -> value factor applied before the bias (0 = 1): the
                               inverse of the power-of-two weight scale of S3_FMT_FP16C */
} s3_umma_tuning;

int s3_init(int device);
const char* s3_last_error(void);
int s3_version(void);
int s3_sm_count(int device);

/* ---- generic fp32 direct convolution (CUDA cores; every geometry) ------------------------
 * x: (n, z, y, x, cin) f32.  w: keras kernel layout (kz, ky, kx, cin, cout) f32.
 * bias: [cout] or NULL.  residual: conv-output geometry (n, Zo, Yo, Xo, cout) f32 or NULL,
 * added after the activation.  post_scale/post_shift: [cout] or NULL, applied last
 * (un-normalisation, abstract.py:240-275).  y: mapped f32 destination or NULL.
 * y_hi / y_lo: optional 16-bit padded+mirrored destinations (see s3_pack_act_pad16). */
int s3_conv_fwd_f32(const s3_conv_desc* d, const float* x, const float* w, const float* bias,
                    const float* residual, const float* post_scale, const float* post_shift,
                    float* y, void* y_hi, void* y_lo, s3_stream stream);

/* Narrow 3-D convolution (3x3x3, stride 1, pad 1, cin <= 8, cout <= 8, plain output map) on
 * the warp-level tensor cores: x and w are rounded to bf16, products accumulate in fp32.  The
 * single-pass "bf16" precision mode uses it for the generators' high-resolution output
 * convolution (last FlexiblePadding -> Conv3D -> Cropping3D of
 * sup3r/configs/spatiotemporal/gen_*.json); same operands / epilogue as s3_conv_fwd_f32.
 * Input: x (f32) or x_bf16 (unpadded (n, z, y, x, 8) bf16, e.g. the 16-bit depth_to_space
 * destination of s3_conv_fwd_umma) -- exactly one of the two. */
int s3_conv_fwd_small_bf16(const s3_conv_desc* d, const float* x, const void* x_bf16,
                           const float* w, const float* bias, const float* residual,
                           const float* post_scale, const float* post_shift, float* y,
                           s3_stream stream);

/* The same with fp16 operands (the S3_FMT_FP16C mode's output layer; x_fp16 = the unpadded fp16
 * depth_to_space destination of s3_conv_fwd_umma). */
int s3_conv_fwd_small_fp16(const s3_conv_desc* d, const float* x, const void* x_fp16,
                           const float* w, const float* bias, const float* residual,
                           const float* post_scale, const float* post_shift, float* y,
                           s3_stream stream);

/* Adjoint of the convolution w.r.t. its input: dy in conv-output geometry -> dx (n,z,y,x,cin).
 * Stands in for tape.gradient through keras Conv* (abstract.py:1230-1238). */
int s3_conv_dgrad_f32(const s3_conv_desc* d, const float* dy, const float* w, float* dx,
                      s3_stream stream);
/* Weight / bias gradients: dw (kz,ky,kx,cin,cout), dbias [cout] (either may be NULL).
 * Results OVERWRITE the destinations.  Deterministic: the voxels are split over CTAs, every split
 * writes its partial dw to `scratch` (>= s3_conv_wgrad_scratch_bytes(); may be NULL when that is
 * 0) and a second kernel adds the splits in a fixed order; the bias gradient likewise. */
size_t s3_conv_wgrad_scratch_bytes(const s3_conv_desc* d);
int s3_conv_wgrad_f32(const s3_conv_desc* d, const float* x, const float* dy, float* dw,
                      float* dbias, void* scratch, s3_stream stream);

/* Weight gradient on tcgen05 (tape.gradient w.r.t. the conv kernels, abstract.py:1190-1238) of
 * the 3x3x3 stride-1 reflect-pad-1 convolutions with 64 output channels:
 *   dw[dz][dy][dx][ci][co] = scale * sum_v x_pad[v + (dz,dy,dx)][ci] * g[v][co]
 * x_hi: fp16 padded input (n, z+2, y+2, x+2, 64) with its REFLECT halo (s3_pack_act_pad16_ex);
 * g_hi: fp16 output gradient padded with g_halo (1 or 2) ZERO voxels per side,
 * (n, z+2h, y+2h, x+2h, 64) -- the tensor the input-gradient convolution consumes has h = 2.
 * Voxels are the GEMM's K dimension (MN-major SWIZZLE_128B operands straight from the TMA boxes),
 * split over CTAs; ws (>= s3_conv_wgrad_umma_ws_bytes) holds the partial sums, reduced in a fixed
 * order.  dw: (3, 3, 3, cin, 64) f32, overwritten; cin <= 64 (channels beyond cin are padding). */
size_t s3_conv_wgrad_umma_ws_bytes(int n, int z, int y, int x);
int s3_conv_wgrad_umma(const void* x_hi, const void* g_hi, int g_halo, int n, int z, int y, int x,
                       int cin, float scale, float* dw, void* ws, size_t ws_bytes,
                       s3_stream stream);

/* ---- tcgen05 implicit-GEMM convolution (cin == 64, 3x3[x3], stride 1, reflect-1) ---------
 * x_hi/x_lo: padded+mirrored 16-bit activations (lo NULL = single pass; bf16 lo = the 3-pass
 * split-precision product hi*hi + lo*hi + hi*lo; S3_FMT_FP16C lo = e4m3 correction rows, one
 * extra e4m3 pass).  w_hi/w_lo: from s3_pack_weights_umma[_c].
 * The SkipConnection addend is either `residual` (f32, output layout) or the pair res_hi/res_lo
 * (16-bit padded layout of y_hi; value = hi + lo; res_lo may be NULL).
 * y_hi: padded+mirrored 16-bit destination for plain output maps; for depth_to_space /
 * depth_to_time maps (8-channel runs) an UNPADDED 16-bit tensor of the mapped geometry. */
int s3_conv_fwd_umma(const s3_conv_desc* d, const void* x_hi, const void* x_lo, const void* w_hi,
                     const void* w_lo, const float* bias, const float* residual,
                     const void* res_hi, const void* res_lo, const float* post_scale,
                     const float* post_shift, float* y, void* y_hi, void* y_lo,
                     const s3_umma_tuning* tune, s3_stream stream);
/* rows of the packed weight tensor per tap (cout rounded up to a multiple of 16) */
int s3_umma_npad(int cout);
/* Weight layout the tcgen05 kernel expects for a conv with this rank / cout:
 * 0 = tap-major (taps, npad, 64); 1 = "zcat" (9 (dy,dx), 3 (dz), npad, 64) (3-D, 3*npad <= 256). */
int s3_umma_weight_layout(int ndim, int cout, int split);
/* w (taps, cin=64, cout) f32 (keras tap order dz, dy, dx) -> 16-bit packed tensor in `layout`,
 * cout rows zero padded to npad.  w_lo NULL = no split. */
int s3_pack_weights_umma(const float* w, int taps, int cin, int cout, void* w_hi, void* w_lo,
                         int fmt, int layout, s3_stream stream);
/* S3_FMT_FP16C weights: w_hi = fp16(w * scale), w_corr = e4m3 correction rows; `scale` is a
 * power of two chosen by the caller so that max|w| * scale lies in [2^13, 2^14) (the kernel is
 * then called with s3_umma_tuning.acc_scale = 1 / scale). */
int s3_pack_weights_umma_c(const float* w, int taps, int cin, int cout, void* w_hi, void* w_corr,
                           float scale, int layout, s3_stream stream);
/* The same packing of a VIEW of the keras kernel w (taps, src_cin, src_cout), so that the
 * training step never materialises sliced / padded / flipped copies of its weights:
 *   adjoint == 0: rows ci0 .. ci0+63 (input channels), columns co0 .. co0+cout-1;
 *   adjoint == 1: the operand of the input-gradient convolution (tape.gradient w.r.t. the
 *                 layer input, abstract.py:1230-1238): taps flipped, row r = output channel
 *                 ci0 + r of w, column c = input channel co0 + c of w.
 * Elements outside w are zero. */
int s3_pack_weights_umma_view(const float* w, int taps, int src_cin, int src_cout, int ci0,
                              int co0, int adjoint, int cout, void* w_hi, void* w_corr,
                              float scale, int layout, s3_stream stream);
/* f32 (n, z, y, x, c) -> 16-bit (n, z+2*pz, y+2, x+2, c), pz = (ndim == 3), reflect halo.
 * fmt S3_FMT_FP16C (c == 64): lo receives the e4m3 correction rows (same byte geometry). */
int s3_pack_act_pad16(const float* x, int ndim, int n, const int32_t dims[3], int c, void* hi,
                      void* lo, int fmt, s3_stream stream);
/* The same with the halo content chosen: S3_PAD_REFLECT (forward operands) or S3_PAD_ZERO (the
 * zero-extended output gradient of the tcgen05 dgrad: tape.gradient through the reflect-padded
 * convolution, sup3r/models/abstract.py:1230-1238, is a zero-padded correlation with the flipped
 * kernel followed by the adjoint of the reflect pad, s3_pad_bwd). */
int s3_pack_act_pad16_ex(const float* x, int ndim, int n, const int32_t dims[3], int c, void* hi,
                         void* lo, int fmt, int halo_mode, s3_stream stream);
/* Same with a zero halo of `halo_width` voxels (1 or 2): width 2 is the operand of the input-
 * gradient convolution of a reflect-padded layer (the gradient on the padded extent, zero-padded
 * by one more voxel), without materialising the padded fp32 tensor. */
int s3_pack_act_pad16_hw(const float* x, int ndim, int n, const int32_t dims[3], int c, void* hi,
                         void* lo, int fmt, int halo_mode, int halo_width, s3_stream stream);
/* inverse (interior only); lo may be NULL.  For tests. */
int s3_unpack_act_pad16(const void* hi, const void* lo, int ndim, int n, const int32_t dims[3],
                        int c, float* x, int fmt, s3_stream stream);

/* ---- stand-alone layers (eager path; the literal reference op order) ---------------------
 * Shapes are given as up to 5 leading extents + channels: dims[0..4] (unused leading = 1). */
/* tf.pad (phygnn FlexiblePadding.call) over dims (n, z, y, x) + channel pads. */
int s3_pad_fwd(const float* x, float* y, const int32_t dims[5], const int32_t lo[5],
               const int32_t hi[5], int mode, s3_stream stream);
int s3_pad_bwd(const float* dy, float* dx, const int32_t dims[5], const int32_t lo[5],
               const int32_t hi[5], int mode, s3_stream stream);
/* keras Cropping2D/3D (also used as its own adjoint helper: bwd zero-fills the border). */
int s3_crop_fwd(const float* x, float* y, const int32_t dims[5], const int32_t lo[5],
                const int32_t hi[5], s3_stream stream);
int s3_crop_bwd(const float* dy, float* dx, const int32_t dims[5], const int32_t lo[5],
                const int32_t hi[5], s3_stream stream);
/* LeakyReLU / Activation.  bwd takes the layer OUTPUT y (sign / value recompute). */
int s3_act_fwd(const float* x, float* y, size_t n, int act, float alpha, s3_stream stream);
int s3_act_bwd(const float* y, const float* dy, float* dx, size_t n, int act, float alpha,
               s3_stream stream);
/* y = a + b (SkipConnection second call, Sup3rAdder).  b is broadcast with period nb. */
int s3_add(const float* a, const float* b, float* y, size_t n, size_t nb, s3_stream stream);
/* SpatialExpansion / SpatioTemporalExpansion.  x: (n, z, y, x, c); method 0 nearest, 1 d2t. */
int s3_expand_fwd(const float* x, float* y, int ndim, int n, const int32_t dims[3], int c,
                  int spatial_mult, int temporal_mult, int method, int t_roll, s3_stream stream);
int s3_expand_bwd(const float* dy, float* dx, int ndim, int n, const int32_t dims[3], int c,
                  int spatial_mult, int temporal_mult, int method, int t_roll, s3_stream stream);
/* Sup3rConcat: y[v, :ca] = a[v], y[v, ca:] = b[v].  split = adjoint. */
int s3_concat_fwd(const float* a, int ca, const float* b, int cb, float* y, size_t nvox,
                  s3_stream stream);
int s3_concat_bwd(const float* dy, float* da, int ca, float* db, int cb, size_t nvox,
                  s3_stream stream);
/* y[v, c] = x[v, c] * scale[c] + shift[c]   (norm_input / un_norm_output, abstract.py:197-275) */
int s3_channel_affine(const float* x, float* y, size_t nvox, int c, const float* scale,
                      const float* shift, s3_stream stream);
/* Dense: y[m, n] = act(x[m, :] @ w[:, n] + b[n]);  w is (k, n) row-major (keras). */
int s3_dense_fwd(const float* x, const float* w, const float* b, float* y, int m, int k, int n,
                 int act, float alpha, s3_stream stream);
int s3_dense_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw,
                 float* db, int m, int k, int n, s3_stream stream);

/* ---- losses and optimiser ------------------------------------------------------------------
 * Content loss (base.py:478-503): kind 0 MeanSquaredError, 1 MeanAbsoluteError over the first
 * `c_use` of `c` channels of gen/true (exo channels sliced off).  loss: device scalar
 * (overwritten).  dgen: NULL or gradient of (weight * loss) w.r.t. gen (zeros on unused ch).
 * scratch: S3_LOSS_SCRATCH_FLOATS device floats (block partial sums, added in block order: the
 * loss is bit-reproducible from run to run). */
#define S3_LOSS_SCRATCH_FLOATS 1025
int s3_content_loss(const float* gen, const float* truth, size_t nvox, int c, int c_use, int kind,
                    float weight, float* loss, float* dgen, float* scratch, s3_stream stream);
/* Relativistic average discriminator loss (base.py:505-549) for logits (b,) each.
 * loss: device scalar; d_real / d_fake: NULL or gradients scaled by `weight`. */
int s3_loss_disc(const float* out_real, const float* out_fake, int b, float weight, float* loss,
                 float* d_real, float* d_fake, s3_stream stream);
/* keras Adam step over a flat arena (abstract.py:899,912): step is the 1-based iteration. */
int s3_adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1,
                 float beta2, float eps, int64_t step, s3_stream stream);
/* y (fp16, saturating) = x: halves the device -> host transfer of the high-resolution result when
 * the consumer stores 16-bit data anyway (ForwardPass output_dtype, cf. the scaled integer dtypes
 * of sup3r/writers). */
int s3_cast_f16(const float* x, void* y, size_t n, s3_stream stream);
/* out[0] = sum(x), out[1] = sum(|x|), out[2] = #nan-or-inf, out[3] = min, out[4] = max */
int s3_stats(const float* x, size_t n, float* out5, s3_stream stream);
/* per-channel min/max + NaN count for ForwardPass._output_check (forward_pass.py:384-425):
 * out: [c][3] = (min, max, n_nan). */
int s3_channel_check(const float* x, size_t nvox, int c, float* out, s3_stream stream);

/* ---- chunk pipeline around the generator --------------------------------
 * Output post-processing of one cropped hi-res chunk, IN PLACE on data (n_spatial, n_t, n_f):
 * every (pair_u[k], pair_v[k]) channel pair is rotated by the grid angle (cos_sin: (n_spatial, 2) =
 * cos / sin theta per hi-res grid point) and replaced by (windspeed, winddirection [deg]) --
 * sup3r/preprocessing/derivers/utilities.py:204-255 invert_uv, writers/base.py:233-295 -- then
 * every channel is limited to [lo, hi] (clip != 0; utilities/utilities.py:155-220 enforce_limits
 * without nn_fill).  pair_u, pair_v, lo, hi are HOST arrays (read before the launch); data,
 * cos_sin and counts are device pointers.  counts[2 f], counts[2 f + 1] (device, overwritten): points of channel f
 * below lo / above hi before clipping (the reference's warning fractions; the nn_fill trigger). */
int s3_output_transform(float* data, size_t n_spatial, int n_t, int n_f, const float* cos_sin,
                        const int* pair_u, const int* pair_v, int n_pairs, const float* lo,
                        const float* hi, int clip, unsigned long long* counts, s3_stream stream);
/* Batch production (sup3r/preprocessing/batch_queues/base.py:32-87): hr (n, s1, s2, t, f) ->
 * lr (n, s1/s, s2/s, t/te, f): (s x s) block mean, then temporal method 0 subsample, 1 average,
 * 2 total, 3 max, 4 min (utilities/utilities.py:345-523).  4-D tensors: t = te = 1. */
int s3_coarsen(const float* hr, float* lr, int n, int s1, int s2, int t, int f, int s_enhance,
               int t_enhance, int method, s3_stream stream);
/* scipy.ndimage.gaussian_filter(sigma, mode='nearest') over the (s1, s2) planes of x
 * (n, s1, s2, tf) for the features whose bit is set in feature_mask (channel = index % f);
 * weights: 2 radius + 1 normalised taps (host-computed); tmp: scratch of the size of x
 * (sup3r/preprocessing/batch_queues/utilities.py:57-104 smooth_data). */
int s3_gauss_smooth2d(const float* x, float* tmp, float* y, int n, int s1, int s2, int tf, int f,
                      unsigned feature_mask, const float* weights, int radius, s3_stream stream);
/* Sampler: out[b] = data[i0:i0+s1, j0:j0+s2, t0:t0+t, :] for origins (n, 3) int32 on the device
 * (the device-resident stand-in for sup3r/preprocessing/samplers/base.py get_next). */
int s3_gather_samples(const float* data, int S1, int S2, int T, int F, const int* origins, int n,
                      int s1, int s2, int t, float* out, s3_stream stream);
/* Empirical quantile delta mapping of a low-res chunk (sup3r/bias/bias_transforms.py:490-619
 * _apply_qdm -> rex.utilities.bc_utils.QuantileDeltaMapping, dist "empirical"):
 *   tau = F_mf(x), x_oh = F_oh^-1(tau), x_mh = F_mh^-1(tau),
 *   out = relative ? x_oh * (x / x_mh) : x_oh + (x - x_mh)
 * data / out: (n_sites, n_times) fp32; window[t]: time-window index of time step t; params_*:
 * (n_sites, n_windows, n_quantiles) fp32 quantile values of the observed-historical, modeled-
 * historical and modeled-future distributions (pass params_mh as params_mf for no_trend);
 * quantiles: the n_quantiles levels (device doubles).  delta_denom_zero / delta_denom_min /
 * delta_range[2] / out_range[2]: HOST pointers or NULL.  Interpolation restates np.interp in
 * double precision.  tau_fut (n_sites, fp32) + k_factor (n_sites, n_windows, DOUBLES: the
 * reference's k_range clipping promotes them), or both NULL: the PresRat
 * step of local_presrat_bc (bias_transforms.py:1117-1120): results below tau_fut become 0, the
 * others are scaled by the window's K factor.  n_bad: 2 device counters {non-finite results,
 * NaN results} (local_qdm_bc raises on the first, local_presrat_bc on the second). */
int s3_qdm_bc(const float* data, const int* window, const float* params_oh,
              const float* params_mh, const float* params_mf, const double* quantiles,
              const float* tau_fut, const double* k_factor, int n_sites, int n_times,
              int n_windows, int n_quantiles, int relative,
              const double* delta_denom_zero, const double* delta_denom_min,
              const double* delta_range, const double* out_range, float* out,
              unsigned long long* n_bad, s3_stream stream);
/* Gradient SUM over NVLink peer memory (sup3r/models/abstract.py:785-805 _sum_parallel_grad):
 * out[i] = peer_ptrs[0][i] + peer_ptrs[1][i] + ... in RANK ORDER (deterministic, identical on
 * every rank).  peer_ptrs: HOST array of `world` device pointers, one per rank, to the ranks' flat
 * fp32 arenas in symmetric (peer-mapped) memory; the caller brackets the call with cross-GPU
 * barriers (arenas written / arenas read).  n: floats, a multiple of 4. */
int s3_peer_sum_f32(const void* const* peer_ptrs, int world, size_t n, float* out,
                    s3_stream stream);
/* The optimiser step fused into the gradient exchange: for every tensor of the table
 * g = sum over ranks (rank order) of the gradient arenas at [off, off + n), then the keras Adam
 * update of s3_adam_step on (p, m, v) -- one launch for the whole network.  segs_dev: DEVICE array
 * of n_seg records {float* p, m, v; uint64 off, n}; max_n: the largest n; step: 1-based. */
int s3_peer_sum_adam(const void* const* peer_ptrs, int world, const void* segs_dev, int n_seg,
                     unsigned long long max_n, float lr, float beta1, float beta2, float eps,
                     long long step, s3_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* SUP3R_B200_H_ */
