/* wu_b200.h — C ABI of libwu_b200.so: the sm_100a kernels behind the cUNet generator hot path.
 *
 * The reference (Sota0726/weather-Unet) has no FFI: its generator is Python calling ATen/cuDNN.
 * Each entry point below therefore names the reference call site(s) (file:line in the reference
 * tree) whose arithmetic it replaces.  Conventions:
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch); the library never allocates
 *     persistent memory and never frees caller memory; scratch comes in through `workspace`;
 *   - activations are NHWC bf16 ([B][H][W][C], C contiguous) unless stated; parameters and
 *     parameter gradients are fp32 in the reference's own layouts (state_dict shapes);
 *   - every call is asynchronous on `stream` (a cudaStream_t), re-entrant, and keeps no mutable
 *     global state; returns WU_OK or an error code, message via wu_last_error() (thread-local);
 *   - no CPU fallback exists: a non-sm_100 device is an error.
 */
#ifndef WU_B200_H_
#define WU_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WU_OK 0
#define WU_ERR_INVALID 1     /* bad shape / argument */
#define WU_ERR_CUDA 2        /* CUDA runtime / driver error */
#define WU_ERR_UNSUPPORTED 3 /* device is not sm_100 */

typedef void* wu_stream_t; /* cudaStream_t */

const char* wu_last_error(void);
int wu_version(void);
/* 0 when the current device can run the kernels (compute capability 10.x). */
int wu_device_check(void);

/* ---- weights ---------------------------------------------------------------------------------
 * nn.Conv2d(cin, cout, 3, padding=1).weight (nets.py:20,22) fp32 [cout][cin][3][3]  ->
 *   w_fprop bf16 [cout][9*cin]  with k = (r*3+s)*cin + ci           (B operand of fprop)
 *   w_dgrad bf16 [cin][9*cout]  with k = ((2-r)*3+(2-s))*cout + co  (B operand of dgrad; may be NULL)
 */
int wu_pack_conv3x3_weights(const float* w, int cout, int cin, void* w_fprop, void* w_dgrad,
                            wu_stream_t stream);

/* ---- 3x3 / stride 1 / pad 1 convolution as tcgen05 implicit GEMM ------------------------------
 * Replaces nn.Conv2d(…,3,padding=1) + nn.ReLU (nets.py:18-24) and, with src1 != NULL, the
 * torch.cat([x, skip], 1) feeding it (cunet.py:62,69,76): the K loop walks src0's c0 channels and
 * then src1's c1 channels, so the concatenated tensor is never materialised.
 *   dst[b,h,w,co] = act( bias[co] + sum_{r,s,ci} src[b,h+r-1,w+s-1,ci] * w_packed[co][(r*3+s)*C+ci] )
 * then, if relu_mask_src != NULL (same shape as dst), dst = relu_mask_src > 0 ? dst : 0.
 * The same entry point is the data-gradient pass (autograd of nets.py:20,22): call it with
 * src0 = dY, w_packed = w_dgrad, bias = NULL, relu = 0 and relu_mask_src = the forward output of
 * the layer below (its ReLU mask).  c0, c1, cout must be multiples of 64; bias may be NULL.
 */
int wu_conv3x3_fprop(const void* src0, int c0, const void* src1, int c1, const void* w_packed,
                     const float* bias, int relu, const void* relu_mask_src, void* dst, int cout,
                     int B, int H, int W, wu_stream_t stream);

/* Same, for "one image x many conditions" inference (inference/inf_1year_signals.py:98-107,
 * dataset.py:200-203): with src1_bcast != 0 the skip source src1 has batch 1 and is shared by all B
 * images of src0 (its TMA batch coordinate is pinned to 0), so the encoder runs once. */
int wu_conv3x3_fprop_bcast(const void* src0, int c0, const void* src1, int c1, int src1_bcast,
                           const void* w_packed, const float* bias, int relu,
                           const void* relu_mask_src, void* dst, int cout, int B, int H, int W,
                           wu_stream_t stream);

/* Convolution + ReLU that also emits AdaIN's statistics of its own output (utils.py:34-39 on the
 * tensor cunet.py:59,66,73 hands to AdaIN): stats fp32 [Bout][chunks][cout][2] = (sum, sum of squares)
 * of the STORED bf16 values per 64-pixel chunk, chunks = wu_conv3x3_stats_chunks(cout, H, W) (0 when
 * the shape has no fused variant: cout must be 128 or a multiple of 256).  Feed it to
 * wu_adain_style_fwd_n instead of running wu_adain_stats over the tensor again. */
int wu_conv3x3_stats_chunks(int cout, int H, int W);
int wu_conv3x3_fprop_stats(const void* src0, int c0, const void* src1, int c1, int src1_bcast,
                           const void* w_packed, const float* bias, void* dst, float* stats, int cout,
                           int B, int H, int W, wu_stream_t stream);

/* The generator's last two layers in one kernel (cunet.py:78-82): dst = relu(conv3x3(src) + bias)
 * with 64 output channels (dconv_up1.2, kept for the backward pass) and, from the same registers,
 *   y[b,o,h,w] = tanh(last_b[o] + sum_c last_w[o][c] * dst[b,h,w,c])       (conv_last + Tanh)
 * last_w fp32 [3][64], last_b fp32 [3] (may be NULL), y fp32 NCHW [B][3][H][W]. */
int wu_conv3x3_fprop_last(const void* src, int cin, const void* w_packed, const float* bias, void* dst,
                          const float* last_w, const float* last_b, float* y, int B, int H, int W,
                          wu_stream_t stream);

/* Convolution + ReLU + nn.MaxPool2d(2) in one kernel (cunet.py:45-46, 48-49): dst as wu_conv3x3_fprop
 * with relu = 1 (kept: it is the skip tensor), pool_dst NHWC bf16 [B][H/2][W/2][cout] = 2x2 max of dst,
 * taken from the staged tile in shared memory.  cout = 64 or 128; H, W even. */
int wu_conv3x3_fprop_pool(const void* src, int cin, const void* w_packed, const float* bias, void* dst,
                          void* pool_dst, int cout, int B, int H, int W, wu_stream_t stream);

/* Weight + bias gradient of the same convolution (autograd of nets.py:20,22):
 *   dw[co][ci][r][s] = sum_{b,h,w} dy[b,h,w,co] * src[b,h+r-1,w+s-1,ci]     (fp32, overwritten)
 *   db[co]           = sum_{b,h,w} dy[b,h,w,co]                              (fp32, may be NULL)
 * src is the two-source concatenation as in fprop.  Split-K partials live in `workspace`
 * (wu_conv3x3_wgrad_workspace_bytes); the result is deterministic for a fixed shape.
 */
size_t wu_conv3x3_wgrad_workspace_bytes(int cin_total, int cout, int B, int H, int W);
int wu_conv3x3_wgrad(const void* src0, int c0, const void* src1, int c1, const void* dy, int cout,
                     int B, int H, int W, float* dw, float* db, void* workspace,
                     size_t workspace_bytes, wu_stream_t stream);

/* ---- first layer: Conv2d(3,64,3,padding=1)+ReLU on the NCHW fp32 image (cunet.py:21,45) -------
 * x fp32 NCHW [B][3][H][W]; w fp32 [64][3][3][3]; dst NHWC bf16 [B][H][W][64].  HBM/FMA bound
 * (K = 27), so it is a direct convolution, not a tensor-core GEMM. */
int wu_conv_first_fprop(const float* x, const float* w, const float* bias, void* dst, int B, int H,
                        int W, wu_stream_t stream);
/* dy NHWC bf16 [B][H][W][64] (already ReLU-masked) -> dw fp32 [64][3][3][3], db fp32 [64]. */
size_t wu_conv_first_wgrad_workspace_bytes(int B, int H, int W);
int wu_conv_first_wgrad(const float* x, const void* dy, float* dw, float* db, int B, int H, int W,
                        void* workspace, size_t workspace_bytes, wu_stream_t stream);

/* ---- last layer: Conv2d(64,3,1) + Tanh (cunet.py:39-40,80-82) ---------------------------------
 * x NHWC bf16 [B][H][W][64]; w fp32 [3][64]; y fp32 NCHW [B][3][H][W]. */
int wu_conv_last_tanh_fprop(const void* x, const float* w, const float* bias, float* y, int B,
                            int H, int W, wu_stream_t stream);
/* gy, y fp32 NCHW.  gx NHWC bf16 = relu'(x) * (W^T (gy * (1 - y^2)));  dw [3][64], db [3] fp32. */
size_t wu_conv_last_tanh_bprop_workspace_bytes(int B, int H, int W);
int wu_conv_last_tanh_bprop(const float* gy, const float* y, const void* x, const float* w,
                            void* gx, float* dw, float* db, int B, int H, int W, void* workspace,
                            size_t workspace_bytes, wu_stream_t stream);

/* ---- nn.MaxPool2d(2) (cunet.py:27,46,49,52) ---------------------------------------------------
 * src NHWC bf16 [B][H][W][C] -> dst [B][H/2][W/2][C]; C % 8 == 0, H and W even. */
int wu_maxpool2_fwd(const void* src, void* dst, int B, int H, int W, int C, wu_stream_t stream);
/* Gradient arriving at a skip tensor `y` (post-ReLU conv output): the pooled branch routes
 * g_pool to the first maximum of each 2x2 window (PyTorch tie rule), the decoder's skip branch
 * adds g_skip (may be NULL), and the ReLU mask of y is applied:
 *   g[b,h,w,c] = y > 0 ? g_skip + (argmax ? g_pool[b,h/2,w/2,c] : 0) : 0 */
int wu_maxpool2_bwd(const void* y, const void* g_pool, const void* g_skip, void* g, int B, int H,
                    int W, int C, wu_stream_t stream);

/* ---- AdaIN (utils.py:26-51) fused with Upsample(x2, bilinear, align_corners) + Dropout --------
 * Step 1: per-(b,c) partial sums of x and x^2 over H*W (utils.py:36-38).
 *   partial fp32 [B][nchunk][C][2], nchunk = wu_adain_stats_chunks(H*W). */
int wu_adain_stats_chunks(int HW);
int wu_adain_stats(const void* x, float* partial, int B, int HW, int C, wu_stream_t stream);
/* Step 2: style = l1(cond) (utils.py:46) -> y_mean / y_std over the 4 style numbers per channel,
 * combined with the instance statistics into an affine map (all fp32, [B][C]):
 *   mean, rstd = 1/sqrt(var_unbiased + eps), ystd, scale = ystd*rstd, shift = ymean - mean*scale.
 * cond fp32 [B][nc]; lw fp32 [4C][nc]; lb fp32 [4C].  x_bcast != 0: the statistics (and x in step 3)
 * have batch 1 and serve all B conditions (one image x many signals). */
int wu_adain_style_fwd(const float* cond, const float* lw, const float* lb, const float* partial,
                       float* mean, float* rstd, float* ystd, float* scale, float* shift, int B,
                       int C, int nc, int HW, float eps, int x_bcast, wu_stream_t stream);
/* Same with an explicit chunk count (statistics produced by wu_conv3x3_fprop_stats). */
int wu_adain_style_fwd_n(const float* cond, const float* lw, const float* lb, const float* partial,
                         float* mean, float* rstd, float* ystd, float* scale, float* shift, int B,
                         int C, int nc, int HW, int nchunk, float eps, int x_bcast, wu_stream_t stream);
/* Step 3 (cunet.py:59-61): u[b,Y,X,c] = keep * bilinear_x2(x*scale+shift)[b,Y,X,c] / (1-p).
 * Dropout: p_drop == 0 -> none; else if mask != NULL it is a uint8 NHWC [B][2h][2w][C] keep mask;
 * else keep bits come from a Philox4x32-7 stream keyed by (seed, 8-channel vector index): 15 bits
 * per element, keep iff u15 >= round(p * 32768), so the realised rate is 1-p to within 2^-16.
 * Requires B * (h/2 + 1) <= 65535 (grid.y).
 * keep_bits (required when p_drop > 0): uint8 [B][2h][2w][C/8], bit j = channel 8v+j kept; the
 * backward pass reads it instead of replaying the RNG. */
int wu_adain_up_drop_fwd(const void* x, const float* scale, const float* shift, void* u,
                         uint8_t* keep_bits, int B, int h, int w, int C, float p_drop, uint64_t seed,
                         const uint8_t* mask, int x_bcast, wu_stream_t stream);
/* Same with a device-resident draw counter: `epoch` (device uint32, may be NULL = 0) is read by the
 * kernel and its low 16 bits enter the Philox counter next to the image index, so the same launch
 * replayed from a CUDA graph draws a fresh mask once the caller has advanced *epoch (nn.Dropout draws
 * a new mask per forward: cunet.py:61,68,75).  epoch == NULL or *epoch == 0 reproduces
 * wu_adain_up_drop_fwd bit for bit. */
int wu_adain_up_drop_fwd_epoch(const void* x, const float* scale, const float* shift, void* u,
                               uint8_t* keep_bits, int B, int h, int w, int C, float p_drop,
                               uint64_t seed, const uint32_t* epoch, const uint8_t* mask, int x_bcast,
                               wu_stream_t stream);
/* AdaIN without the fused upsample / dropout (utils.py:49-50 on its own): out = x*scale + shift,
 * scale / shift from wu_adain_style_fwd.  x, out NHWC bf16 [B][HW][C]. */
int wu_adain_apply(const void* x, const float* scale, const float* shift, void* out, int B, int HW,
                   int C, wu_stream_t stream);
/* Backward, step 1: gz = adjoint(dropout o upsample)(gu) at low resolution, plus per-(b,c)
 * partial sums S1 = sum gz, S2 = sum gz * xhat (xhat = (x-mean)*rstd), in one pass over gu.
 *   gz bf16 [B][h][w][C]; partial fp32 [B][nchunk][C][2], nchunk = wu_adain_bwd_chunks(h, w, C). */
int wu_adain_bwd_chunks(int h, int w, int C);
int wu_adain_up_drop_bwd(const void* gu, const void* x, const float* mean, const float* rstd,
                         void* gz, float* partial, int B, int h, int w, int C, float p_drop,
                         const uint8_t* keep_bits, wu_stream_t stream);
/* Backward, step 2: reduce the nchunk partials, emit k1 = S1/N and k2 = S2/(N-1) ([B][C] fp32), the apply
 * coefficients coef fp32 [3][B][C] (A = rstd*ystd, Bc = -A*k2*rstd, Cc = A*(k2*rstd*mean - k1)) and the
 * gradients of l1.weight / l1.bias (utils.py:31,46), overwritten: dlw [4C][nc], dlb [4C].
 * gh is caller scratch, fp32 [B][4C] (gradient of the style vector l1(cond)). */
int wu_adain_style_bwd(const float* cond, const float* lw, const float* lb, const float* partial,
                       int nchunk, const float* ystd, const float* mean, const float* rstd, float* k1,
                       float* k2, float* coef, float* gh, float* dlw, float* dlb, int B, int C, int nc,
                       int HW, wu_stream_t stream);
/* Backward, step 3: gx = relu'(x) * (A*gz + Bc*x + Cc)  ==  relu'(x) * rstd*ystd*(gz - k1 - xhat*k2),
 * bf16 [B][h][w][C]. */
int wu_adain_bwd_apply(const void* gz, const void* x, const float* coef, void* gx, int B, int HW,
                       int C, wu_stream_t stream);

/* ---- bias + LeakyReLU on NHWC bf16 (discriminator blocks, nets.py:26-33) ---------
 * fwd, in place: x = leaky_relu(x + bias[c], slope); slope == 1 is a plain bias add.
 * bwd: g = gy * (y > 0 ? 1 : slope) (g may alias gy), db[c] = sum_px g[px][c] (fp32, overwritten).
 * x, y, gy, g: bf16 [npix][C]; C a power of two in [8, 2048]. */
int wu_bias_act_fwd(void* x, const float* bias, float slope, long long npix, int C,
                    wu_stream_t stream);
size_t wu_bias_act_bwd_workspace_bytes(int C);
int wu_bias_act_bwd(const void* gy, const void* y, void* g, float* db, float slope, long long npix,
                    int C, void* workspace, size_t workspace_bytes, wu_stream_t stream);

/* ---- discriminator stem (disc.py:28, nets.py:26-33 with in_channels = 3) --------
 *   x fp32 NCHW [B][3][H][W] -> h1 = Conv2d(3,3,3,padding=1)(x) fp32 NCHW (no activation)
 *   -> c1 = LeakyReLU(Conv2d(3,64,3,padding=1,stride=2)(h1)) NHWC bf16 [B][H/2][W/2][64].
 * Weights arrive already spectrally normalised (fp32, [cout][cin][3][3]).  K = 27: FMA / HBM bound. */
int wu_conv3to3_fprop(const float* x, const float* w, const float* bias, float* y, int B, int H,
                      int W, wu_stream_t stream);
int wu_conv3to64_s2_fprop(const float* h1, const float* w, const float* bias, float slope, void* dst,
                          int B, int Hin, int Win, wu_stream_t stream);
/* g = gradient at the stride-2 convolution's output, already LeakyReLU-masked (wu_bias_act_bwd with
 * NULL-equivalent bias handling is not needed: use slope masking of the caller), NHWC bf16. */
size_t wu_conv3to64_s2_wgrad_workspace_bytes(int B, int Hin, int Win);
int wu_conv3to64_s2_wgrad(const float* h1, const void* g, float* dw, float* db, int B, int Hin,
                          int Win, void* workspace, size_t workspace_bytes, wu_stream_t stream);
/* "GEMM first, shift afterwards": g is contracted once with all 27 (ci, r, s) weight columns on tcgen05
 * (a 1x1 convolution into a bf16 scratch tensor), then scattered to (2u + r - 1, 2v + s - 1).  Hin, Win
 * even; `workspace` (wu_conv3to64_s2_dgrad_workspace_bytes, 128-byte aligned) holds weights + scratch. */
size_t wu_conv3to64_s2_dgrad_workspace_bytes(int B, int Hin, int Win);
int wu_conv3to64_s2_dgrad(const void* g, const float* w, float* g_h1, int B, int Hin, int Win,
                          void* workspace, size_t workspace_bytes, wu_stream_t stream);
/* Backward of the 3->3 convolution: g_x fp32 NCHW (may be NULL: real images / detached fakes need
 * none), dw [3][3][3][3] (may be NULL: the generator update discards the discriminator's weight
 * gradients), db [3] (may be NULL); workspace is only needed with dw. */
size_t wu_conv3to3_bprop_workspace_bytes(void);
int wu_conv3to3_bprop(const float* g_h1, const float* x, const float* w, float* g_x, float* dw,
                      float* db, int B, int H, int W, void* workspace, size_t workspace_bytes,
                      wu_stream_t stream);

/* ---- discriminator trunk: 3x3 / stride 2 / pad 1 convolution on tcgen05 ----------
 * Replaces spectral_norm(nn.Conv2d(cin, cout, 3, padding=1, stride=2)) + nn.LeakyReLU(0.2)
 * (nets.py:26-33, disc.py:12-15,28-31) and its autograd.  The stride-1 convolution in front of it
 * (nets.py:28-29, no activation) is wu_conv3x3_fprop with relu = 0.  Weights arrive spectrally
 * normalised and packed by wu_pack_conv3x3_weights.  src NHWC bf16 [B][Hin][Win][cin];
 * dst / dy NHWC bf16 [B][Ho][Wo][cout], Ho = ceil(Hin/2), Wo = ceil(Win/2).
 *   fprop: dst = lrelu(bias + conv_s2(src, w_packed), slope)   (slope 1 = no activation)
 *   dgrad: dx[b,yi,xi,ci] = sum_{r,s,co : yi+1-r, xi+1-s even} dy[b,(yi+1-r)/2,(xi+1-s)/2,co] * w[co][ci][r][s]
 *          (w_dgrad from wu_pack_conv3x3_weights; dy already LeakyReLU-masked: wu_bias_act_bwd)
 *   wgrad: dw[co][ci][r][s] = sum_{b,yo,xo} dy[b,yo,xo,co] * src[b,2yo+r-1,2xo+s-1,ci]; db optional.
 * cin, cout multiples of 64; the N dimension (cout for fprop / wgrad, cin for dgrad) must be 64, 128
 * or a multiple of 256. */
int wu_conv3x3_s2_fprop(const void* src, int cin, const void* w_packed, const float* bias,
                        float slope, void* dst, int cout, int B, int Hin, int Win,
                        wu_stream_t stream);
int wu_conv3x3_s2_dgrad(const void* dy, int cout, const void* w_dgrad, void* dx, int cin, int B,
                        int Hin, int Win, wu_stream_t stream);
size_t wu_conv3x3_s2_wgrad_workspace_bytes(int cin, int cout, int B, int Hin, int Win);
int wu_conv3x3_s2_wgrad(const void* src, int cin, const void* dy, int cout, int B, int Hin, int Win,
                        float* dw, float* db, void* workspace, size_t workspace_bytes,
                        wu_stream_t stream);

/* ---- spectral normalisation of the discriminator's weights, multi-tensor -----------
 * torch.nn.utils.spectral_norm as the reference applies it (nets.py:26-33, disc.py:21,24): one power
 * iteration per training-mode forward on W [rows = out][cols = in*kh*kw] (fp32, contiguous),
 *   v <- normalize(W^T u), u <- normalize(W v), sigma = u . (W v), W_sn = W / sigma   (eps on the norms);
 * eval mode (training == 0) uses the stored u, v.  `tensors`: device array of records
 *   {const float* w; float* u; float* v; float* u_snap; float* v_snap; float* sigma; float* t; float* s;
 *    float* part; float* w_sn; const float* g; float* dw; bf16* wf; bf16* wd; int32 rows; int32 cols}
 * (120 bytes).  wf / wd non-NULL (3x3 convolution weights, cols = 9 * cin): W / sigma is written
 * directly in the packed operand layouts of wu_pack_conv3x3_weights and w_sn is left untouched;
 * with t [cols], s [rows], part [wu_sn_parts()] scratch; u_snap / v_snap / sigma are this forward's
 * values, kept for the backward pass.  Work items {int32 tensor; int32 begin; int32 count; int32 index}:
 *   wtu_chunks : begin = first column, wu_sn_wtu_cols() columns each, index = chunk number in its tensor
 *   wv_chunks  : begin = first row, wu_sn_wv_rows() rows each, index = chunk number in its tensor
 *   elem_chunks: begin / count in units of 1024 elements.
 * Backward of W_sn = W / sigma(W) with u, v constant:  dw = g / sigma - (<g, w> / sigma^2) u v^T.
 *   dot_chunks : like elem_chunks, at most wu_sn_parts()/2 per tensor, index = chunk number
 *   bwd_chunks : like elem_chunks, index = that tensor's number of dot chunks. */
int wu_sn_parts(void);
int wu_sn_wtu_cols(void);
int wu_sn_wv_rows(void);
int wu_sn_forward(const void* tensors, int n_tensors, const void* wtu_chunks, int n_wtu_chunks,
                  const void* wv_chunks, int n_wv_chunks, const void* elem_chunks, int n_elem_chunks,
                  int training, float eps, wu_stream_t stream);
int wu_sn_backward(const void* tensors, const void* dot_chunks, int n_dot_chunks,
                   const void* bwd_chunks, int n_bwd_chunks, wu_stream_t stream);

/* ---- multi-tensor Adam that also emits the packed bf16 weights (t_cls_train.py:184-185,273,311;
 * SURVEY §8 f2 / k15) --------------------------------------------------------------------------
 * torch.optim.Adam semantics (L2 weight decay added to the gradient, bias correction, no amsgrad)
 * for a whole parameter list in one launch.  `tensors`: device array of 64-byte records
 *   {float* param; const float* grad; float* exp_avg; float* exp_avg_sq; int64 numel;
 *    bf16* w_fprop; bf16* w_dgrad; int32 cout; int32 cin}
 * w_fprop == NULL: ordinary tensor.  Otherwise `param` is a 3x3 convolution weight [cout][cin][3][3]
 * (cout %% 16 == 0, cin %% 64 == 0) and the launch rewrites its operand layouts (see
 * wu_pack_conv3x3_weights; w_dgrad may be NULL) from the updated master, so the next
 * forward(x, c) (t_cls_train.py:302,242) runs on the weights g_opt.step() (:273) just produced
 * without a separate pack pass.
 * `chunks`: device array of {int32 tensor; int32 count; int64 start} (16 bytes), one CTA each; for a
 * packed tensor count = 0 and start = tile index over (cout/16) x (cin/64) tiles (row-major, see
 * wu_adam_pack_tile).
 * `step` is the 1-based step count of this update when step_state == NULL.  With step_state (device,
 * 16 bytes: {int32 step; float bc1; float rsqrt_bc2; int32 pad}, zero-initialised by the caller) the
 * count lives on the device and is advanced by the call itself: the update can then be replayed
 * from a CUDA graph. */
int wu_adam_multi(const void* tensors, const void* chunks, int n_chunks, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int step, void* step_state,
                  wu_stream_t stream);
/* Tile of a packed weight one chunk covers: *co output channels x *ci input channels (x 9 taps). */
int wu_adam_pack_tile(int* co, int* ci);

/* ---- per-sample L1 distance (t_cls_train.py:255,259-266; ops.py:22-24) ------------
 * d[s] = mean_i |a[s][i] - b[s][i]| over the n elements of sample s (fp32, contiguous), one pass; the
 * generator loss terms g_loss_l1 = mean_s d[s] and loss_con = mean_s d[s] / (lambda_s + eps) follow
 * from it.  Backward: ga[s][i] = sign(a - b) * gd[s] / n.  n %% 4 == 0, 16-byte aligned pointers. */
size_t wu_l1_per_sample_workspace_bytes(int B);
int wu_l1_per_sample_fwd(const float* a, const float* b, float* d, int B, long long n, void* workspace,
                         size_t workspace_bytes, wu_stream_t stream);
int wu_l1_per_sample_bwd(const float* a, const float* b, const float* gd, float* ga, int B, long long n,
                         wu_stream_t stream);

/* ---- layout helpers (tests, interop with NCHW fp32 PyTorch tensors) ---------------------------*/
int wu_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int B, int C, int H, int W,
                             wu_stream_t stream);
int wu_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int B, int C, int H, int W,
                             wu_stream_t stream);

/* Number of kernels this library has launched in the calling process (all threads). */
unsigned long long wu_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* WU_B200_H_ */
