/*
 * hgk.h -- C-ABI of libhgk.so: the B200 (sm_100a) kernels behind the stacked-hourglass
 * (+ASN agent) training path of zhiqiangdon/pose-adv-aug.
 *
 * The reference has no FFI/plugin layer of its own (it is pure Python on torch 0.3); the
 * drop-in boundary is the Python module API of models/asn_stacked_hg.py and
 * pylib/Criterion.py.  Below that API every torch-0.3 op the
 * reference executes on the hot path is replaced by one entry point of this header; each
 * entry cites the reference op it replaces (paths relative to the reference repo).
 *
 * Conventions (all entry points):
 *   - plain device pointers + sizes, no torch types; all activations are fp32 **NHWC**
 *     unless the name says nchw; the caller (torch) owns, allocates and frees every buffer;
 *   - asynchronous on `stream` (a cudaStream_t passed as void*), no internal syncs, no
 *     allocation, no hidden global state; thread-safe for distinct streams;
 *   - returns 0 on success, <0 on error (HGK_EINVAL bad argument, HGK_ECUDA launch error);
 *     hgk_last_error() gives the message of the calling thread's last error;
 *   - a "virtual activation" (z, scale, shift, relu) denotes the tensor
 *         relu ? max(0, z*scale[c]+shift[c]) : z*scale[c]+shift[c]        (scale==NULL: z)
 *     i.e. BatchNorm(+ReLU) of the stored pre-BN tensor z applied on load, never materialised.
 */
#ifndef HGK_H_
#define HGK_H_

#ifdef __cplusplus
extern "C" {
#endif

#define HGK_OK 0
#define HGK_EINVAL (-1)
#define HGK_ECUDA (-2)

const char* hgk_last_error(void);
int hgk_version(void);
/* 1 if the library was compiled for sm_100a and the current device is compute capability 10.x */
int hgk_device_ok(void);
/* Programmatic dependent launch: the NEXT launch of this thread may start its prologue under the tail of the kernel in front
 * of it in its stream (cudaLaunchAttributeProgrammaticStreamSerialization; every libhgk kernel on the chain waits with
 * griddepcontrol.wait before its first global access).  Only legal when that predecessor is a kernel launch: the caller arms
 * it per launch; unarmed launches keep plain stream order. */
int hgk_pdl_arm(int on);

/* ---- convolution (nn.Conv2d 1x1 / 3x3 p1, models/asn_stacked_hg.py:17,20,23,242-248,279) ----
 * y = [accumulate ? y : 0] + conv(T(x), w) + bias + T_res(res)
 * x: [N,H,W,Cin] virtual activation;  w: [ksize*ksize][Cin][Cout] ("[tap][K][N]", tap = kh*3+kw);
 * flip!=0 uses w[taps-1-tap] (the data-gradient of a 3x3 conv is the same kernel on dz with
 * flipped taps and the [tap][Cout_fwd][Cin_fwd] weight array);  bias/res optional (NULL);
 * stat_sum/stat_sq optional fp64 [Cout] accumulators (+=) of y and y^2 over all pixels: the
 * batch statistics of the following nn.BatchNorm2d (:19,22,25,243), fused into the epilogue.
 * This entry point is the exact-fp32 SIMT kernel (any C % 4 == 0); path must be 0 or 1.       */
int hgk_conv_nhwc(const float* x, const float* x_scale, const float* x_shift, int x_relu,
                  int N, int H, int W, int Cin,
                  const float* w, int ksize, int flip, const float* bias, int Cout,
                  const float* res, const float* res_scale, const float* res_shift, int res_relu,
                  float* y, int accumulate, double* stat_sum, double* stat_sq,
                  int path, void* stream);

/* The same convolution on the tcgen05 tensor cores (TF32 operands, fp32 accumulate in TMEM).
 * Weights come pre-packed by hgk_pack_weights_tc in the UMMA operand layout; w_lo != NULL selects the
 * error-compensated 3xTF32 product (fp32-class accuracy, used for the forward pass), w_lo == NULL
 * plain TF32 (used for data gradients; taps are pre-flipped by the packer).  Shapes:
 * hgk_conv_tc_supported(Cin, Cout, ksize) -> Cin % 32 == 0, Cout in {64,128,256}, ksize in {1,3}. */
int hgk_conv_tc_supported(int Cin, int Cout, int ksize);
int hgk_conv_tc_nhwc(const float* x, const float* x_scale, const float* x_shift, int x_relu,
                     int N, int H, int W, int Cin,
                     const float* w_hi, const float* w_lo, int ksize, const float* bias, int Cout,
                     const float* res, const float* res_scale, const float* res_shift, int res_relu,
                     float* y, int accumulate, double* stat_sum, double* stat_sq, void* stream);
/* Data gradient of a convolution whose INPUT was relu(bn(bz)), with the reduction pass of that BatchNorm's
 * backward fused into the epilogue:  dy = [accumulate ? dy : 0] + conv^T(dz) + extra;
 * sum_g += sum_p g, sum_gx += sum_p g*xhat,  g = dy*[bz*bscale+bshift > 0], xhat = (bz-bmean)*binvstd.
 * Only valid when this launch produces the COMPLETE gradient of that activation (single consumer). */
int hgk_conv_tc_dgrad_bnstats_nhwc(const float* dz, int N, int H, int W, int Cin,
                                   const float* w_hi, const float* w_lo, int ksize, int Cout,
                                   const float* extra, float* dy, int accumulate,
                                   const float* bz, const float* bscale, const float* bshift, int brelu,
                                   const float* bmean, const float* binvstd,
                                   double* sum_g, double* sum_gx, void* stream);
/* hgk_conv_tc_nhwc + the nn.BatchNorm2d finaliser of its statistics (hgk_bn_finalize) executed by the LAST CTA of the
 * convolution (ticket counter, re-armed by the kernel: `ticket` must be zero before the first launch).  Requires
 * stat_sum / stat_sq; count = N*H*W.  Replaces nn.Conv2d + the batch-statistics half of nn.BatchNorm2d
 * (models/asn_stacked_hg.py:36-47) in one launch. */
int hgk_conv_tc_bn_nhwc(const float* x, const float* x_scale, const float* x_shift, int x_relu,
                        int N, int H, int W, int Cin,
                        const float* w_hi, const float* w_lo, int ksize, const float* bias, int Cout,
                        const float* res, const float* res_scale, const float* res_shift, int res_relu,
                        float* y, int accumulate, double* stat_sum, double* stat_sq,
                        const float* gamma, const float* beta, float eps, float momentum,
                        float* running_mean, float* running_var, float* scale, float* shift,
                        float* save_mean, float* save_invstd, unsigned int* ticket, void* stream);
/* hgk_conv_tc_dgrad_bnstats_nhwc + hgk_bn_bwd_finalize of that BatchNorm executed by the last CTA. */
/* Forward 3x3 convolution with TF32 + 2xBF16 products (fp32-class: x*w ~= xh*wh [tf32] + bf16(xl)*bf16(wh) + bf16(xh)*bf16(wl),
 * ~3 * 2^-20 per product; 4 instead of 6 tensor-core instructions per 16 channels and tap).  Same contract as
 * hgk_conv_tc_bn_nhwc / hgk_conv_tc_nhwc; w_x2 is the "lo" buffer written by hgk_pack_weights_tc in mode 2.  Covered shapes:
 * hgk_conv_tc_x2_supported (3x3 with H and W multiples of 16 and 64 / 128 output channels on the image-tile kernel; 1x1 with
 * 128 / 256 output channels on the persistent kernel; never a split-K layer). */
int hgk_conv_tc_x2_supported(int N, int H, int W, int Cin, int Cout, int ksize);
int hgk_conv_tc_x2_nhwc(const float* x, const float* x_scale, const float* x_shift, int x_relu,
                        int N, int H, int W, int Cin,
                        const float* w_hi, const float* w_x2, int ksize, const float* bias, int Cout,
                        const float* res, const float* res_scale, const float* res_shift, int res_relu,
                        float* y, int accumulate, double* stat_sum, double* stat_sq, void* stream);
int hgk_conv_tc_bn_x2_nhwc(const float* x, const float* x_scale, const float* x_shift, int x_relu,
                           int N, int H, int W, int Cin,
                           const float* w_hi, const float* w_x2, int ksize, const float* bias, int Cout,
                           const float* res, const float* res_scale, const float* res_shift, int res_relu,
                           float* y, int accumulate, double* stat_sum, double* stat_sq,
                           const float* gamma, const float* beta, float eps, float momentum,
                           float* running_mean, float* running_var, float* scale, float* shift,
                           float* save_mean, float* save_invstd, unsigned int* ticket, void* stream);
int hgk_conv_tc_dgrad_bnfin_nhwc(const float* dz, int N, int H, int W, int Cin,
                                 const float* w_hi, const float* w_lo, int ksize, int Cout,
                                 const float* extra, float* dy, int accumulate,
                                 const float* bz, const float* bscale, const float* bshift, int brelu,
                                 const float* bmean, const float* binvstd,
                                 double* sum_g, double* sum_gx,
                                 const float* gamma, int training, float* dgamma, float* dbeta,
                                 float* cA, float* cB, float* cC, unsigned int* ticket, void* stream);
/* Data gradient of a convolution whose OUTPUT was followed by BatchNorm(+ReLU), with that BatchNorm's backward "apply"
 * (hgk_bn_bwd_apply) evaluated on load by the image-tile kernel:  g = dL/d relu(bn(gz));
 * dz = gcA*((g*[gz*gscale+gshift > 0] - gcC) - (gz - gmean)*gcB) is the convolution operand and is also written once to
 * dz_out (for the weight-gradient kernel / the shortcut).  dy = [accumulate ? dy : 0] + conv^T(dz) + extra.
 * bz != NULL additionally fuses the reduction + finaliser of the INPUT's BatchNorm exactly as
 * hgk_conv_tc_dgrad_bnfin_nhwc does.  Shapes: hgk_conv_tc_bnapply_supported(N,H,W,Cin,Cout,ksize) (H, W multiples of 16). */
int hgk_conv_tc_bnapply_supported(int N, int H, int W, int Cin, int Cout, int ksize);
int hgk_conv_tc_dgrad_bnapply_nhwc(const float* g, const float* gz, const float* gscale, const float* gshift, int grelu,
                                   const float* gmean, const float* gcA, const float* gcB, const float* gcC,
                                   float* dz_out, int N, int H, int W, int Cin,
                                   const float* w_hi, int ksize, int Cout,
                                   const float* extra, float* dy, int accumulate,
                                   const float* bz, const float* bscale, const float* bshift, int brelu,
                                   const float* bmean, const float* binvstd, double* sum_g, double* sum_gx,
                                   const float* gamma, int training, float* dgamma, float* dbeta,
                                   float* cA, float* cB, float* cC, unsigned int* ticket, void* stream);
/* table: n_entries x 8 int64 {src_off, dst_hi_off, dst_lo_off (-1: none), N, K, taps, mode, BN};
 * mode 0 (forward operand): B[n][k;tap] = W[o=n][i=k][tap];  mode 1 (data-gradient operand):
 * B[n][k;tap] = W[o=k][i=n][taps-1-tap].  Destination: [n-tile][tap][k/32] blocks of [8][BN][4] floats,
 * hi = tf32(w), lo = w - hi. */
int hgk_pack_weights_tc(const float* src_base, float* dst_base, const long long* table, int n_entries,
                        void* stream);

/* weight / bias gradient of the same convolutions (autograd of nn.Conv2d):
 * dw[co*s_co + ci*s_ci + tap*s_tap] += sum_p dz[p,co] * T(x)[p+off(tap),ci];  dbias[co] += sum_p dz[p,co]
 * (strides in elements, so gradients land directly in the OIHW-shaped .grad of the parameter). */
int hgk_conv_wgrad_nhwc(const float* x, const float* x_scale, const float* x_shift, int x_relu,
                        int N, int H, int W, int Cin, const float* dz, int Cout, int ksize,
                        float* dw, long long s_co, long long s_ci, long long s_tap,
                        float* dbias, void* stream);

/* Weight gradient on the tcgen05 tensor cores (plain TF32 operands, fp32 accumulate): accumulates (+=, vector
 * atomics) into a TAP-MAJOR destination dw[tap][Cout][Cin] -- for 1x1 convolutions that is the OIHW .grad itself;
 * 3x3 gradients go to a scratch buffer that hgk_unpack_add_grads adds into the OIHW .grad views
 * (table: n_entries x 5 int64 {src_off, dst_off, O, I, taps}).  dbias[co] += sum_p dz[p,co] (optional).
 * hgk_conv_wgrad_tc_supported(Cin, Cout, ksize): Cin in {64,128,256}, Cout % 4 == 0, ksize in {1,3}. */
int hgk_conv_wgrad_tc_supported(int Cin, int Cout, int ksize);
int hgk_conv_wgrad_tc_nhwc(const float* x, const float* x_scale, const float* x_shift, int x_relu,
                           int N, int H, int W, int Cin, const float* dz, int Cout, int ksize,
                           float* dw_tap_major, float* dbias, void* stream);
int hgk_unpack_add_grads(const float* src_base, float* dst_base, const long long* table, int n_entries,
                         void* stream);

/* repack many OIHW conv weights in one launch.  table: n_entries x 6 int64 on the device:
 * {src_offset, dst_offset, O, I, taps, mode}; mode 0: dst[tap][i][o] = src[o][i][tap],
 * mode 1: dst[tap][o][i] = src[o][i][tap].  Offsets in elements from src_base / dst_base.    */
int hgk_pack_weights(const float* src_base, float* dst_base, const long long* table, int n_entries,
                     void* stream);

/* ---- stem: nn.Conv2d(3,64,7,stride=2,padding=3), models/asn_stacked_hg.py:223,283 ----
 * img: NCHW [N,3,H,W] (the reference's input layout); w: OIHW [Cout,3,7,7]; y: NHWC [N,H/2,W/2,Cout] */
int hgk_stem_conv7_fwd(const float* img, int N, int H, int W, const float* w, const float* bias, int Cout,
                       float* y, double* stat_sum, double* stat_sq, void* stream);
int hgk_stem_conv7_wgrad(const float* img, int N, int H, int W, const float* dz, int Cout,
                         float* dw, float* dbias, void* stream);
/* Tensor-core form of the same forward convolution: space-to-depth.  xs [N,H/2,W/2,16] (NHWC) holds
 * xs[sy][sx][(dy*2+dx)*3 + c] = img[c][2 sy + dy][2 sx + dx] (channels 12..15 zero), ws [Cout,32,4,4] (OIHW; channels 16..31
 * zero: the weight packer works in blocks of 32) the matching rearrangement of w; the stem is then hgk_conv_tc_bn_nhwc /
 * hgk_conv_tc_nhwc (or their _x2 forms) over xs with the packed ws and ksize = 4 (4x4 taps at offsets -2 .. +1; Cin = 16,
 * Cout = 64, H/2 and W/2 multiples of 16). */
int hgk_stem_s2d_image(const float* img, int N, int H, int W, float* xs, void* stream);
int hgk_stem_s2d_weight(const float* w, int Cout, float* ws, void* stream);
/* The same with bn1's BatchNorm-backward apply evaluated on load: g = dL/d relu(bn1(z)), z = the stem's pre-BN output;
 * dz = cA*((g*[z*scale+shift > 0] - cC) - (z - mean)*cB) is formed in shared memory (the expression of hgk_bn_bwd_apply) and
 * never written to global memory. */
int hgk_stem_conv7_wgrad_bnapply(const float* img, int N, int H, int W, const float* g, const float* z,
                                 const float* scale, const float* shift, int relu, const float* mean,
                                 const float* cA, const float* cB, const float* cC, int Cout,
                                 float* dw, float* dbias, void* stream);

/* ---- nn.BatchNorm2d (eps 1e-5, momentum 0.1), models/asn_stacked_hg.py:19,22,25,224,243 ----
 * train: (sum, sumsq, count) -> scale = gamma*invstd, shift = beta - mean*scale, saved mean/invstd,
 * running stats update (unbiased variance).  eval: scale/shift from the running statistics.       */
int hgk_bn_finalize(const double* sum, const double* sq, long long count, const float* gamma, const float* beta,
                    float eps, float momentum, float* running_mean, float* running_var,
                    float* scale, float* shift, float* save_mean, float* save_invstd, int C, void* stream);
int hgk_bn_eval_prepare(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                        float eps, float* scale, float* shift, float* save_mean, float* save_invstd, int C,
                        void* stream);
/* backward of y = relu?(bn(z)): g = dy * [z*scale+shift > 0];  sum_g += sum g;  sum_gx += sum g*xhat */
int hgk_bn_bwd_reduce(const float* dy, const float* z, const float* scale, const float* shift, int relu,
                      const float* mean, const float* invstd, long long P, int C,
                      double* sum_g, double* sum_gx, void* stream);
/* hgk_bn_bwd_reduce + hgk_bn_bwd_finalize executed by the last CTA of the reduction (`ticket` zero before the first launch) */
int hgk_bn_bwd_reduce_fin(const float* dy, const float* z, const float* scale, const float* shift, int relu,
                          const float* mean, const float* invstd, long long P, int C,
                          double* sum_g, double* sum_gx, const float* gamma, int training, float* dgamma, float* dbeta,
                          float* cA, float* cB, float* cC, unsigned int* ticket, void* stream);
/* dgamma += sum_gx, dbeta += sum_g; coefficients of dz = cA*(g - cC - (z-mean)*cB)  (training: full BN
 * backward, cA = gamma*invstd, cB = mean(g*xhat)*invstd, cC = mean(g); eval: cB = cC = 0) */
int hgk_bn_bwd_finalize(const double* sum_g, const double* sum_gx, long long count, const float* gamma,
                        const float* mean, const float* invstd, int training, float* dgamma, float* dbeta,
                        float* cA, float* cB, float* cC, int C, void* stream);
/* in place: dy <- dz = cA*((dy*mask) - cC - (z-mean)*cB) */
int hgk_bn_bwd_apply(float* dy, const float* z, const float* scale, const float* shift, int relu,
                     const float* mean, const float* cA, const float* cB, const float* cC, long long P, int C,
                     void* stream);

/* ---- nn.MaxPool2d(2,2) (:69,227,371) on a virtual activation; backward recomputes the argmax ---- */
int hgk_maxpool2_fwd(const float* x, const float* x_scale, const float* x_shift, int x_relu,
                     int N, int H, int W, int C, float* y, void* stream);
int hgk_maxpool2_bwd(const float* x, const float* x_scale, const float* x_shift, int x_relu,
                     int N, int H, int W, int C, const float* dy, float* dx, int accumulate, void* stream);
/* ---- nn.Upsample(scale_factor=2) nearest + skip add (:70,193-203); a_up=0: plain add (:408-417) ----
 * y[N,H,W,C] = Ta(a[up ? (h/2,w/2) : (h,w)]) + Tb(b)                                                   */
int hgk_add_fwd(const float* a, const float* a_scale, const float* a_shift, int a_relu, int a_up,
                const float* b, const float* b_scale, const float* b_shift, int b_relu,
                int N, int H, int W, int C, float* y, void* stream);
/* da[N,H/2,W/2,C] = [accumulate ? da : 0] + 2x2 window sums of dy[N,H,W,C] */
int hgk_upsample2_bwd(const float* dy, int N, int H, int W, int C, float* da, int accumulate, void* stream);
/* The same two kernels writing the LAST contribution to dL/d relu(bn(z)) of a tensor with several consumers (hourglass level
 * inputs: skip branch + pool; up-path sums): the BatchNorm-backward reduction (sum g, sum g*xhat; hgk_bn_bwd_reduce_fin) and its
 * last-CTA finaliser ride on the launch instead of a pass of their own over g and z.  maxpool: the reduction is over the pooled
 * tensor x itself (its own scale/shift/relu); upsample: over z [N,H/2,W/2,C] next to da.  C/4 must divide 256. */
int hgk_maxpool2_bwd_bnred(const float* x, const float* x_scale, const float* x_shift, int x_relu, int N, int H, int W, int C,
                           const float* dy, float* dx, int accumulate, const float* mean, const float* invstd,
                           double* sum_g, double* sum_gx, const float* gamma, int training, float* dgamma, float* dbeta,
                           float* cA, float* cB, float* cC, unsigned int* ticket, void* stream);
int hgk_upsample2_bwd_bnred(const float* dy, int N, int H, int W, int C, float* da, int accumulate, const float* z,
                            const float* scale, const float* shift, int relu, const float* mean, const float* invstd,
                            double* sum_g, double* sum_gx, const float* gamma, int training, float* dgamma, float* dbeta,
                            float* cA, float* cB, float* cC, unsigned int* ticket, void* stream);
/* dst = [accumulate ? dst : 0] + src  (gradient fan-in of `x + y + tmp_in`, :334) */
int hgk_add_into(const float* src, float* dst, long long n, int accumulate, void* stream);

/* ---- inter-stack head `x = x + forth_conv(y) + in_conv(out_conv(y))` (models/asn_stacked_hg.py:329-334) ----
 * in_conv(out_conv(y)) is linear in y: the J->C and the C->C 1x1 convolutions are folded into ONE C->C convolution with
 * Wc = Wf + Wi Wo, bc = bf + bi + Wi bo (OIHW fp32, Wf [C,C], Wi [C,J], Wo [J,C]); the pair below is the weight-space part
 * (forward combination; exact chain rule back: dWf += dWc, dWi += dWc Wo^T, dWo += Wi^T dWc, dbf/dbi += dbc, dbo += Wi^T dbc). */
int hgk_head_combine_fwd(const float* w_forth, const float* b_forth, const float* w_in, const float* b_in,
                         const float* w_out, const float* b_out, float* w_comb, float* b_comb, int C, int J, void* stream);
int hgk_head_combine_bwd(const float* dw_comb, const float* db_comb, const float* w_in, const float* w_out,
                         float* dw_forth, float* db_forth, float* dw_in, float* db_in, float* dw_out, float* db_out,
                         int C, int J, void* stream);

/* ---- layout at the module boundary (the reference API is NCHW) ---- */
int hgk_nchw_to_nhwc(const float* x, int N, int C, int H, int W, float* y, void* stream);
int hgk_nhwc_to_nchw(const float* x, const float* x_scale, const float* x_shift, int x_relu,
                     int N, int H, int W, int C, float* y, void* stream);

/* ---- ASN head: nn.AvgPool2d(k) + nn.Linear (:375-377,431-435) ---- */
int hgk_avgpool_fwd(const float* x, const float* x_scale, const float* x_shift, int x_relu,
                    int N, int H, int W, int C, int k, float* y, void* stream);
int hgk_avgpool_bwd(const float* dy, int N, int H, int W, int C, int k, float* dx, int accumulate, void* stream);
int hgk_linear_fwd(const float* x, const float* w, const float* b, int M, int K, int Nout, float* y, void* stream);
int hgk_linear_bwd(const float* x, const float* w, const float* dy, int M, int K, int Nout,
                   float* dx, float* dw, float* db, void* stream);

/* ---- losses ----
 * inline MSE of stack-hg.py:156-159: loss += sum((o-t)^2)*inv_numel; dout = [acc? dout:0] + gscale*2*(o-t)*inv_numel */
int hgk_mse_fwd_bwd(const float* out, const float* target, long long n, float inv_numel, float gscale,
                    float* dout, int accumulate, double* loss, void* stream);
/* pylib/Criterion.py:12-18 and :4-10; kind 0 = weighted_L2, 1 = weighted_sigmoid_crossentropy.
 * fwd: loss(double, +=);  bwd: dpred = gout[0] * dloss/dpred (gout: device scalar)             */
int hgk_criterion_fwd(int kind, const float* pred, const float* gt, const float* weight, long long n,
                      double* loss, void* stream);
int hgk_criterion_bwd(int kind, const float* pred, const float* gt, const float* weight, long long n,
                      const float* gout, float* dpred, void* stream);

/* ---- torch.optim.RMSprop(alpha,eps,momentum=0,weight_decay=0), stack-hg.py:51-52,165 ----
 * one launch over the flat buffers: g' = g*grad_scale; v = alpha v + (1-alpha) g'^2; p -= lr g'/(sqrt(v)+eps) */
int hgk_rmsprop_flat(float* p, const float* g, float* v, long long n, float lr, float alpha, float eps,
                     float grad_scale, void* stream);
/* The same update with (lr, alpha, eps, grad_scale) read from a 4-float DEVICE array at execution time, so a captured CUDA
 * graph follows the reference's learning-rate schedule (adjust_lr, utils/util.py:105; stack-hg.py:106) without re-capture. */
int hgk_rmsprop_flat_dev(float* p, const float* g, float* v, long long n, const float* hyper, void* stream);
/* y[i] = (float)x[i]  (double loss accumulators -> fp32 scalars) */
int hgk_f64_to_f32(const double* x, float* y, int n, float mul, void* stream);

/* ==== callers either side of the training path (row a8) ==== */

/* ---- pylib/Evaluation.py: get_preds (:6-23) and final_preds (:169-193) ----
 * scores: NCHW [N,J,H,W].  preds [N,J,2] = 1-based (x, y) of the first maximum of every map, (0,0) where max <= 0
 * (row = floor(idx / H) + 1 exactly as the reference, which divides by size(2)).  mode 1 adds final_preds' quarter-
 * pixel shift (needs 1 < x < res0, 1 < y < res1), + 0.5, and -- when tinv != NULL -- the inverse crop transform:
 * tinv[n] is the row-major 2x3 fp64 top of inv(GetTransform(center, scale, rot, res0, 200)) computed by the host
 * exactly as the reference does; result = trunc(tinv * (x-1, y-1, 1)) + 1.  maxval [N,J] optional.                  */
int hgk_heatmap_peaks(const float* scores, int N, int J, int H, int W, int mode, int res0, int res1,
                      const double* tinv, float* preds, float* maxval, void* stream);
/* mode 2 of hgk_heatmap_peaks = HumanPts.heatmap2pts (pylib/HumanPts.py:118-137): (idx % W, floor(idx / W) + 0.5), masked.
 * ---- HumanPts.pts2heatmap + draw_gaussian (pylib/HumanPts.py:36-48,82-116): ground-truth heat-map rendering ----
 * pts [M,2] (x, y); heatmap [M,H,W] = zeros + the size x size blob g (host-computed as the reference does) pasted at
 * (int(x - size/2), int(y - size/2)), clipped; points outside (0,W] x (0,H] give a zero map and a zero valid_pts row. */
int hgk_pts2heatmap(const float* pts, int M, int H, int W, const float* g, int size, float* heatmap, float* valid_pts,
                    void* stream);
/* ---- calc_dists (:25-39) + dist_acc (:41-54) + the averaging loop of accuracy / accuracy_origin_res (:56-104) ----
 * dists [J,N]: |preds - target|_2 / normalize[n] where both target coordinates > boundary, else -1;
 * acc [n_idx+1] (optional): acc[k+1] = share of valid dists[idxs[k]] <= thr (-1 if none valid), acc[0] = their mean. */
int hgk_pck_accuracy(const float* preds, const float* target, const float* normalize, int N, int J,
                     float boundary, float thr, const int* idxs, int n_idx, float* dists, float* acc, void* stream);
/* dist_acc (:41-54) of a flat vector: out[0] = share of entries != -1 that are <= thr (-1 if none) */
int hgk_dist_acc(const float* dists, int n, float thr, float* out, void* stream);
/* per_person_pckh (:106-167): acc_vec[n] over the joints idxs; gt_preds = get_preds of the ground-truth heat-maps */
int hgk_per_person_pckh(const float* dists, const float* gt_preds, int N, int J, const int* idxs, int n_idx,
                        float thr, float* acc_vec, void* stream);
/* ---- flip test (stack-hg.py:225-232; pylib/HumanAug.py:179-210) ----
 * f(b)[n,c,h,w] = b[n, perm[c], h, flip_w ? W-1-w : w], perm = the n_pairs (i1,i2) channel swaps of the HOST array
 * `pairs` applied in order;  out = a ? (a + f(b)) / 2 : f(b).  NCHW, W % 4 == 0, C <= 32, out != b.                 */
int hgk_flip_merge_nchw(const float* a, const float* b_flipped, int N, int C, int H, int W, const int* pairs,
                        int n_pairs, int flip_w, float* out, void* stream);
/* ---- agent sampling (joint-train-pose-s-r-agent.py:252-271,344-363) ----
 * probs [N,K] = softmax(logits); index[n] = np.random.choice(K, 1, p=probs[n]) for the uniform u[n] in [0,1)
 * (first k whose normalised fp64 running sum exceeds u).  probs or index may be NULL.                                */
int hgk_softmax_sample(const float* logits, int N, int K, const double* u, float* probs, long long* index,
                       void* stream);
/* ---- ASN dropout (models/asn_stacked_hg.py:79-100): y = T(x) * nearest_upsample(mask [N,MH,MW]) ; x, y NHWC ---- */
int hgk_mask_mul_fwd(const float* x, const float* x_scale, const float* x_shift, int x_relu, const float* mask,
                     int N, int H, int W, int C, int MH, int MW, float* y, void* stream);
int hgk_mask_mul_bwd(const float* g, const float* mask, int N, int H, int W, int C, int MH, int MW, float* gx,
                     int accumulate, void* stream);

/* ---- image crop / rotate / resize augmentation (pylib/HumanAug.py:117-175 `crop`, reached by load_batch_data of
 * joint-train-pose-s-r-agent.py:425-450 through gen_img_heatmap, data/joint_train_s_r_agent.py:198-204) ----
 * Byte-exact GPU versions of the pixel operations `crop` performs through scipy.misc / PIL.  Images are H x W x 3
 * interleaved; uint8 unless said otherwise; the host (pose_adv_aug_b200/pylib/HumanAug.py) computes the crop geometry
 * exactly as the reference does with numpy and strings these launches together.
 * hgk_aug_minmax: out2 = {min, max} (doubles) of the region [y0,y1) x [x0,x1) of a float32 (is_u8 = 0) or uint8 image,
 *   together with 0 when include_zero (the crop window has zero padding): the cmin / cmax of scipy.misc.bytescale.
 *   scratch3: three device words in the idle state {0xFFFFFFFF, 0, 0}; the launch leaves them idle again (stream-ordered
 *   re-use).                                                                                                            */
int hgk_aug_minmax(const void* img, int is_u8, int H, int W, int y0, int y1, int x0, int x1, int include_zero,
                   unsigned int* scratch3, double* out2, void* stream);
/* the zero-padded crop window (`new_img`, ref :156-164) byte-scaled as scipy.misc.toimage does for a float64 array:
 * window (y, x) = src (y + oy, x + ox) inside [ny0,ny1) x [nx0,nx1), 0 elsewhere; out Hn x Wn x 3                        */
int hgk_aug_window_bytes(const void* src, int is_u8, int SH, int SW, int oy, int ox, int ny0, int ny1, int nx0, int nx1,
                         const double* minmax, int Hn, int Wn, unsigned char* out, void* stream);
/* scipy.misc.toimage of a whole float32 image (the pre-shrink of ref :131): bytescale in float32 arithmetic             */
int hgk_aug_image_bytes_f32(const float* src, int H, int W, const double* minmax, unsigned char* out, void* stream);
/* PIL Image.resize(BILINEAR) = ImagingResample, 8 bits per channel: tap count per output index (returns it), the
 * fixed-point tap table (bounds [out_size][2] = first input index, count; kk [out_size][ksize]), and the two passes:
 * tmp [in_h][out_w][3] from the in_h x in_w sub-image at (y_off, x_off) of src [SH][SW][3]; out [out_h][out_w][3] from tmp */
int hgk_aug_resample_ksize(int in_size, int out_size);
int hgk_aug_resample_coeffs(int in_size, int out_size, int ksize, int* bounds, int* kk, void* stream);
int hgk_aug_resize_h(const unsigned char* src, int SH, int SW, int y_off, int x_off, int in_h, int in_w, int out_w,
                     const int* bounds, const int* kk, int ksize, unsigned char* tmp, void* stream);
int hgk_aug_resize_v(const unsigned char* tmp, int in_h, int out_w, int out_h, const int* bounds, const int* kk, int ksize,
                     unsigned char* out, void* stream);
/* PIL Image.rotate(BILINEAR) = ImagingGenericTransform(affine, bilinear), fill 0; m6 = HOST array of the six affine
 * coefficients (output pixel centre -> input position), computed by the host as Image.rotate does                       */
int hgk_aug_rotate(const unsigned char* in, int H, int W, const double* m6, unsigned char* out, void* stream);
/* The whole crop pipeline of a BATCH in about a dozen launches (blockIdx.y = image): desc = N rows of hgk_aug_desc_fields()
 * int64 fields (struct AugDesc, csrc/warp.cu: source pointer and size, pre-shrink size, crop window, rotation flag / padding,
 * final-resize input, offsets of every intermediate into the two arenas), given both as a HOST copy (grid sizing) and a DEVICE
 * copy; mats_dev [N][9]: six rotation coefficients + three colour gains (a source flagged `chw` is the resident 3 x H x W
 * image, read W-flipped when `flip`, times its gains, clamped to [0,1]: data/joint_train_s_r_agent.py:160-168); arena_u8 / arena_i: byte and int scratch arenas the descriptors point into;
 * minmax [N][4] doubles; scratch [N][2][3] words in the idle state {0xFFFFFFFF, 0, 0}; out_stack [N][res][res][3] uint8.
 * Same bytes as the per-image entry points above. */
int hgk_aug_desc_fields(void);
int hgk_aug_crop_batch(const long long* desc_host, const long long* desc_dev, const double* mats_dev, int N, int res,
                       unsigned char* arena_u8, int* arena_i, double* minmax, unsigned int* scratch,
                       unsigned char* out_stack, void* stream);
/* utils/imutils.im_to_torch on a batch of crops: [N][res][res][3] uint8 -> [N][3][res][res] float32, / 255 per image only
 * when its maximum exceeds 1                                                                                            */
int hgk_aug_to_chw_float(const unsigned char* img, int N, int res, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HGK_H_ */
