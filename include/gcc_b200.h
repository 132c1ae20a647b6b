/* gcc_b200 C-ABI: B200 (sm_100a) kernels for the GCC cooperative-compression training step.
 *
 * The reference (SJLeo/GCC) is pure Python on torch; it has no FFI of its own.  The boundary that
 * these entry points replace is the set of ATen calls reached from Pix2PixModel.optimize_parameters
 * / optimizer_netD_arch (models/Pix2Pix.py:565-593).  Each declaration cites the reference call
 * site whose arithmetic it implements.  See INTEGRATION.md for the ctypes binding.
 *
 * Conventions: plain pointers are DEVICE pointers unless stated; activations are NHWC bf16 with a
 * physical channel count that is a multiple of 8 (extra channels are zero); `stream` is a
 * cudaStream_t passed as void*; every function returns 0 on success, non-zero on failure
 * (gcc_last_error() describes it).  No function allocates or keeps caller memory.
 *
 * Data parallel: there are no gcc_nccl_* entry points.  The gradient exchange (and, with --sync_bn, the statistics /
 * loss-mean exchange) goes through torch.distributed's NCCL process group on the flat gradient arenas that these
 * kernels accumulate into (gcc_b200/pix2pix.py::_allreduce_grads); the library itself never talks to NCCL.
 */
#ifndef GCC_B200_H
#define GCC_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- runtime ---- */
int gcc_abi_version(void);
const char* gcc_last_error(void);
int gcc_check_device(void); /* non-zero unless the current device is sm_100 */
/* call once per host thread before any other entry point: binds `device`'s primary context to the
 * calling thread for this library's (statically linked) CUDA runtime. */
int gcc_bind_thread(int device);
long long gcc_launch_count(void); /* kernels launched by this library so far (host counter) */

/* ---- dense contractions on tcgen05 tensor cores (conv_gemm.cu) ----
 * gcc_conv_gemm_bf16: y[n,oy,ox,y_coff+r] = act(bias[r] + sum_{kh,kw,c} x[n,iy,ix,c] * w[r][kh*KW+kw][c])
 *   transposed=0: iy = stride*oy + kh - pad             nn.Conv2d          (models/Pix2Pix.py:31,216,227,280-300)
 *   transposed=1: oy = stride*iy + kh - pad             nn.ConvTranspose2d (models/Pix2Pix.py:40-56,253-256)
 *   (also used for the data gradients of both, with the transposed weight pack)
 *   x: [N,H,W,Cx] bf16, w: [R][T=KH*KW][Cw] bf16, y: [N,OH,OW,Cy] bf16, bias: [R] fp32 or NULL.
 *   act: 0 none, 1 leaky-relu(slope), 2 tanh.  stride in {1,2}.
 *   w_per_image=1 (1x1 only): w is [N][R][Cw], one matrix per image (Gram-loss backward dF = F M).
 *   splitk_ws / ws_elems: optional fp32 scratch (contents undefined on entry and on return), NULL disables both uses:
 *   (a) >= N*OH*OW*round8(R) elements: split-K for layers with very few output pixels (U-Net inner levels; not with
 *   `stats`); (b) >= 148*128*256 elements: tail-wave split for layers whose tile count leaves the last wave of the
 *   148-SM persistent grid at most half full (512 tiles = 3.46 waves: the PatchGAN 1024 -> 512 data gradient) -- the
 *   K loop of those last tiles is cut in parts that run side by side, a small second kernel sums the parts and
 *   applies bias / activation / statistics.  Results of (b) are deterministic (plain stores, fixed summation order).
 *   stats: optional zeroed fp32 [2][stats_ld]; the epilogue adds per-output-channel sum / sum of squares of the
 *   stored bf16 values (the statistics nn.BatchNorm2d needs next, Pix2Pix.py:34,288).
 */
int gcc_conv_gemm_bf16(const void* x, int N, int H, int W, int Cx, const void* w, int R, int T, int Cw,
                       const float* bias, void* y, int OH, int OW, int Cy, int y_coff, int transposed, int KH,
                       int KW, int stride, int pad, int act, float slope, int w_per_image, float* splitk_ws,
                       long long ws_elems, float* stats, int stats_ld, void* stream);
/* gcc_wgrad_gemm_bf16: dw[b][r][kh*KW+kw][c] (+)= scale * sum_{n,oy,ox} p[n,oy,ox,r] * q[n,stride*oy+kh-pad,stride*ox+kw-pad,c]
 *   weight gradient of Conv2d (p = dy, q = x) and ConvTranspose2d (p = x, q = dy); with batched=1,
 *   KH=KW=1, p == q it is the per-sample Gram matrix f f^T (models/Pix2Pix.py:733-740).
 *   p: [N,OH,OW,Cp] bf16, q: [N,H,W,Cq] bf16, dw: fp32 [N if batched][R][KH*KW][C]. */
int gcc_wgrad_gemm_bf16(const void* p, int N, int OH, int OW, int Cp, const void* q, int H, int W, int Cq,
                        float* dw, int R, int C, int KH, int KW, int stride, int pad, int batched, int accumulate,
                        float scale, void* stream);
void gcc_debug_force_block_n(int bn);
/* timing experiments / A-B switches: bit0 skip conv epilogue stores, bit1 skip TMEM loads (both: garbage results),
 * bit4 wgrad 128-row tiles, bit5 per-launch GEMM trace, bit7 no image mode, bit11 no tail-wave split. */
void gcc_debug_set_flags(int f);
/* test hook: k-blocks per tile from which the conv kernel's tail-wave split is taken (0 = the default, 128) */
void gcc_debug_set_tail_min_kb(int kb);
/* Host-side launch planning, pure arithmetic (no device needed; tests/test_host_cpu.py):
 * gcc_plan_conv_tail: K parts (0 = none) the tail-wave split cuts the last base_tiles % sms tiles of a persistent conv
 * launch into; gcc_plan_wgrad_splits: split-K factor of a weight-gradient launch of base_ctas CTAs over total_pb
 * 64-pixel blocks (BN = tile columns, MT = 128-row accumulators per CTA, c8 = image mode). */
int gcc_plan_conv_tail(int base_tiles, int min_kb, int sms, int tail_min_kb);
int gcc_plan_wgrad_splits(int base_ctas, int total_pb, int BN, int MT, int c8);

/* ---- norm / gate / activation blocks (norm.cu) ----
 * One block = [BatchNorm2d | InstanceNorm2d | identity] -> [DifferentiableOP gate] -> [(Leaky)ReLU]
 * (models/Pix2Pix.py:26-35,201,272-341; models/DifferentiableOp.py:22-59).
 *   z = gamma*(x-mean)*rstd + beta ; g = mask(alpha,thr)*z ; y = act(g) ; optional y2 = act2(g) written
 *   into channels [y2_coff, y2_coff+Cp) of a wider NHWC buffer (U-Net skip concat, Pix2Pix.py:77).
 *   act/act2: 0 none, 1 leaky-relu(slope), 2 relu.  per_sample=1 -> instance norm statistics.
 *   sums: fp32 [N if per_sample else 1][2][Cp] (sum x, sum x^2); NULL sums = identity (no norm).
 *   alpha NULL = no gate.  gamma/beta NULL = 1/0.  mask = (sign(alpha-thr)+1)/2 evaluated in fp32.
 *   gate_after=1 (identity norm only): y = mask*act(z), the conv -> LeakyReLU -> gate order of the first
 *   MaskNLayerDiscriminator layer (Pix2Pix.py:320-322), whose gate gradient is sum dy*act(z). */
int gcc_norm_stats_bf16(const void* x, int N, long long HW, int Cp, int per_sample, float* sums, void* stream);
/* the same, ACCUMULATING into caller-zeroed `sums` (no memset node per call) */
int gcc_norm_stats_acc_bf16(const void* x, int N, long long HW, int Cp, int per_sample, float* sums, void* stream);
int gcc_norm_apply_bf16(const void* x, void* y, int N, long long HW, int Cp, int C, int per_sample, const float* sums,
                        const float* gamma, const float* beta, const float* alpha, float thr, float eps,
                        float* running_mean, float* running_var, float momentum, int act, float slope,
                        int gate_after, void* y2, int y2_Cp, int y2_coff, int act2, long long stat_count,
                        void* stream);
int gcc_norm_apply_eval_bf16(const void* x, void* y, int N, long long HW, int Cp, int C, const float* running_mean,
                             const float* running_var, const float* gamma, const float* beta, const float* alpha,
                             float thr, float eps, int act, float slope, void* y2, int y2_Cp, int y2_coff, int act2,
                             void* stream);
/* backward of the block: dx (bf16, may be NULL), dgamma/dbeta/dalpha (fp32 [C], may be NULL; dalpha is
 * the straight-through gate gradient sum dg*z; parameter gradients ACCUMULATE into their buffers).
 * dy / dy2 are gradient windows of y / y2.
 * Synchronised batch norm (data parallel): `stat_count` > 0 is the number of pixels behind `sums` / `red`
 * when the caller has all-reduced them over the ranks (0 = this device's N*HW); `phase` 1 runs only the reduction
 * (red = this device's sum dg, sum dg*xhat), the caller all-reduces a copy, and `phase` 2 runs only the apply pass with
 * the global `red` for dx and `red_param` (this device's sums, NULL = red) for dgamma / dbeta / dalpha; phase 0 = both
 * (the same argument `stat_count` exists on gcc_norm_apply_bf16).  `phase | 4`: `red` is pre-zeroed by the caller
 * (otherwise the reduction zeroes it with a memset node of its own). */
int gcc_norm_bwd_bf16(const void* x, int N, long long HW, int Cp, int C, int per_sample, const float* sums,
                      const float* gamma, const float* beta, const float* alpha, float thr, float eps, int act,
                      float slope, int gate_after, const void* dy, int dy_Cp, int dy_coff, const void* dy2, int dy2_Cp,
                      int dy2_coff, int act2, float* red, void* dx, float* dgamma, float* dbeta, float* dalpha,
                      long long stat_count, int phase, const float* red_param, void* stream);

/* ---- elementwise / layout (elementwise.cu) ---- */
/* set_input boundary (models/Pix2Pix.py:453-458): NCHW fp32 <-> NHWC bf16 channel windows */
int gcc_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int N, int C, long long HW, int Cp, int c_off,
                              int zero_to, void* stream);
int gcc_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int N, int C, long long HW, int Cp, int c_off,
                              int accumulate, void* stream);
/* torch.cat / its gradient split (models/Pix2Pix.py:77,467,471,516) */
int gcc_copy_channels_bf16(const void* src, int Cs, int s_off, void* dst, int Cd, int d_off, int C, long long npix,
                           int accumulate, void* stream);
/* mode 1 leaky-relu, 2 relu, 3 tanh; bwd takes the forward input (1,2) or output (3) as ref */
/* torch.cat([a, b], 1) for two 8-channel-padded images with ca + cb <= 8 logical channels (cat(real_A, fake_B),
 * models/Pix2Pix.py:467,471,516) and its backward split (da / db may be NULL); one 16-byte vector per pixel. */
int gcc_cat_small_bf16(const void* a, const void* b, void* y, int ca, int cb, long long npix, void* stream);
int gcc_split_small_bf16(const void* dy, void* da, void* db, int ca, int cb, long long npix, void* stream);
int gcc_act_fwd_bf16(const void* x, void* y, long long n, int mode, float slope, void* stream);
int gcc_act_bwd_bf16(const void* ref, const void* dy, void* dx, long long n, int mode, float slope, void* stream);
/* nn.Dropout(p) (models/Pix2Pix.py:64): counter-based mask from (*seed_dev, salt, index); apply to dy for bwd */
int gcc_dropout_bf16(const void* x, void* y, long long n, float p, const void* seed_dev, int salt, void* stream);
int gcc_add_bf16(const void* a, const void* b, void* y, long long n, void* stream);
/* fp32 [D0][T][D1] parameter -> bf16 [D0][T][D1p] and/or bf16 [D1][T][D0p] GEMM operand packs */
int gcc_pack_weight_bf16(const float* src, void* direct, void* transposed, int D0, int T, int D1, int D1p, int D0p,
                         void* stream);
int gcc_bias_grad_bf16(const void* dy, long long npix, int Cp, int c_off, int C, float* out, int accumulate,
                       void* stream);
/* nn.ReflectionPad2d (models/Pix2Pix.py:161,215,259); backward=1: x is dy on the padded grid, y is dx */
int gcc_reflect_pad_bf16(const void* x, void* y, int N, int H, int W, int Cp, int pad, int backward, void* stream);
/* depthwise 3x3 with fused ReflectionPad2d(1) (SeparableConv2d, models/Pix2Pix.py:132-145) */
int gcc_dw3x3_fwd_bf16(const void* x, const float* w, const float* bias, void* y, int N, int H, int W, int Cp, int C,
                       void* stream);
int gcc_dw3x3_bwd_bf16(const void* x, const void* dy, const float* w, void* dxp, void* dx, float* dw, float* dbias,
                       int N, int H, int W, int Cp, int C, int accumulate, void* stream);

/* k4 s2 p1 layers with <= 8 image-side channels (first conv of PatchGAN / U-Net, Pix2Pix.py:31,280,320; last
 * ConvTranspose of the U-Net, Pix2Pix.py:40): 16 taps x 8 channels become one 128-wide GEMM dimension.
 * im2col: col[n,oh,ow,(kh*4+kw)*8+c] = img[n,2oh+kh-1,2ow+kw-1,c]; img [N,H,W,8], col [N,H/2,W/2,128].
 * col2im: img[n,iy,ix,c] = act(bias[c] + sum_taps col[...]) (order 0: index tap*8+c, 1: c*16+tap, 2: tap*4+c with C <= 4);
 * act 2 = tanh.
 * unpad_wgrad: g[r][tap][c] += tmp[r][tap*8+c] for c < C. */
int gcc_im2col_k4s2_c8(const void* img, void* col, int N, int H, int W, void* stream);
int gcc_col2im_k4s2_c8(const void* col, int Ccol, int order, int C, const float* bias, int act, void* img, int N,
                       int H, int W, void* stream);
int gcc_unpad_wgrad_c8(const float* tmp, float* g, int R, int C, void* stream);

/* k4 s1 p1 Conv2d with <= 8 output channels (PatchGAN logits head, Pix2Pix.py:300,343): the 1x1 GEMM
 * ycol[pix, co*16+tap] = x[pix,:].w[co,tap,:] reads x once; fold sums the 16 shifted partials (+bias) into
 * y [N,H-1,W-1,8]; unfold builds dcol[n,iy,ix,tap*8+c] = dy[n,iy-kh+1,ix-kw+1,c] for the data/weight gradients;
 * unpad_wgrad_rows: g[c][tap][k] += tmp[tap*8+c][k]. */
int gcc_fold_k4s1_c8(const void* ycol, int Ccol, int C, const float* bias, void* y, int N, int H, int W, void* stream);
int gcc_unfold_k4s1_c8(const void* dy, void* dcol, int N, int H, int W, void* stream);
int gcc_unpad_wgrad_rows(const float* tmp, float* g, int C, int K, void* stream);

/* ImagePool.query (utils/image_pool.py:22-54) with device-resident state, so a captured CUDA graph draws fresh
 * decisions on every replay.  images: bf16 [b][elems_per_image] (the generated batch), pool: bf16
 * [pool_size][elems_per_image], state_dev: int64 [2] = {images stored so far, random counter} (zero-initialised by the
 * caller, advanced here), dec_ws: int32 [2 b] scratch, out: bf16 [b][elems_per_image] = the images the discriminator
 * sees.  Same policy as the reference: fill the pool first, then with probability 1/2 swap with a random slot. */
int gcc_image_pool_query_bf16(const void* images, void* pool, long long* state_dev, int* dec_ws, void* out, int b,
                              long long elems_per_image, int pool_size, void* stream);

/* ---- stem / head convolutions: <= 8 channels on the image side, large kernels, stride 1 ----
 * (k7 stem / head of the MobileResNet generator, models/Pix2Pix.py:216,259 and CycleGAN; k9 stem / head of the SRResNet,
 * models/SRGAN.py:150,190; VGG19's first k3 conv, models/GANLoss.py:110).
 * Stem ("row window"): x is a PRE-PADDED 8-channel NHWC image [N][Hrows][Wp][8] followed by >= 128 readable bytes (windows
 * at the end of a row run 7 pixels into the next one, against zero weights); the taps of one kernel row are contiguous there, so a tensor map whose window positions are
 * one pixel (16 bytes) apart delivers K-major [pixels][8 kw x 8 c] tiles: K = KH * ceil(KW/8) * 64 instead of KH*KW
 * k-blocks that are 7/8 zeros.  w_rowpack: bf16 [R][KH*ceil(KW/8)][64] from gcc_rowwin_weight_pack_bf16 (src = the
 * arena's direct pack [R][KH*KW][8]); the weight gradient comes back as fp32 [R][KH*ceil(KW/8)][64] and is accumulated
 * into the arena layout [R][KH*KW][Cin] by gcc_rowwin_wgrad_unpack_f32.  OH = padded rows - KH + 1, OW = Wp - KW + 1. */
int gcc_conv_rowwin_bf16(const void* x, int N, int Hrows, int Wp, const void* w_rowpack, int R, int KH, int KW,
                         const float* bias, void* y, int OH, int OW, int Cy, int act, float slope, float* stats, int stats_ld,
                         void* stream);
int gcc_wgrad_rowwin_bf16(const void* dy, int N, int OH, int OW, int Cp, const void* x, int Hrows, int Wp, float* dw, int R,
                          int KH, int KW, void* stream);
int gcc_rowwin_weight_pack_bf16(const void* src, void* out, int R, int KH, int KW, void* stream);
int gcc_rowwin_wgrad_unpack_f32(const float* tmp, float* g, int R, int KH, int KW, int Cin, void* stream);
/* Head ("fold"): C <= CG (4 or 8) output channels.  One 1x1 GEMM (gcc_conv_gemm_bf16) over the pre-padded input computes
 * ycol[pix][(kh*KW+kw)*CG + c] = x[pix,:] . w[c][kh][kw][:]; gcc_fold_taps_bf16 sums the shifted partials:
 *   y[n,oy,ox,c] = act(bias[c] + sum ycol[n, oy + dir*kh + off, ox + dir*kw + off, (kh*KW+kw)*CG + c])   (act 2 = tanh)
 * over the ycol grid [N][GH][GW][Ccol] (terms outside are skipped); dir = +1: the conv itself, dir = -1: the data gradient
 * of a stem conv on its padded input grid.  gcc_unfold_taps_bf16 builds dcol[n,gy,gx,(tap)*CG+c] = dy[n,gy-kh-off,gx-kw-off,c]
 * for the head's data / weight gradient GEMMs; gcc_fold_weight_pack_bf16 builds the GEMM weight operands from the arena
 * packs (mode 0: [(tap,c)][K] from direct [C][T][Kp]; mode 1: [K][(tap,c)] from transposed [K][T][Cp]; rows_p = padded
 * number of (tap,c) rows / columns);
 * gcc_fold_wgrad_unpack_f32: g[c][t][k] += tmp[(t*CG+c)][k]. */
int gcc_fold_taps_bf16(const void* ycol, int Ccol, int CG, int KH, int KW, int C, const float* bias, int act, void* y, int N,
                       int GH, int GW, int OH, int OW, int dir, int off, void* stream);
int gcc_unfold_taps_bf16(const void* dy, void* dcol, int Ccol, int CG, int KH, int KW, int N, int GH, int GW, int OH, int OW,
                         int off, void* stream);
int gcc_fold_weight_pack_bf16(const void* src, void* out, int mode, int C, int T, int CG, int K, int Kp, int Cp, int rows_p,
                              void* stream);
int gcc_fold_wgrad_unpack_f32(const float* tmp, float* g, int C, int T, int CG, int K, void* stream);
/* nn.Conv2d zero padding made explicit for the stem path (backward = 1: crop the gradient); `slack` extra zero rows per image */
int gcc_zero_pad_bf16(const void* x, void* y, int N, int H, int W, int Cp, int pad, int slack, int backward, void* stream);

/* ---- "slab" kernels of the InstanceNorm networks (slab.cu) ----
 * MobileResnetBlock (models/Pix2Pix.py:147-197) and CycleGAN's InstanceNorm PatchGAN (models/CycleGAN.py:140-178) at the
 * resolutions where one CTA owns all H*W <= 4096 pixels of 8 channels of one sample: InstanceNorm2d's statistics become a
 * block reduction and each chain is ONE launch reading its input once (gcc_slab_supported tells whether a shape fits).
 *   gcc_dw_in_slab_fwd:  z = InstanceNorm(depthwise3x3(ReflectionPad2d(1)(x)) + b)     (SeparableConv2d conv.0 + conv.1)
 *   gcc_dw_in_slab_bwd:  dx, dw += , dbias += from dz; the depthwise output is recomputed from x (saved: x, stats)
 *   gcc_in_act_slab_fwd: z = act(InstanceNorm(y)) (+ res)   act 0 none / 1 leaky-relu(slope) / 2 relu; res = block skip
 *   gcc_in_act_slab_bwd: dy from dz (the skip's own gradient is dz)
 * stats: fp32 [N][Cp][2] = (mean, rstd) per (sample, channel), written by the forward kernels for their backward. */
int gcc_slab_supported(int H, int W);
int gcc_dw_in_slab_fwd_bf16(const void* x, const float* w, const float* bias, void* z, float* stats, int N, int H, int W,
                            int Cp, int C, float eps, void* stream);
int gcc_dw_in_slab_bwd_bf16(const void* x, const void* dz, const float* w, const float* bias, const float* stats, void* dx,
                            float* dw, float* dbias, int N, int H, int W, int Cp, int C, void* stream);
int gcc_in_act_slab_fwd_bf16(const void* y, const void* res, void* z, float* stats, int N, long long HW, int Cp, int C,
                             float eps, int act, float slope, void* stream);
int gcc_in_act_slab_bwd_bf16(const void* y, const void* dz, const float* stats, void* dy, int N, long long HW, int Cp, int C,
                             int act, float slope, void* stream);

/* ---- loss reductions (loss.cu) ---- */
/* GANLoss (models/GANLoss.py:38-59). mode: 0 hinge, 1 lsgan, 2 vanilla, 3 wgangp.
 * kind: 0 D-real, 1 D-fake, 2 G.  out is an fp32 device scalar that the caller zeroes. */
int gcc_gan_loss_fwd_bf16(const void* pred, long long npix, int Cp, int C, int mode, int kind, float* out,
                          void* stream);
int gcc_gan_loss_bwd_bf16(const void* pred, long long npix, int Cp, int C, int mode, int kind, const float* gout,
                          void* dpred, void* stream);
/* mode 0: out += mean|a-b| (criterionL1, Pix2Pix.py:520); mode 1: out += mean (a-b)^2 (criterionMSE, :543) */
int gcc_diff_reduce_bf16(const void* a, const void* b, long long npix, int Cp, int C, int mode, float* out,
                         void* stream);
/* mode 0: da = gout*sign(a-b)/count ; mode 1 (sqrt(MSE)): da = gout*(a-b)/(count*sqrt(*msq)) ;
 * mode 2 (plain MSE, models/CycleGAN.py:513-514): da = gout*2(a-b)/count */
int gcc_diff_bwd_bf16(const void* a, const void* b, long long npix, int Cp, int C, int mode, const float* gout,
                      const float* msq, void* da, void* stream);
int gcc_sqdiff_reduce_f32(const float* a, const float* b, long long n, float* out, void* stream);
int gcc_gram_bwd_matrix(const float* gs, const float* gt, int B, int C, int Cp, float gram_scale, const float* gout,
                        const float* msq, int mse, void* m, void* stream);
int gcc_scalar_sqrt(const float* in, float* out, void* stream);

/* ---- optimizer (optim.cu) ---- */
/* torch.optim.Adam step over a flat fp32 arena (Pix2Pix.py:382-440). hyper_dev: device fp32[5] =
 * {lr, beta1, beta2, eps, step(int bits)}; the step counter is incremented on the device. */
int gcc_adam_step_f32(float* p, const float* g, float* m, float* v, long long n, float* hyper_dev, void* stream);
/* L1_sparsity (Pix2Pix.py:554-563): g += lambda * sign(w) */
int gcc_l1_sparsity_f32(const float* w, float* g, long long n, float lambda, void* stream);
/* DifferentiableOP.clip_alpha (DifferentiableOp.py:51-53) */
int gcc_clamp_f32(float* x, long long n, float lo, float hi, void* stream);
/* re-pack all conv weights of a net after an optimizer step: device int64 table [count][8] */
int gcc_pack_weights_table(const void* table_dev, int count, void* stream);

/* ---- SRGAN additions (srgan.cu) ----
 * nn.PReLU() with one learnable slope (models/SRGAN.py:49-50,89): y = x > 0 ? x : a x; bwd: dx = dy (x > 0 ? 1 : a),
 * dslope += sum_{x <= 0} dy x (dx / dslope may be NULL). */
int gcc_prelu_fwd_bf16(const void* x, void* y, long long n, const float* slope_dev, void* stream);
int gcc_prelu_bwd_bf16(const void* x, const void* dy, void* dx, long long n, const float* slope_dev, float* dslope,
                       void* stream);
/* nn.PixelShuffle(2) (models/SRGAN.py:88) on NHWC: dst[n,2h+i,2w+j,c] = src[n,h,w,4c+2i+j]; inverse = 1 is the
 * backward permutation (src = gradient wrt the shuffled tensor, dst = gradient wrt the conv output). */
int gcc_pixel_shuffle2_bf16(const void* src, void* dst, int N, int H, int W, int C, int Cin_p, int Cout_p, int inverse,
                            void* stream);
/* nn.MaxPool2d(2, 2) of torchvision's VGG19 features (models/GANLoss.py:110-134); bwd routes dy to the first
 * maximum of each window in row-major order. */
int gcc_maxpool2_fwd_bf16(const void* x, void* y, int N, int H, int W, int Cp, void* stream);
int gcc_maxpool2_bwd_bf16(const void* x, const void* dy, void* dx, int N, int H, int W, int Cp, void* stream);
/* convert_image('[-1, 1]' -> 'imagenet-norm') (data/sr_dataset.py:15-64) and its gradient: y = x * scale[c] + shift[c]
 * on an 8-channel (3 logical) NHWC image; shift_dev may be NULL. */
int gcc_channel_affine8_bf16(const void* x, void* y, long long npix, int C, const float* scale_dev,
                             const float* shift_dev, void* stream);
/* AdaptiveAvgPool2d((1,1)) + Linear(C, 1) (models/SRGAN.py:231-245): sums = per-sample channel sums as written by
 * gcc_norm_stats_bf16(per_sample = 1); logits bf16 [N][8] (channel 0, the layout the GAN-loss kernels read).
 * bwd: dlogit bf16 [N][8], dx bf16 [N,HW,Cp] (may be NULL), dw/db accumulate. */
int gcc_pool_linear_fwd(const float* sums, int N, long long HW, int Cp, int C, const float* w, const float* b,
                        void* logits, void* stream);
int gcc_pool_linear_bwd(const void* dlogit, const float* sums, const float* w, int N, long long HW, int Cp, int C,
                        void* dx, float* dw, float* db, void* stream);

/* ---- SAGAN additions (sagan.cu) ----
 * SpectralNorm._update_u_v (models/SAGAN.py:25-38): one power iteration IN PLACE on u [height], v [width] over the
 * fp32 weight w_bar [height][width] (the arena's channels-last conv weight, height = weight.shape[0]);
 * t_out [height] = w_bar v (saved for the u gradient), *sigma_out = u . (w_bar v); scratch2: 2 floats. */
int gcc_spectral_norm_fwd(const float* w_bar, float* u, float* v, int height, int width, float* t_out, float* sigma_out,
                          float* scratch2, void* stream);
/* bf16 GEMM operand packs of w_bar / sigma (the weight the spectral-normed conv uses, :38) */
int gcc_pack_weight_scaled_bf16(const float* src, const float* sigma_dev, void* direct, void* transposed, int D0, int T,
                                int D1, int D1p, int D0p, void* stream);
/* backward of w = w_bar / sigma, sigma = u . (w_bar v): dw_bar += (dw_eff - (s/sigma) u v^T) / sigma with
 * s = sum(dw_eff * w_bar); du += dsigma * t_saved, dv += dsigma * w_bar^T u, dsigma = -s / sigma^2 (du, dv may be NULL:
 * the vectors only carry a gradient after set_requires_grad(netD, True), :553-559). scratch1: 1 float. */
int gcc_spectral_norm_bwd(const float* dw_eff, const float* w_bar, const float* u, const float* v, const float* sigma,
                          const float* t_saved, int height, int width, float* dw_bar, float* du, float* dv, float* scratch1,
                          void* stream);
/* Self_Attn core (models/SAGAN.py:96-104): q, k bf16 [N][L][dp] (d logical channels), v bf16 [N][L][Cp];
 * probs bf16 [N][L][L] = softmax_j(q_i . k_j), out[n][i][c] = sum_j probs[i][j] v[j][c].
 * The two matrix products run as batched tcgen05 GEMMs (one weight matrix per image), the logits stay fp32 until the
 * softmax; backward: dprobs = dout v^T and dq = de k on the same kernel, dk = de^T q and dv = probs^T dout on the
 * batched weight-gradient kernel.  ws: device scratch of gcc_attn_workspace_bytes(N, L, dp, Cp, backward) bytes;
 * de_scratch: bf16 [N][L][L]. */
long long gcc_attn_workspace_bytes(int N, int L, int dp, int Cp, int backward);
int gcc_attn_fwd_bf16(const void* q, const void* k, const void* v, int N, int L, int d, int dp, int C, int Cp, void* probs,
                      void* out, void* ws, long long ws_bytes, void* stream);
int gcc_attn_bwd_bf16(const void* q, const void* k, const void* v, const void* probs, const void* dout, int N, int L,
                      int d, int dp, int C, int Cp, void* de_scratch, void* dq, void* dk, void* dv, void* ws,
                      long long ws_bytes, void* stream);
/* out = gamma * a + x (Self_Attn's residual, :106); bwd: da = gamma dy, dgamma += sum dy a (dx = dy) */
int gcc_scale_add_bf16(const void* a, const void* x, const float* gamma_dev, void* y, long long n, void* stream);
int gcc_scale_add_bwd_bf16(const void* dy, const void* a, const float* gamma_dev, void* da, float* dgamma, long long n,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GCC_B200_H */
