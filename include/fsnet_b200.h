/*
 * fsnet_b200 -- C ABI of the B200-native FSNet training-step hot path.
 *
 * The reference (Owen-Liuyuxuan/FSNet) has NO native interface on this path: every operation below is
 * a chain of stock PyTorch ops in Python.  Each entry point therefore cites the reference Python it
 * replaces (file:line relative to the reference checkout); INTEGRATION.md shows the ctypes binding a
 * reference maintainer would add.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host; the caller owns every buffer
 *     (the kernels never allocate);
 *   - `stream` is a cudaStream_t passed as void*; kernels are enqueued on it and the call returns
 *     without synchronising;
 *   - return value: 0 on success, a negative fsnet_status otherwise; fsnet_last_error() gives text;
 *   - the library is re-entrant; it keeps no mutable global state except the last-error string
 *     (thread-local) and lazily cached function attributes / tensor-map driver entry point.
 *   - images are fp32 NCHW as the reference's data dict delivers them;
 *     network activations are fp32/bf16 NHWC.
 */
#ifndef FSNET_B200_H
#define FSNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  FSNET_OK = 0,
  FSNET_ERR_INVALID = -1,   /* bad argument (null pointer, non-positive size, unsupported shape) */
  FSNET_ERR_CUDA = -2,      /* a CUDA runtime / driver call failed */
  FSNET_ERR_UNSUPPORTED = -3
} fsnet_status;

/* mask element types for `patched_mask` (the reference delivers fp64) */
#define FSNET_MASK_NONE 0
#define FSNET_MASK_F32 1
#define FSNET_MASK_F64 2

/* flags of the fused reprojection kernels */
#define FSNET_FLAG_OVERLAP_MASK 1u   /* overlapped_mask=True: nearest-sample patched_mask, invalid := 100 */
#define FSNET_FLAG_MOTION_MASK 2u    /* 'motion_mask' branch: min over the two reprojections only   */
#define FSNET_FLAG_PACKED_MASK 4u    /* the 4th component of every packed pixel holds patched_mask at that pixel (1 without a mask):
                                        written by fsnet_identity_photometric_masked, read by fsnet_warp_ssim_fwdbwd */

int fsnet_abi_version(void);
const char* fsnet_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * camera set-up.  Replaces the per-(scale, frame) host round trip of
 * monodepth2_decoder.py:82-90 (P2.cpu().numpy(), np.linalg.pinv, .cuda()) and Project3D's
 * P = (K @ T)[:, :3, :] (monodepth_utils.py:155).
 *   P2   [B,3,4] fp32     T0, T1 [B,4,4] fp32 (cam_T_cam for frame_ids[1], frame_ids[2])
 *   cam  [B,2,21] fp32 out: inv(K) row-major (9) followed by P row-major (12), per source frame
 * ------------------------------------------------------------------------------------------- */
int fsnet_camera_setup(const float* P2, const float* T0, const float* T1, int B, float* cam, void* stream);

/* ---------------------------------------------------------------------------------------------
 * PoseNet output -> cam_T_cam: rot_from_axisangle (Rodrigues, axis = v / (|v| + 1e-7)), get_translation_matrix and
 * transformation_from_parameters (monodepth_utils.py:298-337, 31-44, 46-63) in one launch; M = T*R, or R^T * T(-t) with invert.
 *   axisangle, translation [B,3] fp32      T [B,4,4] fp32 out
 * backward: grad_T [B,4,4] -> grad_axisangle, grad_translation [B,3]
 * ------------------------------------------------------------------------------------------- */
int fsnet_pose_matrix(const float* axisangle, const float* translation, int B, int invert, float* T, void* stream);
int fsnet_pose_matrix_bwd(const float* axisangle, const float* translation, const float* grad_T, int B, int invert,
                          float* grad_axisangle, float* grad_translation, void* stream);

/* ---------------------------------------------------------------------------------------------
 * identity photometric terms: 0.85*mean_c SSIM(src_f, tgt) + 0.15*mean_c |tgt - src_f| for both
 * source frames.  monodepth2_decoder.py:248-254 (+ :118-128, monodepth_utils.py:184-215).  They do
 * not depend on the scale, so they are computed once per step.  The same pass also emits the
 * RGBX-packed copy of the three images that the per-scale kernels gather from (one 128-bit load per
 * bilinear corner instead of three scalar ones).
 *   tgt, src0, src1 [B,3,H,W] fp32        ident [B,2,H,W] fp32 out
 *   packed [3,B,H,W,4] fp32 out (target, source 0, source 1; 16-byte aligned) or NULL
 * ------------------------------------------------------------------------------------------- */
int fsnet_identity_photometric(const float* tgt, const float* src0, const float* src1,
                               int B, int H, int W, float* ident, float* packed, void* stream);
/* Same pass; the fourth component of each packed pixel additionally carries patched_mask[b, y, x] as fp32 (1 when mask == NULL).
 * The frame-pair training kernel then gets the loss weight of a pixel with its target colour and the overlap mask of a warped
 * pixel (nearest sample of patched_mask, monodepth2_decoder.py:110-116) with the bilinear corners it loads anyway: the nearest
 * pixel IS one of the four corners.  Pass FSNET_FLAG_PACKED_MASK to fsnet_warp_ssim_fwdbwd. */
int fsnet_identity_photometric_masked(const float* tgt, const float* src0, const float* src1, const void* mask, int mask_dtype,
                                      int B, int H, int W, float* ident, float* packed, void* stream);

/* ---------------------------------------------------------------------------------------------
 * fused per-scale reprojection loss, forward.  One launch replaces, for one scale,
 * _generate_images_pred (monodepth2_decoder.py:61-116: F.interpolate, BackprojectDepth,
 * Project3D, two F.grid_sample per frame), compute_reprojection_loss for both frames (:118-128),
 * the 100.0 overwrite (:231-235), the tie-break noise and 4-way min (:257-263), the patched-mask
 * product and the two sums of :292.
 *   depth_s [B,1,hs,ws]   packed [3,B,H,W,4] from fsnet_identity_photometric   mask [B,H,W] (mask_dtype) or NULL
 *   cam     [B,2,21] from fsnet_camera_setup
 *   ident   [B,2,H,W] from fsnet_identity_photometric (ignored with FSNET_FLAG_MOTION_MASK)
 *   noise   [B,2,H,W] fp32 standard-normal draws (scaled by 1e-5 in the kernel) or NULL
 *   motion  [B,H,W] fp32 motion mask (only with FSNET_FLAG_MOTION_MASK) or NULL
 *   accum   [2] fp64, must be zeroed by the caller: accum[0] += sum(min * mask), accum[1] += sum(mask)
 *   sel     [B,H,W] uint8 arg-min index (0,1 identity; 2,3 reprojection) or NULL
 *   pred0   [2,3,H,W] fp32 warped sources of batch sample 0 (the reference's `hm` images) or NULL
 * ------------------------------------------------------------------------------------------- */
int fsnet_warp_ssim_fwd(const float* depth_s, int hs, int ws, const float* packed,
                        const void* mask, int mask_dtype, const float* cam,
                        const float* ident, const float* noise, const float* motion,
                        unsigned flags, int B, int H, int W,
                        double* accum, uint8_t* sel, float* pred0, void* stream);

/* ---------------------------------------------------------------------------------------------
 * fused per-scale reprojection loss, backward (recomputes the forward; nothing is saved).
 * Autograd of the same reference lines.  d loss / d depth_s and, if grad_P != NULL, d loss / d P
 * (the pose path of MonoDepthMeta, monodepth2_model.py:42-43).
 *   accum      the forward's [2] fp64 (accum[1] = sum(mask) is read)
 *   gout       [1] fp32 device scalar: d L / d (this scale's photometric term), i.e. the incoming
 *              gradient of the total loss divided by num_scales
 *   grad_depth [B,1,hs,ws] fp32, must be zeroed by the caller when hs != H (scatter-add)
 *   grad_P     [B,2,12] fp32, zeroed by the caller, or NULL
 * ------------------------------------------------------------------------------------------- */
int fsnet_warp_ssim_bwd(const float* depth_s, int hs, int ws, const float* packed,
                        const void* mask, int mask_dtype, const float* cam,
                        const float* ident, const float* noise, const float* motion,
                        unsigned flags, int B, int H, int W,
                        const double* accum, const float* gout,
                        float* grad_depth, float* grad_P, void* stream);

/* ---------------------------------------------------------------------------------------------
 * fused forward + backward of the per-scale reprojection loss for training steps: ONE pass computes the masked sum of
 * the per-pixel minimum (the forward value) and, by the same recomputation-free walk, d loss / d depth_s (and d loss /
 * d P).  The normaliser sum(patched_mask) of monodepth2_decoder.py:292 does not depend on the depth, so it is produced
 * up front by fsnet_mask_sum; the upstream gradient enters as `gout` (d total / d this scale's term -- the executor
 * runs the kernel with the unit gradient 1/num_scales during the forward and rescales in backward, which is exact
 * because the loss is linear in it).  Saves the separate forward launch (~100 us per scale at cfg2).
 *   fsnet_mask_sum: out[k*out_stride + 1] += sum(mask) for k < n_out (mask == NULL: += n); `out` zeroed by the caller
 *   fsnet_warp_ssim_fwdbwd: arguments as fsnet_warp_ssim_bwd (+ the MEI ray table, NULL for the pinhole camera);
 *   accum [2] fp64: accum[1] = sum(mask) on entry (read), accum[0] += sum(min * mask)
 * ------------------------------------------------------------------------------------------- */
int fsnet_mask_sum(const void* mask, int mask_dtype, long long n, double* out, int n_out, int out_stride, void* stream);
int fsnet_warp_ssim_fwdbwd(const float* lut, const int* lut_idx,
                           const float* depth_s, int hs, int ws, const float* packed,
                           const void* mask, int mask_dtype, const float* cam,
                           const float* ident, const float* noise, const float* motion,
                           unsigned flags, int B, int H, int W,
                           double* accum, const float* gout,
                           float* grad_depth, float* grad_P, void* stream);

/* ---------------------------------------------------------------------------------------------
 * MEI (unified omnidirectional) fisheye camera -- FishEyeDecoder (monodepth2_decoder.py:350-420)
 * with MeiCameraProjection (monodepth/networks/utils/mei_fisheye_utils.py).
 *
 * fsnet_mei_lut: the cached per-calibration ray table of image2cam (mei_fisheye_utils.py:139-170:
 * X=(u-u0)/gamma1, Y=(v-v0)/gamma2, Newton solve of the radial distortion :70-79, bisection of the
 * mirror equation :85-101, mask := 0 where no root or Z < 0.05, masked entries := -1, X,Y *= Z+xi),
 * built on the device in fp64 like numba evaluates it.  The reference keys its cache with values
 * pulled to the host by .item() (:151-154); here the table slots carry the calibration they were
 * built for, and a slot is rebuilt only when its calibration changed (no host synchronisation).
 *   P2      [B,3,4] fp32 (gamma1, gamma2, u0, v0 = P2[0,0], P2[1,1], P2[0,2], P2[1,2])
 *   calib   [B,3]   fp64 (xi, k1, k2 of calib_meta[b]; fp64 because numba solves with the Python doubles)
 *   header  [B,8]   fp64 persistent state, zero-initialised by the caller once
 *   lut_idx [B]     int32 out: table slot of sample b (first sample with the same calibration)
 *   lut     [B,H,W,4] fp32 persistent (X, Y, Z, mask), 16-byte aligned
 * fsnet_camera_setup_mei: cam [B,2,21] = {gamma1, gamma2, u0, v0, xi, k1, k2, 0, 0, T[:3,:4]} per frame.
 * fsnet_warp_ssim_mei_fwd / _bwd: the fused per-scale loss with points = LUT * norm, p = T [points;1],
 * (u,v) = cam2image(p) (:23-51, :379-387) and the overlap mask additionally multiplied by the LUT mask
 * (:409).  Arguments as fsnet_warp_ssim_fwd / _bwd; `norm_s` is the decoder output (the ray norm);
 * grad_T [B,2,12] is d loss / d cam_T_cam[:, :3, :4] per frame (or NULL).
 * fsnet_mei_depth: FishEyeDecoder.get_prediction (:415-420), depth = Z_lut * norm.
 * ------------------------------------------------------------------------------------------- */
int fsnet_mei_lut(const float* P2, const double* calib, int B, int H, int W,
                  double* header, int* lut_idx, float* lut, void* stream);
int fsnet_camera_setup_mei(const float* P2, const double* calib, const float* T0, const float* T1, int B,
                           float* cam, void* stream);
int fsnet_warp_ssim_mei_fwd(const float* lut, const int* lut_idx,
                            const float* norm_s, int hs, int ws, const float* packed,
                            const void* mask, int mask_dtype, const float* cam,
                            const float* ident, const float* noise, const float* motion,
                            unsigned flags, int B, int H, int W,
                            double* accum, uint8_t* sel, float* pred0, void* stream);
int fsnet_warp_ssim_mei_bwd(const float* lut, const int* lut_idx,
                            const float* norm_s, int hs, int ws, const float* packed,
                            const void* mask, int mask_dtype, const float* cam,
                            const float* ident, const float* noise, const float* motion,
                            unsigned flags, int B, int H, int W,
                            const double* accum, const float* gout,
                            float* grad_norm, float* grad_T, void* stream);
int fsnet_mei_depth(const float* norm, const float* lut, const int* lut_idx, int B, int H, int W,
                    float* depth, void* stream);

/* ---------------------------------------------------------------------------------------------
 * k x k box average of `planes` = B*C image planes [H,W] -> [H/k, W/k]: F.adaptive_avg_pool2d(original_image_0, (h, w)) of
 * monodepth2_decoder.py:219, once per step and scale; fsnet_smooth_* then take the pooled image with H = h, W = w.
 * ------------------------------------------------------------------------------------------- */
int fsnet_box_pool(const float* img, int planes, int H, int W, int k, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * edge-aware smoothness on mean-normalised disparity (monodepth2_decoder.py:214-219,294-296,
 * monodepth_utils.py:168-181), forward and backward.  `img` is original_image_0 at full resolution;
 * the 2^s x 2^s box average (adaptive_avg_pool2d) is taken inside the kernel.
 *   disp [B,1,h,w]  img [B,3,H,W] with H = h*k, W = w*k
 *   sums [B,3] fp64 workspace, zeroed by the caller: per-sample sum(disp), then backward scratch
 *   out  [1] fp64, zeroed by the caller: += weight * (mean|dx|e^-|dIx| + mean|dy|e^-|dIy|)
 *   gout [1] fp32 device scalar (d L / d smooth term); grad_disp [B,1,h,w] fp32 out (overwritten)
 * ------------------------------------------------------------------------------------------- */
int fsnet_smooth_fwd(const float* disp, const float* img, int B, int h, int w, int H, int W,
                     float weight, double* sums, double* out, void* stream);
int fsnet_smooth_bwd(const float* disp, const float* img, int B, int h, int w, int H, int W,
                     float weight, double* sums, const float* gout, float* grad_disp, void* stream);

/* ---------------------------------------------------------------------------------------------
 * softmax-over-depth-bins head (depth_encoder.py:76-88,115-121; monodepth_utils.py:19-24) and the
 * sigmoid head (depth_encoder.py:104-109), forward and backward.
 *   logits [B,n,h,w] fp32 NCHW (channel stride = h*w) or NHWC (channels_last != 0)
 *   bins [n] fp32; scale [B] fp32 (fx/base_fx) or NULL; depth, disp [B,1,h,w] out
 * ------------------------------------------------------------------------------------------- */
int fsnet_depth_head_fwd(const float* logits, const float* bins, const float* scale, int B, int n, int h, int w,
                         int channels_last, int sigmoid_head, float min_depth, float max_depth,
                         float* depth, float* disp, void* stream);
int fsnet_depth_head_bwd(const float* logits, const float* bins, const float* scale, int B, int n, int h, int w,
                         int channels_last, int sigmoid_head, float min_depth, float max_depth,
                         const float* grad_depth, const float* grad_disp, float* grad_logits, void* stream);

/* ---------------------------------------------------------------------------------------------
 * training augmentation of uint8 frames on the device: the pixel work of the reference's CPU list RandomWarpAffine ->
 * RandomMirror -> colour jitter -> Normalize (vision_base/data/augmentations/augmentations.py:91-109,200-226,377-498,527-592)
 * with OpenCV's arithmetic (cv2.warpAffine fixed-point coordinates, float HSV); the random parameters are drawn on the host.
 *   frames [B,F,H0,W0,3] uint8 (zero-padded to a common H0 x W0), mask [B,H0,W0] uint8 or NULL
 *   plan [B,16] fp64 per sample: [15] geometry 0 = affine: [0:6] inverse 2x3 matrix (output -> source), 1 = resize + zero pad:
 *        [0:4] = scale_x, scale_y, w_eff, h_eff;  [6] mirror, [7:10] colour op codes in order (0 none, 1 brightness, 2 contrast,
 *        3 saturation), [10:13] their values (NaN = HSV round trip only), [13] h0, [14] w0 (valid region of the padded source)
 *   mean_std [6] fp32;  image, original [F,B,3,H,W] fp32 out: (aug/255 - mean)/std and warped/255;  mask_out [B,H,W] fp64 or NULL
 * ------------------------------------------------------------------------------------------- */
int fsnet_augment_frames(const unsigned char* frames, const unsigned char* mask, const double* plan, int B, int F, int H0, int W0,
                         int H, int W, const float* mean_std, float* image, float* original, double* mask_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * distillation loss of the second training stage (monodepth2_decoder.py:185-203, scaled branch; the sigmoid of
 * MultiChannelDepthDecoderUncertain.forward, depth_encoder.py:190, is applied here), value + unit gradients in one pass.
 *   pred, teacher [n] fp32: the student's and the (frozen) teacher's depth map of one scale
 *   ulogit [n] fp32: un-activated uncertainty head output, or NULL (is_uncertain_distill = False: plain L1)
 *   out [1] fp64, accumulated: += mean_i( |t-p| / u + log(u + 1e-5) ),  u = sigmoid(ulogit)
 *   grad_pred, grad_ulogit [n] fp32 out or NULL: d out / d pred, d out / d ulogit;  uncertain [n] fp32 out or NULL: u
 * ------------------------------------------------------------------------------------------- */
int fsnet_distill_loss(const float* pred, const float* teacher, const float* ulogit, long long n, double* out,
                       float* grad_pred, float* grad_ulogit, float* uncertain, void* stream);

/* ---------------------------------------------------------------------------------------------
 * loss finalisation: turns the per-scale accumulators into the reference's loss_dict entries
 * (monodepth2_decoder.py:292-303).  acc [S,4] fp64 = {num, den, smooth, unused} per scale.
 *   out [2*S+2] fp64: loss/s (S), smooth_loss/s (S), total_loss, spare
 * ------------------------------------------------------------------------------------------- */
int fsnet_loss_finalize(const double* acc, int S, double* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * NHWC views.  Element (n, y, x, c) of a view lives at
 *     ((n*(h+2*ring) + y+ring) * (w+2*ring) + x+ring) * c_total + c_off + c
 * `ptr` addresses the ring origin.  "planes" views are bf16: the hi plane at `ptr`, the lo plane
 * (x - hi) right behind it (n*(h+2*ring)*(w+2*ring)*c_total elements later); fp32 views are single.
 * The ring of a planes buffer holds the replicate padding of the interior.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  void* ptr;
  int n, h, w, c;      /* logical size of the view */
  int ring;            /* 0 or 1 pixel of materialised border around h x w */
  int c_total, c_off;  /* channel stride of the underlying buffer and first channel of the view */
} fsnet_view;

/* ---------------------------------------------------------------------------------------------
 * convolution on the tcgen05 tensor cores (implicit GEMM, TMA-fed, fp32 accumulation in TMEM).
 * Replaces nn.Conv2d as used by resnet.py:21-89,119 / blocks.py:43-45 / depth_encoder.py:62 /
 * pose_decoder.py:18-21 of the reference; with transposed / flipped weights it is also the data
 * gradient of those layers.
 *   in        planes view (Cin = in->c, multiple of 16)
 *   use_ring  0: zero padding (TMA out-of-bounds fill over the interior)
 *             1: the ring holds materialised (replicate) padding and is read as data
 *   w_hi/w_lo weights [Cout, KH, KW, Cin] bf16 planes (fsnet_weight_planes)
 *   nprod     3: hi*hi + lo*hi + hi*lo (forward, ~fp32 accuracy)   1: hi*hi only (gradients)
 *   bias      [Cout] fp32 added in the epilogue, or NULL;  relu != 0 applies max(.,0)
 *   out       fp32 view [N, Ho, Wo, Cout]; accumulate != 0 adds to its current content
 *   stats     [2*Cout] fp64, accumulated: per-channel sum and sum of squares of the result
 *             (train-mode BatchNorm statistics, nn.BatchNorm2d in resnet.py / blocks.py), or NULL
 * fsnet_conv_wgrad: acc[Cout, KH, KW, Cin] (fp32, zeroed by the caller) += dy^T * im2col(x); dy is a
 * bf16 plane view [N, Ho, Wo, Cout], x the planes view the forward convolution read.
 * ------------------------------------------------------------------------------------------- */
int fsnet_conv(const fsnet_view* in, int use_ring, const void* w_hi, const void* w_lo, int Cout, int KH, int KW,
               int stride, int pad, int nprod, const float* bias, int relu, const fsnet_view* out, int accumulate,
               double* stats, void* stream);
int fsnet_conv_wgrad(const fsnet_view* x, int use_ring, const fsnet_view* dy, int KH, int KW, int stride, int pad,
                     float* acc, void* stream);

/* ---------------------------------------------------------------------------------------------
 * element-wise / reduction kernels around the convolutions (fsnet_b200/csrc/act_tc.cu)
 *   fsnet_image_to_planes  fp32 NCHW image [N,C,H,W] -> planes with dst->c >= C channels (extra = 0)
 *   fsnet_weight_planes    fp32 [Cout,Cin,KH,KW] -> [Cout_pad,KH,KW,Cin_pad] hi/lo planes and the
 *                          data-gradient operand [Cin_pad,KH,KW (flipped),Cout_pad] (hi), any may be NULL
 *   fsnet_wgrad_to_param   fp32 [Cout_pad,KH,KW,Cin_pad] accumulator -> parameter gradient layout
 *   fsnet_bn_finalize      nn.BatchNorm2d statistics: scale_shift[2C] = (gamma*invstd, beta-mean*gamma*invstd),
 *                          mean_invstd[2C] saved for backward, running stats updated (momentum, unbiased var),
 *                          `stats` re-zeroed; training == 0 uses the running statistics
 *   fsnet_act_planes       dst planes = relu?(raw*scale+shift + residual), optional nearest x2 (`up`=2) into a
 *                          channel slice of a concat buffer; residual: 0 none, 1 planes, 2 raw fp32 with its
 *                          own scale_shift (the down-sample branch, resnet.py:138-145)
 *   fsnet_copy_planes      planes -> channel slice of another planes buffer (skip connection, depth_encoder.py:99-101)
 *   fsnet_maxpool_planes / fsnet_maxpool_bwd   nn.MaxPool2d(3, 2, 1) (resnet.py:122) and its gradient; `argmax`
 *                          [N,Ho,Wo,C] uint8 = window position (0..8) of the first maximum, written by the
 *                          forward (may be NULL there) and read by the backward
 *   fsnet_bn_bwd_reduce / fsnet_bn_bwd_apply   gradient of (ReLU o BatchNorm): sums[2C] = (sum g, sum g*xhat);
 *                          dy (bf16 plane) = gamma*invstd*(g - mean g - xhat*mean g*xhat); mean_invstd == NULL
 *                          means "no BatchNorm" (dy = masked g); a dy view with ring > 0 gets its ring written as zeros
 *                          (zero padding read as data by the folded data gradient); `up`=2 reads the gradient through the adjoint
 *                          of the nearest x2 up-sampling; res_mode 1/2 writes/accumulates the masked gradient
 *                          into another fp32 view (identity residual); dgamma / dbeta (fp32 [c_real], may be NULL) receive
 *                          the BatchNorm parameter gradients; the ReLU mask is `mask` (activation planes)
 *                          or, when mask == NULL and mask_scale_shift != NULL, raw*scale+shift > 0
 *   fsnet_fold_ring        adjoint of replicate padding: adds the ring of a ringed fp32 gradient into its border
 *   fsnet_add_slice        dst (+)= channel slice of src (fp32 views)
 *   fsnet_zero_insert      bf16 plane -> zero-stuffed x2 plane (stride-2 data gradient as a stride-1 convolution)
 * ------------------------------------------------------------------------------------------- */
int fsnet_image_to_planes(const float* img, int C, const fsnet_view* dst, void* stream);
/* same, choosing what the ring (any width) holds: zero_ring = 0 replicate copies of the border, 1 zeros
 * (the 7x7 stem's zero padding, materialised for the folded-tap convolution path) */
int fsnet_image_to_planes_ring(const float* img, int C, const fsnet_view* dst, int zero_ring, void* stream);
int fsnet_weight_planes(const float* w, int Cout, int Cin, int KH, int KW, int Cout_pad, int Cin_pad,
                        void* fwd_hi, void* fwd_lo, void* dgrad_hi, void* stream);
/* one launch for every convolution of a network: `table_device` is a DEVICE array of n_layers descriptors */
typedef struct {
  const float* w;            /* fp32 [cout, cin, kh, kw] parameter */
  void* fwd_hi; void* fwd_lo; void* dgrad_hi;   /* outputs as in fsnet_weight_planes (fwd_lo / dgrad_hi may be NULL) */
  int cout, cin, kh, kw, cout_pad, cin_pad;
} fsnet_weight_desc;
int fsnet_weight_planes_batched(const fsnet_weight_desc* table_device, int n_layers, void* stream);
int fsnet_wgrad_to_param(const float* acc, int Cout, int Cin, int KH, int KW, int Cout_pad, int Cin_pad, float* grad,
                         int accumulate, void* stream);
/* all layers of one backward pass in one launch: accumulator / gradient positions are element offsets from the two
 * base pointers (the executor's pooled accumulator and its flat gradient buffer), so the device table is static */
typedef struct {
  long long acc_off, grad_off;
  int cout, cin, kh, kw, cout_pad, cin_pad;
} fsnet_wgrad_desc;
int fsnet_wgrad_to_param_batched(const fsnet_wgrad_desc* table_device, int n_layers, const float* acc_base, float* grad_base,
                                 void* stream);
int fsnet_bn_finalize(double* stats, double count, const float* gamma, const float* beta, const float* conv_bias,
                      float* running_mean, float* running_var, long long* num_batches, float momentum, float eps,
                      int training, int C, float* scale_shift, float* mean_invstd, void* stream);

/* ---------------------------------------------------------------------------------------------
 * SyncBatchNorm statistics over NVLink / NVSwitch peer memory (fsnet_b200/csrc/peer.cu).  Replaces the all_gather of
 * torch.nn.SyncBatchNorm (the reference converts every BatchNorm, scripts/train.py:100-102) on the data-parallel path.
 *   fsnet_peer          one exchange slot: `bufs` / `flags` are DEVICE arrays [world] of pointers to every rank's data (fp64) and
 *                       flag (u32) buffers, all mapped into this process (symmetric memory); the slot occupies
 *                       [slot_off, slot_off + world*n) doubles and [flag_off, flag_off + world) flags in each of them;
 *                       `seq` is this slot's use counter in LOCAL device memory (zero-initialised, same history on every rank)
 *   fsnet_peer_allreduce_f64   data[0..n) := sum over ranks, bit-identical on every rank (one CTA, one-shot: every rank stores
 *                       into every peer's slot, release/acquire flags at system scope)
 *   fsnet_bn_finalize_sync     the same exchange on stats[2C] fused in front of fsnet_bn_finalize's arithmetic (training mode);
 *                       `count` = elements per channel over ALL ranks
 * Every rank must issue the same sequence of calls (like any collective); the calls are CUDA-graph capturable.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* bufs; const void* flags;
  int rank, world;
  long long slot_off;
  int flag_off;
  void* seq;
} fsnet_peer;
int fsnet_peer_allreduce_f64(double* data, int n, const fsnet_peer* peer, void* stream);
int fsnet_bn_finalize_sync(double* stats, double count, const float* gamma, const float* beta, const float* conv_bias,
                           float* running_mean, float* running_var, long long* num_batches, float momentum, float eps,
                           int C, float* scale_shift, float* mean_invstd, const fsnet_peer* peer, void* stream);
int fsnet_act_planes(const fsnet_view* raw, const float* scale_shift, int res_mode, const fsnet_view* res,
                     const float* res_scale_shift, int relu, int up, const fsnet_view* dst, void* stream);
int fsnet_copy_planes(const fsnet_view* src, const fsnet_view* dst, void* stream);
int fsnet_maxpool_planes(const fsnet_view* src, const fsnet_view* dst, uint8_t* argmax, void* stream);
int fsnet_maxpool_bwd(const fsnet_view* src, const uint8_t* argmax, const fsnet_view* grad_dst, const fsnet_view* grad_src,
                      int accumulate, void* stream);
int fsnet_bn_bwd_reduce(const fsnet_view* g, int up, const fsnet_view* mask, const float* mask_scale_shift, const fsnet_view* raw,
                        const float* mean_invstd, double* sums, void* stream);
int fsnet_bn_bwd_apply(const fsnet_view* g, int up, const fsnet_view* mask, const float* mask_scale_shift, const fsnet_view* raw,
                       const float* mean_invstd, const float* gamma, double* sums, double count,
                       const fsnet_view* dy, int res_mode, const fsnet_view* res, float* dgamma, float* dbeta, int c_real,
                       void* stream);
int fsnet_fold_ring(const fsnet_view* g, void* stream);
int fsnet_add_slice(const fsnet_view* dst, const fsnet_view* src, int accumulate, void* stream);
int fsnet_zero_insert(const fsnet_view* src, const fsnet_view* dst, void* stream);

/* ---------------------------------------------------------------------------------------------
 * fused multi-tensor clip_grad_norm_ + Adam (base_training_hooks.py:46-49: torch.nn.utils.clip_grad_norm_
 * then optimizer.step(); optimizers.py:8: torch.optim.Adam).  SURVEY.md 8(f) N1.
 *   table  device array of fsnet_adam_tensor: parameter, gradient, exp_avg, exp_avg_sq (fp32, n elements) and
 *          chunk_start = index of the tensor's first 4096-element chunk (prefix sum, n_chunks in total)
 *   sumsq  [1] fp64, zeroed by the caller; fsnet_grad_sumsq adds sum(g^2) over all tensors
 *   hyper  [8] fp64 device: lr, beta1, beta2, eps, weight_decay, max_norm (<= 0: no clipping), step, unused;
 *          fsnet_adam_step increments `step`, scales the gradients by min(1, max_norm / (sqrt(sumsq) + 1e-6)) on the
 *          fly and applies torch.optim.Adam's update (bias correction, L2 weight decay, no amsgrad)
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
  long long chunk_start;
} fsnet_adam_tensor;
int fsnet_grad_sumsq(const fsnet_adam_tensor* table, int n_tensors, long long n_chunks, double* sumsq, void* stream);
int fsnet_adam_step(const fsnet_adam_tensor* table, int n_tensors, long long n_chunks, const double* sumsq, double* hyper, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FSNET_B200_H */
