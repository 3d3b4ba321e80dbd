/*
 * cfun_b200 -- C ABI of the B200-native CFUN volumetric hot path (libcfun_b200.so).
 *
 * Conventions (SURVEY.md 8b "C-ABI the replacement must export"):
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - the caller (PyTorch host code) owns every buffer, outputs are pre-allocated;
 *   - activations are fp32, channels-last-3d: memory order N, D, H, W, C  ("NDHWC"); a
 *     torch tensor of logical shape [N,C,D,H,W] with memory_format=torch.channels_last_3d;
 *   - convolution weights keep the checkpoint ABI layout (Cout, Cin, kD, kH, kW) fp32
 *     (reference state_dict, model.py:1329-1339); packed copies are internal;
 *   - `stream` is a cudaStream_t passed as void*; no entry point synchronises it or allocates;
 *   - return 0 on success, negative CFUN_ERR_* otherwise; cfun_last_error() describes it.
 *
 * Each entry point names the reference interface it replaces (file:line in Wuziyi616/CFUN).
 */
#ifndef CFUN_B200_H
#define CFUN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFUN_OK 0
#define CFUN_ERR_INVALID (-1)   /* bad argument / unsupported shape          */
#define CFUN_ERR_CUDA (-2)      /* a CUDA runtime / driver call failed        */
#define CFUN_ERR_WORKSPACE (-3) /* workspace too small                        */
#define CFUN_ERR_NO_DEVICE (-4) /* no sm_100 device                           */

const char* cfun_last_error(void);
int cfun_version(void);
/* number of CUDA kernels this library has launched so far in this process */
unsigned long long cfun_launch_count(void);
/* 1 when the current device is compute capability 10.x (tcgen05 / TMA paths usable). */
int cfun_device_is_sm100(void);
/* Measurement hook (no reference counterpart: the reference times with time.time(), model.py:1354,1560).  After
 * cfun_kernel_timing(1) every tensor-core conv entry point brackets its MAIN kernel (not the operand packs) with CUDA
 * events on the caller's stream; cfun_last_kernel_ms waits for the most recent one and returns its duration.  bench.py
 * uses it for the live roofline of the dominant kernels.  cfun_kernel_timing(0) switches it off (the default). */
int cfun_kernel_timing(int on);
int cfun_last_kernel_ms(float* ms);

/* ------------------------------------------------------------------------------------------
 * conv3d  -- replaces every nn.Conv3d on the path: backbone.py:14-55,124 ; model.py:131-134,
 * 713-717,758-760 ; mask_branch.py:23-89 (forward) and their autograd backward (model.py:1640).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int N, Cin, Din, Hin, Win;
  int Cout, Dout, Hout, Wout;
  int kD, kH, kW;
  int sD, sH, sW;
  int pD, pH, pW;
} cfun_conv3d_desc;

#define CFUN_CONV_ALGO_AUTO 0
#define CFUN_CONV_ALGO_SIMT 1 /* fp32 CUDA-core implicit GEMM (any shape)                       */
#define CFUN_CONV_ALGO_TC 2   /* tcgen05 implicit GEMM, split-bf16 x3 operands, fp32 TMEM accum */
#define CFUN_CONV_ALGO_TC1 3  /* tcgen05, single bf16 pass ("fast mode", NOT parity grade)      */

#define CFUN_PASS_FWD 0
#define CFUN_PASS_BWD_DATA 1
#define CFUN_PASS_BWD_WEIGHT 2

/* epilogue flags for cfun_conv3d_fwd */
#define CFUN_EPI_BIAS 1
#define CFUN_EPI_RELU 2

size_t cfun_conv3d_workspace_size(const cfun_conv3d_desc* d, int pass, int algo);
/* which algorithm AUTO resolves to for (desc, pass): one of CFUN_CONV_ALGO_{SIMT,TC} */
int cfun_conv3d_pick_algo(const cfun_conv3d_desc* d, int pass);
/* 1 when `algo` can execute (desc, pass) on the current device */
int cfun_conv3d_supported(const cfun_conv3d_desc* d, int pass, int algo);
int cfun_conv3d_fwd(const cfun_conv3d_desc* d, const float* x, const float* w, const float* bias, float* y,
                    int epi_flags, int algo, void* ws, size_t ws_bytes, void* stream);
int cfun_conv3d_bwd_data(const cfun_conv3d_desc* d, const float* dy, const float* w, float* dx, int algo, void* ws,
                         size_t ws_bytes, void* stream);
/* dw (Cout,Cin,kD,kH,kW) is overwritten; dbias may be NULL */
int cfun_conv3d_bwd_weight(const cfun_conv3d_desc* d, const float* x, const float* dy, float* dw, float* dbias,
                           int algo, void* ws, size_t ws_bytes, void* stream);

/* Fused backward for the 3x3x3 / stride-1 convs whose three passes run on the halo-family tcgen05 kernels (the U-Net
 * mask branch mask_branch.py:23-89, FPN smoothing model.py:133-134, RPN.conv_shared model.py:713): the split-bf16 operand
 * pack of X written by the forward is kept for the weight gradient, and dY is packed once for both gradients.
 *   cfun_conv3d_pack_bytes            bytes of the X pack the caller must own (0 = shape not eligible, use the calls above)
 *   cfun_conv3d_fwd_keep_pack         = cfun_conv3d_fwd, leaving the pack in xpack (128-byte aligned)
 *   cfun_conv3d_bwd_fused             = cfun_conv3d_bwd_data (dx may be NULL) + cfun_conv3d_bwd_weight (dw / dbias may be NULL)
 * Forward workspace: cfun_conv3d_workspace_size(d, CFUN_PASS_FWD, CFUN_CONV_ALGO_AUTO). */
size_t cfun_conv3d_pack_bytes(const cfun_conv3d_desc* d);
size_t cfun_conv3d_bwd_fused_workspace_size(const cfun_conv3d_desc* d);
int cfun_conv3d_fwd_keep_pack(const cfun_conv3d_desc* d, const float* x, const float* w, const float* bias, float* y,
                              int epi_flags, void* xpack, size_t xpack_bytes, void* ws, size_t ws_bytes, void* stream);
/* cfun_conv3d_fwd_keep_pack whose epilogue also accumulates the InstanceNorm statistics of y (mask_branch.py:18: every
 * U-Net conv is followed by InstanceNorm3d): stat_acc [N][Cout][2] doubles = per-(sample, channel) sum and sum of squares,
 * zeroed by the call; cfun_instnorm_finalize turns them into mean / rstd.  Saves the separate read of y that
 * cfun_instnorm_stats makes. */
int cfun_conv3d_fwd_stats(const cfun_conv3d_desc* d, const float* x, const float* w, const float* bias, float* y,
                          int epi_flags, void* xpack, size_t xpack_bytes, double* stat_acc, void* ws, size_t ws_bytes,
                          void* stream);
int cfun_conv3d_bwd_fused(const cfun_conv3d_desc* d, const void* xpack, size_t xpack_bytes, const float* dy, const float* w,
                          float* dx, float* dw, float* dbias, void* ws, size_t ws_bytes, void* stream);
/* cfun_conv3d_fwd_keep_pack on leaky_relu(x * scale[n][c], slope) (scale NULL = 1) -- the LeakyReLU / Dropout3d that precede a
 * conv (mask_branch.py:127-131) applied on the way into its operand pack; the activated tensor is never written.  The
 * backward (cfun_conv3d_bwd_fused on the kept pack) yields the gradient w.r.t. the activated input; cfun_affine_act_bwd(x,
 * scale, 0, ...) turns it into the gradient of x. */
int cfun_conv3d_preact_supported(const cfun_conv3d_desc* d);
int cfun_conv3d_fwd_keep_pack_preact(const cfun_conv3d_desc* d, const float* x, const float* scale, float slope, const float* w,
                                     const float* bias, float* y, int epi_flags, void* xpack, size_t xpack_bytes, void* ws,
                                     size_t ws_bytes, void* stream);
/* The decoder's InstanceNorm -> LeakyReLU -> Upsample(x2, nearest) -> conv (mask_branch.py:91-103) without the upsampled
 * tensor: cfun_instnorm_up2_pack applies the norm coefficients a, b [N][C] and the activation to the low-resolution x and
 * writes the 2x2x2-replicated result straight into the conv's operand pack (hi, lo: [G][N*(2D+2P)][2H][2W][8], G =
 * align16(C)/8, halves of cfun_conv3d_pack_bytes); cfun_conv3d_fwd_stats_packed runs the conv on that pack. */
int cfun_instnorm_up2_pack(const float* x, const float* a, const float* b, int N, int D, int H, int W, int C, float slope,
                           void* hi, void* lo, int G, int P, void* stream);
int cfun_conv3d_fwd_stats_packed(const cfun_conv3d_desc* d, void* xpack, size_t xpack_bytes, const float* w, float* y,
                                 double* stat_acc, void* ws, size_t ws_bytes, void* stream);
/* cfun_conv3d_fwd_stats on the channel concatenation [a (C1) | b (C2)] (the U-Net decoder's torch.cat((up, skip), 1) -> conv,
 * mask_branch.py:185-205): the operand pack is built from the two tensors, the concatenated tensor never exists.
 * stat_acc may be NULL. */
int cfun_conv3d_cat_supported(const cfun_conv3d_desc* d, int C1, int C2);
int cfun_conv3d_fwd_stats_cat(const cfun_conv3d_desc* d, const float* a, int C1, const float* b, int C2, const float* w, float* y,
                              void* xpack, size_t xpack_bytes, double* stat_acc, void* ws, size_t ws_bytes, void* stream);
/* The same backward with the dY pack made by the caller (cfun_instnorm_bwd_apply_pack writes the InstanceNorm backward's
 * result straight into it, so the conv's output gradient never exists in fp32):
 *   cfun_conv3d_dy_pack_geometry  -> bytes of hi + lo (0 = not eligible), channel groups, zero planes per side */
size_t cfun_conv3d_dy_pack_geometry(const cfun_conv3d_desc* d, int* groups, int* pad_planes);
int cfun_conv3d_bwd_fused_packed(const cfun_conv3d_desc* d, const void* xpack, size_t xpack_bytes, void* ypack,
                                 size_t ypack_bytes, const float* w, float* dx, float* dw, void* ws, size_t ws_bytes,
                                 void* stream);

/* Classifier.conv1 (model.py:758): a kernel-size == input-size conv, i.e. a [M,K]x[Nout,K]^T product with
 * K = Cin*kD*kH*kW (221184) and M = #RoIs (12).  x is NCDHW-contiguous [M,K]; w is [Nout,K]. */
int cfun_fc_fwd(int M, int Nout, long long K, const float* x, const float* w, const float* bias, float* y, void* stream);
int cfun_fc_bwd_data(int M, int Nout, long long K, const float* dy, const float* w, float* dx, void* stream);
int cfun_fc_bwd_weight(int M, int Nout, long long K, const float* dy, const float* x, float* dw, float* dbias,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * normalisation / activation / resampling passes around the convs
 *   InstanceNorm3d(affine=False, eps=1e-5) + LeakyReLU(0.01) + Dropout3d + nearest x2 + cat:
 *   mask_branch.py:18-20,91-122,124-220 ; frozen BatchNorm3d + ReLU + residual: backbone.py:26-114 ;
 *   MaxPool3d(2,2): backbone.py:127 ; F.upsample x2: model.py:144.
 * ------------------------------------------------------------------------------------------ */
/* per-(n,c) mean and 1/sqrt(var+eps) over S = D*H*W voxels of an NDHWC tensor (biased variance).
 * acc is a zeroed scratch of 2*N*C doubles. */
int cfun_instnorm_stats(const float* x, int N, long long S, int C, float eps, double* acc, float* mean, float* rstd,
                        void* stream);
/* y[n, up(s), c_off + c] = act( x[n,s,c] * a[n*a_nstride + c] + b[n*a_nstride + c] (+ r[n,s,c]) )
 *   act = leaky relu with `slope` (0 -> ReLU, 1 -> identity); a/b may be NULL (1/0);
 *   a_nstride = C for per-(n,c) coefficients, 0 for per-channel; up in {1,2}: nearest x`up` upsampling of the
 *   output; the destination has Ctot channels (concat fusion). */
int cfun_affine_act_fwd(const float* x, const float* a, const float* b, int a_nstride, const float* r, float* y, int N,
                        int D, int H, int W, int C, int Ctot, int c_off, int up, float slope, void* stream);
/* backward of the above w.r.t. x (and r): g[n,s,c] = sum_{children} dy[n, child(s), c_off+c] * act'(pre);
 *   pre is recomputed from x (and r).  dx = g * a ; dr = g (if dr != NULL).
 *   When stat_acc != NULL (InstanceNorm backward) it also accumulates, per (n,c), sum(g*a') and sum(g*a'*xhat) as
 *   doubles into stat_acc[2*N*C] where xhat = x*a+b (valid when a,b are the instance-norm coefficients and r==NULL),
 *   and writes g into dx un-scaled; cfun_instnorm_bwd_apply then finishes dx. */
int cfun_affine_act_bwd(const float* x, const float* a, const float* b, int a_nstride, const float* r, const float* dy,
                        float* dx, float* dr, double* stat_acc, int N, int D, int H, int W, int C, int Ctot, int c_off,
                        int up, float slope, void* stream);
/* dx = a * (g - mean_s(g) - xhat * mean_s(g*xhat)) in place on dx (which holds g), xhat = x*a + b with the same
 * per-(n,c) coefficients a[N*C], b[N*C] the forward used (a = rstd, b = -mean*rstd; a dropout channel scale m folds in
 * as a = m*rstd', b = -m*mean*rstd'). */
int cfun_instnorm_finalize(const double* acc, int N, long long S, int C, float eps, float* mean, float* rstd, void* stream);
int cfun_instnorm_bwd_apply(const float* x, const float* a, const float* b, const double* stat_acc, float* dx, int N,
                            long long S, int C, void* stream);
/* The U-Net's level-1 residual sum feeds both the skip connection and the norm (mask_branch.py:132-136:
 * out += residual_1; context_1 = lrelu(out); out = inorm3d_c1(out)):
 *   cfun_add_act_stats             s = a + b, ctx = leaky_relu(s), statistics of s (as cfun_instnorm_stats) in one pass
 *   cfun_instnorm_bwd_apply_extra  cfun_instnorm_bwd_apply + leaky_relu'(x) * dextra: both gradients of s in one pass */
int cfun_add_act_stats(const float* a, const float* b, float* s, float* ctx, int N, long long S, int C, float slope, float eps,
                       double* acc, float* mean, float* rstd, void* stream);
int cfun_instnorm_bwd_apply_extra(const float* x, const float* a, const float* b, const double* stat_acc, float* dx, int N,
                                  long long S, int C, const float* dextra, float slope, void* stream);
/* cfun_instnorm_bwd_apply writing split-bf16 group-planar pack rows (hi, lo: [G][N*(D+2P)][H][W][8], zero planes
 * included) instead of fp32; g = the buffer cfun_affine_act_bwd left the un-normalised gradient in */
int cfun_instnorm_bwd_apply_pack(const float* x, const float* a, const float* b, const double* stat_acc, const float* g,
                                 int N, int D, int H, int W, int C, void* hi, void* lo, int G, int P, void* stream);
/* torch.cat([a, b], dim=1) of two NDHWC tensors with M = N*D*H*W rows (mask_branch.py:189,197,204,211) and its backward */
int cfun_cat2_channels(const float* a, int C1, const float* b, int C2, float* out, long long M, void* stream);
int cfun_split2_channels(const float* cat, int C1, int C2, float* a, float* b, long long M, void* stream);
int cfun_maxpool2_fwd(const float* x, float* y, int N, int D, int H, int W, int C, void* stream);
int cfun_maxpool2_bwd(const float* x, const float* y, const float* dy, float* dx, int N, int D, int H, int W, int C,
                      void* stream);
/* debugging aid for the tcgen05 pipelines: out[0] != 0 means an mbarrier wait timed out (out[0]-1 = wait site,
 * out[1..5] = blockIdx.x, blockIdx.y, threadIdx.x, parity, spins).  Synchronises the device; reading resets the record. */
int cfun_tc_debug_status(int* out8_host);
/* bring-up aid: 0 = stream not capturing, 1 = capturing into a CUDA graph, 2 = its capture has been invalidated */
int cfun_stream_capture_status(void* stream);
/* split fp32 -> (hi, lo) bf16 pairs, channel-padded NDHWC, the operand format of the tcgen05 convs */
int cfun_pack_split_bf16(const float* x, void* hi, void* lo, long long rows, int C, int Cpad, void* stream);
/* group-planar split pack of an NDHWC activation, the operand format of the halo / hx / weight-gradient kernels:
 * hi, lo: [G][N*(D+2P)][H][W][8] bf16, P zero planes before and after every sample, G >= ceil(C/8), lo may be NULL.
 * Exposed for layout tests and bandwidth measurements; the convolutions pack internally. */
int cfun_pack_act_gp(const float* x, void* hi, void* lo, int N, int D, int H, int W, int C, int G, int P, void* stream);

/* ------------------------------------------------------------------------------------------
 * RoI crop + trilinear(align_corners=True) resize -- replaces model.RoI_Align (model.py:265-289) and the
 * gather / scatter half of model.pyramid_roi_align (model.py:292-370).
 *   boxes: [n,6] normalised (z1,y1,x1,z2,y2,x2); level[n] selects fmap 0/1 (NULL -> all 0);
 *   fmapK: NDHWC [1,Dk,Hk,Wk,C];  out: NCDHW-contiguous [n,C,pd,ph,pw] when out_ncdhw else NDHWC.
 *   Empty / invalid crops produce zero rows (the reference's bare except, model.py:281-287).
 * ------------------------------------------------------------------------------------------ */
int cfun_roi_crop_resize_fwd(const float* fmap0, int D0, int H0, int W0, const float* fmap1, int D1, int H1, int W1,
                             int C, const float* boxes, const int* level, int n, int pd, int ph, int pw, float* out,
                             int out_ncdhw, void* stream);
int cfun_roi_crop_resize_bwd(float* dfmap0, int D0, int H0, int W0, float* dfmap1, int D1, int H1, int W1, int C,
                             const float* boxes, const int* level, int n, int pd, int ph, int pw, const float* dout,
                             int out_ncdhw, void* stream);
/* model.py:322-332: level = clamp(round(4 + log2(h*w*d)/3), 2, 3) - 2 on normalised boxes */
int cfun_roi_level(const float* boxes, int n, int* level, void* stream);

/* ------------------------------------------------------------------------------------------
 * boxes: sort / decode / clip / NMS / overlaps / targets -- replaces model.proposal_layer (model.py:199-258),
 * model.apply_box_deltas (:155), model.clip_boxes (:185), utils.non_max_suppression + utils.compute_iou
 * (utils.py:122-157, 50-70), model.bbox_overlaps (:377-411), utils.box_refinement (utils.py:92-119) and the
 * GT-mask crop + nearest resize of model.detection_target_layer (model.py:481-493).
 * ------------------------------------------------------------------------------------------ */
/* order[0..n) = indices sorted by (score descending, index ascending).  ws: cfun_sort_workspace_size(n) bytes */
size_t cfun_sort_workspace_size(int n);
int cfun_sort_desc(const float* scores, int n, int* order, void* ws, size_t ws_bytes, void* stream);
/* boxes_out[i] = clip(decode(anchors[order[i]], deltas[order[i]] * std), window), scores_out[i] = scores[order[i]*sstride+soff] */
int cfun_decode_clip(const float* anchors, const float* deltas, const float* scores, int sstride, int soff,
                     const int* order, int k, const float* std6_host, const float* window6_host, float* boxes_out,
                     float* scores_out, void* stream);
/* Greedy 3-D NMS, bit-exact with the reference's unfused fp32 arithmetic.  boxes must already be in descending
 * score order (as in proposal_layer).  keep[max_num] receives indices, *count the number kept.
 * ws: cfun_nms_workspace_size(n). */
size_t cfun_nms_workspace_size(int n);
int cfun_nms3d(const float* boxes, int n, float threshold, int max_num, int* keep, int* count, void* ws,
               size_t ws_bytes, void* stream);
/* rows_out[i] = rows[idx[i]] / div6 (i < *count, rows >= *count zeroed); proposal normalisation model.py:247-253 */
int cfun_gather_boxes(const float* rows, const int* idx, const int* count, int max_rows, const float* div6_host,
                      float* rows_out, void* stream);
/* utils.compute_iou (utils.py:50-70): IoU (with the +1e-6) of one box against n boxes */
int cfun_iou3d_eps(const float* box, const float* boxes, int n, float* out, void* stream);
int cfun_bbox_overlaps3d(const float* boxes1, int n1, const float* boxes2, int n2, float* iou, void* stream);
int cfun_box_refinement(const float* box, const float* gt_box, int n, const float* std6_host, float* deltas, void* stream);
/* model.detection_target_layer (model.py:414-563) without host round trips, split at its one data-dependent point (the
 * two torch.randperm draws on the host generator need the candidate counts):
 *   cfun_roi_candidates  rois[max_rows,6] = boxes_sorted[keep[i]] / div6 (rows >= *count zero) = the proposal normalisation;
 *                        iou_max / assign = max and argmax IoU over the n_gt ground-truth boxes (bbox_overlaps + max);
 *                        pos_list / neg_list = ascending indices with IoU >= / < threshold (the two torch.nonzero calls);
 *                        counts3 = {*count, positives, negatives}.  One launch, one block.
 *   cfun_roi_targets     rows 0..P-1 = rois[pos_list[perm[i]]], rows P..R-1 = rois[neg_list[perm[i]]]; class id and
 *                        box_refinement deltas of the matched ground-truth box for the positives, zeros for the negatives. */
int cfun_roi_candidates(const float* boxes_sorted, const int* keep, const int* count, int max_rows, const float* div6_host,
                        const float* gt_boxes, int n_gt, float iou_threshold, float* rois, float* iou_max, int* assign,
                        int* pos_list, int* neg_list, int* counts3, void* stream);
int cfun_roi_targets(const float* rois, const int* assign, const int* pos_list, const int* neg_list, const long long* perm,
                     int P, int R, const float* gt_boxes, const int* gt_class_ids, const float* std6_host, float* out_rois,
                     long long* class_ids, float* deltas, void* stream);
/* label: int32 [D,H,W] class-id volume; rois [P,6] normalised; out: one-hot float64 [P,ncls,md,mh,mw]
 * (the reference's target layout, model.py:481-493) and/or class index int64 [P,md,mh,mw] (either may be NULL). */
int cfun_mask_target_crop(const int* label, int D, int H, int W, const float* rois, int P, int ncls, int md, int mh,
                          int mw, double* onehot, long long* cls_index, void* stream);

/* ------------------------------------------------------------------------------------------
 * 3-D Sobel edge loss -- replaces model.compute_mrcnn_mask_edge_loss (model.py:938-981) forward + backward.
 *   pred: probabilities NDHWC [P, M,M,M, ncls]; tgt_index int64 [P,M,M,M] (class id per voxel; plane j is
 *   (tgt==j+1)); classes 1..7 (the reference's literal range(7)); loss = sum_ij mse(mag_p, mag_t) / P.
 *   magnitude = sqrt(g0^2 + g1^2 + g0^2) -- response 0 twice, response 2 unused (model.py:969-972).
 * ------------------------------------------------------------------------------------------ */
/* The crop is Md x Mh x Mw (cubic for the heart configs, (32,80,80) / (64,160,160) for LiTS); classes 1..ncls-1 contribute.
 * mode 0: heart, MSE between the magnitudes sqrt(g0^2+g1^2+g0^2) (model.py:969-975); mode 1: LiTS, MSE between the raw three
 * Sobel responses (LiTS_2017/model.py:967-975). */
size_t cfun_sobel_edge_workspace_size(int P, int Md, int Mh, int Mw, int ncls, int mode);
int cfun_sobel_edge_loss_fwd(const float* pred, const long long* tgt_index, int P, int Md, int Mh, int Mw, int ncls, int mode,
                             float* loss, void* ws, size_t ws_bytes, void* stream);
/* dpred = grad_scale[0] * dloss/dpred (every element written) */
int cfun_sobel_edge_loss_bwd(const float* pred, const long long* tgt_index, int P, int Md, int Mh, int Mw, int ncls, int mode,
                             const float* grad_scale, float* dpred, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Mask cross-entropy -- replaces nn.CrossEntropyLoss over the mask logits (model.py:909-935; with class weights
 * LiTS_2017/model.py:926).  logits channels-last [V][C] fp32 (V = P*d*h*w), target int64 [V] class ids, weight NULL or [C].
 * loss = sum_v w[y_v] * (logsumexp(x_v) - x_v[y_v]) / sum_v w[y_v]   (torch's 'mean' reduction).
 * acc: 2 doubles owned by the caller (zeroed by the forward, read by the backward).
 * backward: dlogits[v][c] = grad_scale[0] * w[y_v] * (softmax(x_v)[c] - [c == y_v]) / sum_v w[y_v].
 * ------------------------------------------------------------------------------------------ */
int cfun_mask_ce_fwd(const float* logits, const long long* target, long long V, int C, const float* weight, double* acc,
                     float* loss, void* stream);
int cfun_mask_ce_bwd(const float* logits, const long long* target, long long V, int C, const float* weight, const double* acc,
                     const float* grad_scale, float* dlogits, void* stream);

/* ------------------------------------------------------------------------------------------
 * optimizer tail -- replaces clip_grad_norm_(5.0) + SGD(momentum, weight decay) (model.py:1538-1545,1641-1644)
 * over one flat fp32 parameter / gradient / momentum buffer.  wd_mask[i] in {0,1} marks decayed elements.
 * ------------------------------------------------------------------------------------------ */
int cfun_sumsq(const float* g, long long n, double* out_acc, void* stream); /* *out_acc += sum g^2 */
int cfun_sgd_clip_step(float* p, const float* g, float* mom, const unsigned char* wd_mask, long long n,
                       const double* sumsq, float max_norm, float lr, float momentum, float weight_decay,
                       int first_step, void* stream);
/* int16 CT volume -> fp32, (x-mean)/std with population std (model.mold_image, model.py:1902-1904);
 * vol is [H,W,D] int16 as stored; out is [1,1,D,H,W] (== NDHWC for C=1).  acc: 2 zeroed doubles. */
int cfun_mold_volume_i16(const short* vol_hwd, int H, int W, int D, double* acc, float* out_dhw, void* stream);

/* ------------------------------------------------------------------------------------------
 * Inference pre / post-processing on the device.
 *
 * cfun_resize_linear3d: utils.resize_image(mode='self') (reference utils.py:389-393, called from MaskRCNN.mold_inputs,
 * model.py:1774-1810): order-1 resize of a raw scan src[H][W][D] to dst[H2][W2][D2], skimage >= 0.19 / scipy.ndimage.zoom
 * (order 1, 'grid-constant', grid_mode=True) semantics in float64, cast back to the scan dtype by C truncation.
 * dtype 0 = int16, 1 = float32 (same type in and out).
 *
 * cfun_unmold_mask_argmax: utils.unmold_mask + np.argmax of MaskRCNN.unmold_detections (reference utils.py:443-460,
 * model.py:1851-1853) fused: trilinear (align_corners=False) resize of one detection's class-probability crop
 * (ncls x md x mh x mw, element (c,z,y,x) at mask[c*class_stride + ((z*mh+y)*mw+x)*voxel_stride]) to the box
 * box6_host = (z1,y1,x1,z2,y2,x2) (host ints, pixels of the original scan), zeros elsewhere, argmax over classes;
 * out[H][W][D] uint8 class ids (the [H,W,D] order unmold_detections returns).
 * ------------------------------------------------------------------------------------------ */
int cfun_resize_linear3d(const void* src, int H, int W, int D, void* dst, int H2, int W2, int D2, int dtype, void* stream);
int cfun_unmold_mask_argmax(const float* mask, int ncls, int md, int mh, int mw, long long class_stride, long long voxel_stride,
                            const int* box6_host, int D, int H, int W, unsigned char* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CFUN_B200_H */
