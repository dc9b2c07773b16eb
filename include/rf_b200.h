/* rf_b200.h - C ABI of the B200-native RetrievalFuse hot path.
 *
 * The reference (nihalsid/retrieval-fuse) has no FFI on this path: it is Python
 * nn.Modules + pyflann.  This header is the drop-in boundary a maintainer binds
 * with ctypes; each entry point names the reference
 * function (file:line under /root/reference) whose arithmetic it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - tensors are fp32, contiguous, NCDHW, exactly as the reference's torch
 *     tensors; indices are int32; kNN distances travel as fp64 until the last
 *     step so that shard merges stay bit-exact;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - return value 0 = ok, non-zero = error, message via rf_last_error()
 *     (thread-local); kernels are asynchronous on `stream`;
 *   - no hidden global state, no allocation: callers pass workspaces.
 *   - there is NO CPU fallback: every function launches sm_100a kernels.
 */
#ifndef RF_B200_H
#define RF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RF_ACT_NONE 0
#define RF_ACT_RELU 1
#define RF_ACT_LEAKY 2 /* slope argument */
#define RF_ACT_TANH 3

const char* rf_last_error(void);
int rf_version(void);
/* sm count / compute capability of `device`; fails if it is not sm_100. */
int rf_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);

/* Handle (SURVEY 8b "no hidden global state besides a handle from rf_create(device) / rf_destroy").  The kernels keep
 * no mutable global state - every entry point takes its buffers, workspaces and stream as arguments - so the handle
 * only records its device and performs the per-device one-time setup (opt-in to > 48 KB dynamic shared memory for the
 * tcgen05 kernels) eagerly; entry points called without a handle do that setup lazily on first use. */
typedef struct rf_handle_s rf_handle;
int rf_create(int device, rf_handle** out);
int rf_destroy(rf_handle* handle);
int rf_handle_device(const rf_handle* handle, int* device, int* sm_count);

/* ---- a2-a4  patch fold / unfold (bit-exact re-indexing) ------------------- */

/* model/attention.py:186-188 Unfold3D.forward: [B,C,S,S,S] -> [B*(S/E)^3,C,E,E,E] */
int rf_unfold3d(const float* x, float* out, int B, int C, int S, int E, void* stream);
/* model/attention.py:170-176 Fold3D.forward: [B*R^3,C,E,E,E] -> [B,C,R*E,R*E,R*E] (contiguous) */
int rf_fold3d(const float* x, float* out, int B, int C, int R, int E, void* stream);
/* model/attention.py:200-203 Unfold3DPadStride.forward and util/patcher.py:14-19
 * Patcher.__call__ (per-axis kernel/pad/stride).  x [B,C,size] -> out
 * [B,n0,n1,n2,C,k0,k1,k2] (the reference then views it as (-1,1,k,k,k) or
 * (-1,C,k,k,k)).  When norm_div != 0 every value (padding included) is mapped
 * v -> (v - norm_sub) / norm_div in fp32, which fuses
 * dataset/patched_scene_dataset.py:127-128 after dataset/scene.py:61 padding. */
int rf_unfold3d_pad_stride(const float* x, float* out, int B, int C, const int size[3], const int kernel[3],
                           const int pad[3], const int stride[3], float pad_val, float norm_sub, float norm_div,
                           void* stream);
/* util/patcher.py:21-30 Patcher.recompose_patches: patches [B,n0*n1*n2,k0,k1,k2]
 * -> out [B,C,size] (padding cropped); later patches overwrite earlier ones. */
int rf_recompose_patches(const float* patches, float* out, int B, int C, const int size[3], const int kernel[3],
                         const int pad[3], const int stride[3], const int count[3], float pad_val, void* stream);

/* ---- building blocks of a5-a8, a13-a16 ----------------------------------- */

/* Implicit-GEMM 3D convolution / linear layer, fp32 accumulate (torch.nn.Conv3d
 * as used by model/retrieval.py:4-388 and model/unet.py:15-16, nn.Linear as in
 * model/retrieval.py:64-84, model/attention.py:29-46).
 *   x  [N,Cin,Di,Hi,Wi]; wt [Cin*KS^3, Cout] = weight.reshape(Cout,-1).T
 *   y  [N,Cout,Do,Ho,Wo], Do = (Di + 2*pad - KS)/stride + 1
 *   optional fused GroupNorm on the INPUT ('g' before 'c', model/unet.py:54-66):
 *     v = (x - gn_mu[n,ci]) * gn_a[n,ci] + gn_beta[ci], zero padding applied
 *     after it; pass gn_mu = NULL to disable;
 *   optional virtual input = concat(x [N,Cin-C2,...], nearest-2x-upsample(x2
 *     [N,C2,Di/2,Hi/2,Wi/2])) along channels (model/unet.py:297-306 Decoder
 *     joining); pass x2 = NULL / C2 = 0 to disable;
 *   epilogue: y = act((acc + bias[co]) * oscale[co] + oshift[co]); any of
 *     bias / oscale / oshift may be NULL (eval BatchNorm3d folds into them). */
int rf_conv3d_fwd(const float* x, const float* x2, int C2, const float* wt, const float* bias, const float* oscale,
                  const float* oshift, const float* gn_mu, const float* gn_a, const float* gn_beta, float* y, int N,
                  int Cin, int Di, int Hi, int Wi, int Cout, int KS, int stride, int pad, int act, float slope,
                  void* stream);
/* y[M,N] = act(x[M,K] @ wt[K,N] + bias) - the same kernel with KS = 1. */
int rf_linear_fwd(const float* x, const float* wt, const float* bias, float* y, int M, int K, int N, int act,
                  float slope, void* stream);
/* torch.nn.GroupNorm statistics (model/unet.py:66): per (n, group) mean and
 * 1/sqrt(var+eps) over cpg*spatial elements of the virtual input (same x / x2
 * concat rule as rf_conv3d_fwd), expanded to gn_mu[n*C+c] = mean and
 * gn_a[n*C+c] = rstd * gamma[c]. */
int rf_groupnorm_stats(const float* x, const float* x2, int C2, const float* gamma, float* gn_mu, float* gn_a, int N,
                       int C, int D, int H, int W, int groups, float eps, void* stream);
/* nn.MaxPool3d(2) (model/unet.py:238): [N,C,D,H,W] -> [N,C,D/2,H/2,W/2] */
int rf_maxpool3d_2(const float* x, float* y, int N, int C, int D, int H, int W, void* stream);
/* F.interpolate(mode='nearest') by exactly 2x (model/unet.py:352-358): -> [N,C,2D,2H,2W] */
int rf_upsample_nearest_2(const float* x, float* y, int N, int C, int D, int H, int W, void* stream);
/* F.normalize(x, dim=1) on rows (util/retrieval.py:38,66): y = x / max(||x||_2, eps) */
int rf_l2_normalize_rows(const float* x, float* y, long M, int D, float eps, void* stream);

/* Tensor-core linear layer (tcgen05, fp16 hi/lo split: Xh.Wh + Xh.Wl + Xl.Wh,
 * fp32 accumulators in TMEM; ~2e-7 relative accuracy).  The weight
 * [N, K] (nn.Linear.weight, row-major) is staged once as a pre-split,
 * pre-swizzled operand image (rf_tc_weight_image, 1024-byte aligned buffer of
 * rf_tc_weight_image_bytes); then y[M,N] = act(x[M,K] (row stride ldx) @ W^T + bias).
 * N in {32, 64, 96, 128, 256, 384, 512}; x, y 16-byte aligned, ldx % 4 == 0. */
size_t rf_tc_weight_image_bytes(int N, int K);
int rf_tc_weight_image(const float* w, int N, int K, void* image, void* stream);
int rf_tc_linear_fwd(const float* x, int ldx, const void* weight_image, const float* bias, float* y, long M, int K, int N,
                     int act, float slope, void* stream);

/* Fused MLP chain on tcgen05 (model/retrieval.py:64-133 Patch04 / Patch05 / Patch04V2 forward +
 * util/retrieval.py:66 normalisation; model/attention.py:29-46 AttentionFeatureEncoder): all layers of
 * y = L_n(act(... act(L_1 x))) in ONE launch, hidden activations never leave the SM (fp16 hi/lo operand
 * planes in shared memory, fp32 accumulators in TMEM).  widths_host[0..n_layers] are the layer widths
 * (each <= 512, widths[0] <= 256, no two consecutive hidden widths above 256); images_host[l] is the operand
 * image of layer l's nn.Linear weight [widths[l+1], widths[l]] (rf_tc_mlp_weight_image); `act`/`slope` apply
 * between layers (not after the last); l2_normalize divides output rows by max(|row|, eps). */
int rf_tc_mlp_supported(const int* widths_host, int n_layers);
size_t rf_tc_mlp_weight_image_bytes(int N, int K);
int rf_tc_mlp_weight_image(const float* w, int N, int K, void* image, void* stream);
int rf_tc_mlp_debug_read(long long* out64); /* tuning aid: phase timestamps of CTA 0's second tile */
int rf_tc_mlp_fwd(const float* x, int ldx, const void* const* images_host, const float* const* bias_host,
                  const int* widths_host, int n_layers, int act, float slope, int l2_normalize, float eps, float* y, int ldy,
                  long M, void* stream);

/* Tensor-core 3D convolution on channels-last activations (U-Net 'gcr' blocks,
 * conv patch encoders).  One SingleConv = rf_cl_gn_stats -> rf_cl_norm_split
 * (normalise once, split to fp16 hi/lo, pad channels to 8) -> rf_tc_conv3d_fwd
 * (implicit GEMM on tcgen05, three fp16 products per K step, fp32 accumulate).
 *   fp32 channels-last tensors are [N,D,H,W,C]; split tensors are two fp16
 *   arrays [N,D,H,W,Cp], Cp = C rounded up to 8.  The optional second source
 *   (x2, half resolution) is virtually nearest-upsampled and concatenated after
 *   x's channels (model/unet.py:297-306).  Cout <= 128.
 *   `scale` arguments are powers of two that move activations / weights into
 *   fp16's comfortable range before the hi/lo split (so that the lo parts are
 *   normal numbers); the conv multiplies its accumulator by out_scale =
 *   1 / (activation scale * weight scale) before bias and activation. */
size_t rf_cl_gn_stats_workspace_bytes(int N, int C);
int rf_cl_gn_stats(const float* x, const float* x2, int C2, const float* gamma, float* gn_mu, float* gn_a, int N, int C,
                   int D, int H, int W, int groups, float eps, void* workspace, void* stream);
int rf_cl_norm_split(const float* x, const float* gn_mu, const float* gn_a, const float* gn_beta, int c_off, int c_tot,
                     void* hi, void* lo, long N, long S, int C, int Cp, float scale, void* stream);
int rf_cl_maxpool3d_2(const float* x, float* y, int N, int D, int H, int W, int C, void* stream);
int rf_cl_transpose(const float* in, float* out, long N, long S, int C, int to_channels_last, void* stream);
size_t rf_tc_conv_weight_image_bytes(int Cout, int C1, int C2, int KS);
int rf_tc_conv_weight_image(const float* w, int Cout, int C1, int C2, int KS, float scale, void* image, void* stream);
int rf_tc_conv3d_fwd(const void* x_hi, const void* x_lo, int C1, const void* x2_hi, const void* x2_lo, int C2,
                     const void* weight_image, const float* bias, float* y, int N, int Di, int Hi, int Wi, int Cout, int KS,
                     int stride, int pad, int act, float slope, float out_scale, int out_ncdhw, void* stream);

/* "Shifted window" variant of the tensor-core convolution for the 3x3x3 / stride 1 /
 * pad 1 layers of the U-Nets (model/unet.py:19-100 create_conv 'gcr'): the
 * normalised activations are written once as COMPACT fp16 hi / lo slot planes,
 * [channel chunk][N][D][H][W] x 16 B (one slot = 8 channels of a voxel; x2, when
 * given, is nearest-upsampled and concatenated after x's channels, each source
 * padded to 8 channels), and the convolution kernel stages a patch / slab of those
 * planes WITH its halo in shared memory through TMA tile loads (the box starts one
 * voxel outside the volume; out-of-bounds slots are zero-filled, so the padding
 * never exists in HBM), addressing all 27 taps in place through no-swizzle UMMA
 * descriptors (no im2col).  rf_halo_act_bytes = size of ONE of hi / lo.
 * rf_tc_conv3d_halo_supported tells whether an item shape fits shared memory and
 * TMEM for this layer; callers use rf_tc_conv3d_fwd otherwise.  D, H, W are the
 * INPUT extents; pad = 1 ('same', the U-Nets) or 0 ('valid', the conv patch
 * encoders of model/retrieval.py: output extents D-2).  interior_only is ignored
 * (kept for ABI stability: the planes had a halo in the first version).
 * y is fp32 channels-last [N,Do,Ho,Wo,Cout] or NCDHW. */
size_t rf_halo_act_bytes(int N, int D, int H, int W, int C1, int C2, int pad);
int rf_cl_norm_split_halo(const float* x, int C1, const float* x2, int C2, const float* gn_mu, const float* gn_a,
                          const float* gn_beta, void* hi, void* lo, int N, int D, int H, int W, int pad, float scale,
                          int interior_only, void* stream);
size_t rf_tc_conv_halo_weight_image_bytes(int Cout, int C1, int C2);
int rf_tc_conv_halo_weight_image(const float* w, int Cout, int C1, int C2, float scale, void* image, void* stream);
int rf_tc_conv3d_halo_supported(int N, int D, int H, int W, int Cout, int C1, int C2, int pad);
int rf_tc_conv3d_halo_geometry(int N, int D, int H, int W, int Cout, int C1, int C2, int pad, int* out8);
int rf_tc_conv3d_halo_debug_read(long long* out64); /* tuning aid: phase timestamps of CTA 0's first items */
int rf_tc_conv3d_halo_fwd(const void* hi, const void* lo, const void* weight_image, const float* bias, float* y, int N, int D,
                          int H, int W, int pad, int Cout, int C1, int C2, int act, float slope, float out_scale,
                          int out_ncdhw, void* stream);

/* Stride-2 'valid' 3x3x3 layers (the down-sampling Conv3d of the conv patch encoders, model/retrieval.py:4-28,187-275)
 * on the same kernel: parity planes of rf_cl_split_parity_planes ([chunk][parity][n][D/2][H/2][W/2] slots, sized by
 * rf_halo_s2_act_bytes), weight image of rf_tc_conv_halo_weight_image; the item's block is staged as its 8 parity
 * sub-blocks.  D, H, W: INPUT extents. */
size_t rf_halo_s2_act_bytes(int N, int D, int H, int W, int C1);
int rf_cl_split_parity_planes(const float* x, int C1, void* hi, void* lo, int N, int D, int H, int W, float scale, void* stream);
int rf_tc_conv3d_halo_s2_supported(int N, int D, int H, int W, int Cout, int C1);
int rf_tc_conv3d_halo_s2_fwd(const void* hi, const void* lo, const void* weight_image, const float* bias, float* y, int N, int D,
                             int H, int W, int Cout, int C1, int act, float slope, float out_scale, int out_ncdhw, void* stream);

/* W-pair variant for small Cout (2 Cout <= 128; model/unet.py:79-100 SingleConv and the 'valid' Conv3d layers of
 * model/retrieval.py:4-388): one GEMM row = the output voxels (w, w + 1), N = 2 Cout; planes W-de-interleaved
 * [chunk][w parity][n][d][h][w/2] (rf_cl_norm_split_halo_wp, sized by rf_halo_act_bytes).  D, H, W: INPUT extents.
 * _supported: 0 = no, 1 = runs, 2 = runs and the item cost model rates it faster than rf_tc_conv3d_halo_fwd. */
int rf_cl_norm_split_halo_wp(const float* x, int C1, const float* x2, int C2, const float* gn_mu, const float* gn_a,
                             const float* gn_beta, void* hi, void* lo, int N, int D, int H, int W, int pad, float scale, void* stream);
size_t rf_tc_conv_halo_wp_weight_image_bytes(int Cout, int C1, int C2);
int rf_tc_conv_halo_wp_weight_image(const float* w, int Cout, int C1, int C2, float scale, void* image, void* stream);
int rf_tc_conv3d_halo_wp_supported(int N, int D, int H, int W, int Cout, int C1, int C2, int pad);
/* Same convolution with MaxPool3d(2) of the activated output taken in the epilogue (model/unet.py:210-253: an encoder
 * level whose full-resolution output only feeds the next level's pooling): y is [N, D/2, H/2, W/2, Cout] channels-last.
 * 'same' padding, Cout % 16 == 0, even extents. */
int rf_tc_conv3d_halo_wp_pool_supported(int N, int D, int H, int W, int Cout, int C1, int C2);
int rf_tc_conv3d_halo_wp_pool_fwd(const void* hi, const void* lo, const void* weight_image, const float* bias, float* y, int N, int D,
                                  int H, int W, int Cout, int C1, int C2, int act, float slope, float out_scale, void* stream);
/* First convolution of a DoubleConv (model/unet.py:103-144) whose epilogue applies the SECOND SingleConv's GroupNorm
 * (model/unet.py:79-100) and writes that layer's operand planes straight from the accumulators: per-sample statistics of
 * the activated output, normalisation, x scale2, fp16 hi / lo split.  out_hi / out_lo: rf_halo_act_bytes(N, Do, Ho, Wo,
 * Cout, 0, 1) bytes each, in rf_cl_norm_split_halo's layout (out_wp = 0) or rf_cl_norm_split_halo_wp's (out_wp = 1).
 * D, H, W: INPUT extents.  Needs Cout <= 64 and an item shape with one whole sample per item (_supported). */
int rf_tc_conv3d_halo_gn_supported(int N, int D, int H, int W, int pad, int Cout, int C1, int C2, int groups2);
int rf_tc_conv3d_halo_gn_fwd(const void* hi, const void* lo, const void* weight_image, const float* bias, int N, int D, int H, int W,
                             int pad, int Cout, int C1, int C2, int act, float slope, float out_scale, const float* gn2_w,
                             const float* gn2_b, int groups2, float eps2, float scale2, void* out_hi, void* out_lo, int out_wp,
                             void* stream);
int rf_tc_conv3d_halo_wp_geometry(int N, int D, int H, int W, int Cout, int C1, int C2, int pad, int* out16, double* scores2);
int rf_tc_conv3d_halo_wp_fwd(const void* hi, const void* lo, const void* weight_image, const float* bias, float* y, int N, int D,
                             int H, int W, int pad, int Cout, int C1, int C2, int act, float slope, float out_scale, void* stream);

/* Fused front of the first DoubleConv of a 'gcr' U-Net encoder on 16^3 single-channel samples (model/unet.py:79-144,
 * the retrieval U-Net's patches, model/refinement.py:64-73): GroupNorm(1,1) -> Conv3d(1,8,3,p=1) -> ReLU ->
 * GroupNorm(groups2, 8) -> x scale -> fp16 hi / lo operand planes of rf_tc_conv3d_halo_fwd (wp = 0) or
 * rf_tc_conv3d_halo_wp_fwd (wp = 1), sized by rf_halo_act_bytes(N,16,16,16,8,0,1).  One persistent CTA per sample;
 * the 8-channel fp32 activations never reach HBM.  x [N,16,16,16], gn2_w / gn2_b [8]: device pointers.  The 218 scalars
 * the FMAs read as kernel-parameter constants come from the HOST: conv_w_host [8][27] (the Conv3d weight [8,1,3,3,3]) and
 * gn1_w / gn1_b by value.  No stream synchronisation (graph-capturable; re-capture after a weight update). */
int rf_unet_front16_fwd(const float* x, float gn1_w, float gn1_b, float eps1, const float* conv_w_host, const float* gn2_w,
                        const float* gn2_b, int groups2, float eps2, float scale, void* hi, void* lo, int N, int wp, void* stream);

/* Single-input-channel layers on the same kernel (the first Conv3d of every patch
 * encoder, model/retrieval.py:4-388, kernel edge 3 or 5, no padding; the first
 * SingleConv of the U-Nets, model/unet.py:79-100, 3^3 'same'): the slot of voxel
 * (d,h,p) holds the 8 consecutive values x[d,h,p-pad .. p-pad+7] of its line
 * ("W-run", zero outside the line, GroupNorm(1 group) applied first when gn_* are
 * given), so one 16-byte K chunk carries all kw taps of a (kd,kh) line and a K = 16
 * MMA step two lines: 5 steps for 3^3, 13 for 5^3.  D, H, W: INPUT extents; planes
 * [N][D][H][Wo] slots, Wo = W + 2 pad - KS + 1; y fp32 channels-last [N,Do,Ho,Wo,Cout]. */
size_t rf_wrun_act_bytes(int N, int D, int H, int W, int KS, int pad);
int rf_cl_norm_split_wrun(const float* x, const float* gn_mu, const float* gn_a, const float* gn_beta, void* hi, void* lo,
                          int N, int D, int H, int W, int KS, int pad, float scale, void* stream);
size_t rf_tc_conv_wrun_weight_image_bytes(int Cout, int KS);
int rf_tc_conv_wrun_weight_image(const float* w, int Cout, int KS, float scale, void* image, void* stream);
int rf_tc_conv3d_wrun_supported(int N, int D, int H, int W, int Cout, int KS, int pad);
int rf_tc_conv3d_wrun_fwd(const void* hi, const void* lo, const void* weight_image, const float* bias, float* y, int N, int D,
                          int H, int W, int KS, int pad, int Cout, int act, float slope, float out_scale, int out_ncdhw,
                          void* stream);

/* First layers (single input channel: the TSDF / occupancy volume): direct
 * convolution, one thread per output voxel, filter bank in shared memory,
 * optional GroupNorm(1 group) on the input (gn_mu/gn_a per sample, gn_beta
 * scalar), bias + activation fused.  x [N,Di,Hi,Wi] fp32, w [Cout,1,KS,KS,KS]
 * (the Conv3d weight as is), y fp32 channels-last [N,Do,Ho,Wo,Cout], Cout <= 32. */
int rf_conv3d_cin1_cl_fwd(const float* x, const float* w, const float* bias, const float* gn_mu, const float* gn_a,
                          const float* gn_beta, float* y, int N, int Di, int Hi, int Wi, int Cout, int KS, int stride, int pad,
                          int act, float slope, void* stream);

/* Pointwise head (model/refinement.py:55-57: the decoder's final Conv3d(nf, 1, 1) + bias + Tanh) on a channels-last
 * volume x [n_voxels, C] -> y [n_voxels] (= NCDHW with one channel); w [C] is the Conv3d weight, bias one float. */
int rf_cl_pointwise_head(const float* x, const float* w, const float* bias, float* y, long n_voxels, int C, int act, float slope,
                         void* stream);

/* ---- a5 + a9  fused query encoder --------------------------------------- */

/* model/retrieval.py:64-84 Patch04.forward (+Patch05/Patch04V2: any ReLU MLP)
 * followed by util/retrieval.py:66 normalisation, activations kept on chip.
 *   x [M, widths[0]] rows; wt[l] = layers[2l].weight.T [widths[l], widths[l+1]];
 *   out [M, widths[n_layers]] unit rows.  n_layers <= 8.  Hidden activations
 *   live in `workspace` (rf_mlp_encode_workspace_bytes), never in caller tensors. */
size_t rf_mlp_encode_workspace_bytes(long M, const int* widths_host, int n_layers);
int rf_mlp_encode_fwd(const float* x, const float* const* wt_host, const float* const* bias_host,
                      const int* widths_host, int n_layers, int l2_normalize, float* out, long M, void* workspace,
                      size_t workspace_bytes, void* stream);

/* ---- a10-a11  exact kNN, merge, demotion -------------------------------- */

/* Canonical exact rule (replaces util/retrieval.py:92 flann nn_index(q, 2K)):
 *   d(q,x) = sum_i (double(q_i) - double(x_i))^2, i ascending, each op rounded
 *   once (no fma); result = k smallest under ascending (d, global row id).
 * bank [n_rows, D] is one shard whose first row has global id row_offset.
 * out_idx [Q,k] int32 global ids, out_d [Q,k] fp64.  D must be 64.
 * method: 0 = auto, 1 = exact fp64 sweep, 2 / 3 = tcgen05 candidate pass (2:
 * one fp16 GEMM; 3: bf16 hi/lo split, K = 192; fp32 accumulators in TMEM) that
 * keeps 16 (k <= 8) or 32 candidates per query and bank slice, then the
 * canonical fp64 re-rank, a proof that no rejected row can enter or tie the
 * top-k, and the fp64 sweep for the queries whose proof fails - enqueued
 * unconditionally and sized by a device-side counter, so no method ever
 * synchronises `stream` (the call can be captured in a CUDA graph).  All
 * methods return identical results; k <= 32. */
size_t rf_knn_workspace_bytes(long Q, long n_rows, int k, int method);
int rf_knn_l2_topk(const float* bank, long n_rows, long row_offset, const float* q, long Q, int D, int k, int method,
                   int* out_idx, double* out_d, void* workspace, size_t workspace_bytes, void* stream);
/* Prepared banks.  The reference loads its FLANN index once per worker and
 * queries it many times (util/retrieval.py:81-92); here the bank's tensor-core
 * operand image (scan order + swizzled 16-bit tiles + |x|^2 range) is built
 * once by rf_knn_bank_prepare into a caller-owned buffer of
 * rf_knn_bank_image_bytes and reused by every rf_knn_l2_topk_prepared call.
 * q_sample (optional, n_sample rows): a sample of the queries the image will
 * serve; their mean direction orders the scan (rows that score high for a
 * typical query first - fewer list insertions, exactness unaffected).  NULL =
 * rows stay in bank order, the right choice when later queries are unknown.  rf_knn_bank_method resolves method 0 for a bank of n_rows (1 = too
 * small for the tensor-core path: use rf_knn_l2_topk). */
int rf_knn_bank_method(long n_rows, int method);
size_t rf_knn_bank_image_bytes(long n_rows, int method);
size_t rf_knn_bank_scratch_bytes(long n_rows, int method);
int rf_knn_bank_prepare(const float* bank, long n_rows, int method, const float* q_sample, long n_sample, void* image,
                        size_t image_bytes, void* scratch, size_t scratch_bytes, void* stream);
size_t rf_knn_prepared_workspace_bytes(long Q, long n_rows, int k, int method);
int rf_knn_l2_topk_prepared(const float* bank, long n_rows, long row_offset, const void* image, int method, const float* q,
                            long Q, int D, int k, int* out_idx, double* out_d, void* workspace, size_t workspace_bytes,
                            void* stream);
/* Diagnostics of the last method-2/3 call that used `workspace` (synchronises
 * the stream): number of queries that needed the exact re-check and the
 * largest observed error of the tensor-core score over all re-ranked
 * candidates (validates the bound). */
int rf_knn_tc_stats(const void* workspace, int* n_unproven, float* max_score_err, void* stream);
/* Merge S sorted candidate lists per query (shards of one GPU sweep or the
 * all-gathered per-rank lists): parts_idx/parts_d [S,Q,k] -> [Q,k]
 * under the same (d, id) order. */
int rf_knn_merge(const int* parts_idx, const double* parts_d, int S, long Q, int k, int* out_idx, double* out_d,
                 void* stream);
/* util/retrieval.py:93-100: gather database[:,0:7] for the 2K hits, move hits
 * whose scene (meta[:,0]) equals query_scene[q] behind the others (stable),
 * keep K.  query_scene[q] < 0 disables demotion for that query.
 * meta [N,7] fp32; out_rows [Q,K,8] = [scene,x0,x1,y0,y1,z0,z1,(float)d];
 * out_idx [Q,K] int32 (may be NULL). */
int rf_knn_demote_rows(const int* idx2k, const double* d2k, const float* meta, const int* query_scene, long Q, int K2,
                       int K, float* out_rows, int* out_idx, void* stream);

/* ---- a12  compose ------------------------------------------------------- */

/* util/retrieval.py:145-164 create_retrieval_from_mapping for non-overlapping
 * patches (patch_stride == patch_size, patched_scene_dataset.py:113-115).
 *   rows [n_chunks*P, K, 8] mapping rows in patch order; dst_extents [P,6] int32
 *   (unpadded destination extents inside a chunk); scene_store [S,sx,sy,sz]
 *   unpadded train targets; out [n_chunks,K,cx,cy,cz];
 *   out[c,k,dst] = scene_store[row.scene][x0:x1,y0:y1,z0:z1] * ratio, or
 *   trunc * ratio when row.scene == -1 (the sentinel row, :160-161); a row with
 *   scene < -1 leaves its destination block untouched (cells nobody owns when
 *   the host has resolved overlapping patches, :156, into single-owner cells).
 *   norm_div != 0 additionally maps v -> (v - norm_sub) / norm_div, the
 *   dataloader's retrieval normalisation (patched_scene_dataset.py:133). */
int rf_compose_gather(const float* rows, const int* dst_extents, const float* scene_store, float* out, int n_chunks,
                      int P, int K, int n_scenes, const int scene_size[3], const int chunk_size[3], float trunc,
                      float ratio, float norm_sub, float norm_div, void* stream);

/* rf_compose_gather followed by Unfold3D(16, 1) of its result (trainer/train_refinement.py:110-111), in one pass:
 * out is [n_chunks, K, P, ex, ey, ez], every destination block stored contiguously in patch order.  The caller
 * guarantees that the P blocks have equal extents and tile the chunk in Unfold3D's patch order ((x, y, z) row-major);
 * a block whose extents do not fit that layout is skipped. */
int rf_compose_gather_patches(const float* rows, const int* dst_extents, const float* scene_store, float* out, int n_chunks,
                              int P, int K, int n_scenes, const int scene_size[3], const int chunk_size[3], float trunc,
                              float ratio, float norm_sub, float norm_div, void* stream);

/* ---- a14  patch attention ------------------------------------------------ */

/* model/attention.py:141-157 PatchedAttentionBlock.forward with
 * AttentionBlock.forward :84-113 (g = o = Identity, blend or additive).
 *   x_back [B,nf,S,S,S], x_retr [B*K,nf,S,S,S] -> out [B,nf,S,S,S]
 *   theta_wt/phi_wt: 4 transposed Linear weights [in,out] each, *_b biases;
 *   theta_img/phi_img: NULL, or the 4 rf_tc_mlp_weight_image buffers of each MLP -
 *   then each MLP is ONE fused tensor-core launch (rf_tc_mlp_fwd) and *_wt is unused;
 *   mode 0: softmax(32*E^3*4 * s); mode 1: hard Gumbel arg-max of 25*s + noise
 *   (noise [B*R^3, K], required); workspace from rf_attention_workspace_bytes. */
size_t rf_attention_workspace_bytes(int B, int nf, int S, int E, int K);
int rf_attention_fuse_fwd(const float* x_back, const float* x_retr, const float* const* theta_wt_host,
                          const float* const* theta_b_host, const float* const* phi_wt_host,
                          const float* const* phi_b_host, const void* const* theta_img_host,
                          const void* const* phi_img_host, const float* gumbel_noise, float* out, int B, int nf, int S,
                          int E, int K, int normalize, int mode, int blend, void* workspace, size_t workspace_bytes,
                          void* stream);
/* The same call with the two re-indexing passes around it folded in (train_refinement.py:108-120 runs
 * Fold3D(P, S/P, nf) on the retrieval U-Net's patches right before the attention, model/refinement.py:48-61 the
 * decoder right after it):
 *   patch_grid P > 1: x_retr is the UN-folded patch batch [B*K*P^3, nf, S/P,S/P,S/P] (model/attention.py:170-176
 *     Fold3D's input, patches in (PX,PY,PZ) order); Fold3D followed by Unfold3D(E) is a permutation of rows, which
 *     the score / blend stage applies as index arithmetic.
 *   out_channels_last != 0: out is [B,S,S,S,nf] (the decoder's channels-last operand) instead of [B,nf,S,S,S].
 *   output_mapping != NULL: attn_no_output_mapping = False (model/attention.py:56-57,95,108): g and o are 1x1x1
 *     convolutions around the weighted sum; both are linear and per voxel, so the caller passes their composition
 *     {Wo Wg [nf,nf] row-major, Wo bg [nf], bo [nf]} (device pointers in a host array); needs out_channels_last == 0.
 * P <= 1, out_channels_last == 0 and output_mapping == NULL is rf_attention_fuse_fwd. */
int rf_attention_fuse_patched_fwd(const float* x_back, const float* x_retr, const float* const* theta_wt_host,
                                  const float* const* theta_b_host, const float* const* phi_wt_host,
                                  const float* const* phi_b_host, const void* const* theta_img_host,
                                  const void* const* phi_img_host, const float* gumbel_noise, float* out, int B, int nf,
                                  int S, int E, int K, int normalize, int mode, int blend, int patch_grid,
                                  int out_channels_last, const float* const* output_mapping, void* workspace,
                                  size_t workspace_bytes, void* stream);
/* model/attention.py:132-139 get_features: theta(unfold(x)), phi(unfold(t)),
 * any(occupancy) per sub-patch.  x,t [B,nf,S,S,S]; occ [B,1,S,S,S] uint8;
 * x_feat,p_feat [B*R^3,32]; occ_any [B*R^3] uint8. */
int rf_attention_features(const float* x, const float* t, const uint8_t* occ, const float* const* theta_wt_host,
                          const float* const* theta_b_host, const float* const* phi_wt_host,
                          const float* const* phi_b_host, const void* const* theta_img_host,
                          const void* const* phi_img_host, float* x_feat, float* p_feat, uint8_t* occ_any, int B, int nf,
                          int S, int E, int normalize, void* workspace, size_t workspace_bytes, void* stream);

/* ---- callers either side of the path ------------------ */

/* dataset/patched_scene_dataset.py:139-146 compute_normals: pad by 1 with pad_val (the target truncation), the
 * three 3x3x3 Sobel cross-correlations (:194-196), n / sqrt(|n|^2 + 1e-5).  x [B,1,D,H,W] -> out [B,3,D,H,W]. */
int rf_sobel_normals(const float* x, float* out, int B, int D, int H, int W, float pad_val, void* stream);
/* util/metrics.py:15-16 (IoU), :66 (Precision), :83 (Recall): per sample the integer sums
 * counts[b] = {sum(pred & target), sum(pred | target), sum(pred), sum(target)} of two bool volumes
 * (one byte per voxel, non-zero = True); counts [B,4] uint64 (zeroed by the call). */
int rf_occupancy_counts(const uint8_t* pred, const uint8_t* target, int B, long voxels_per_sample, unsigned long long* counts,
                        void* stream);
/* external/ChamferDistancePytorch chamfer3D (un-vendored submodule; call site util/metrics.py:46
 * `dist1, dist2, _, _ = cham_loss(points_target, points_pred)`): for every point of a [na,3] the squared distance
 * to its nearest neighbour in b [nb,3] and that neighbour's index (first minimal index).  Call twice for both
 * directions.  d = fma(dz,dz, fma(dy,dy, dx*dx)) in fp32. */
int rf_chamfer_nn(const float* a, int na, const float* b, int nb, float* dist, int* idx, void* stream);
/* model/loss.py:48-69 NTXentLoss.forward(zis, zjs, iou_matrix=None): zis, zjs [N,C] (C <= 128), iou_matrix NULL or
 * [2N,2N]; cosine != 0 selects the cosine similarity (eps 1e-8), else the dot product; loss [1].
 * workspace: rf_ntxent_workspace_bytes(N). */
size_t rf_ntxent_workspace_bytes(int N);
int rf_ntxent_fwd(const float* zis, const float* zjs, int N, int C, const float* iou_matrix, float temperature, float sig_scale,
                  float sig_shift, int cosine, float* loss, void* workspace, size_t workspace_bytes, void* stream);

/* ---- 8f.3  backward passes (training_step_full, trainer/train_refinement.py:74-89) ---------- */

/* Adjoint of rf_conv3d_fwd with respect to the filter: dw [Cout, Cin, KS,KS,KS] (the nn.Conv3d / nn.Linear layout,
 * overwritten) = sum over all output positions of dz [N,Cout,Do,Ho,Wo] x the normalised, zero-padded, virtually
 * upsampled + concatenated input (same arguments as the forward).  KS = 1 with D = H = W = 1 is nn.Linear's
 * weight gradient.  The adjoint with respect to the INPUT is rf_conv3d_fwd itself on the flipped, transposed filter. */
int rf_conv3d_wgrad(const float* x, const float* x2, int C2, const float* gn_mu, const float* gn_a, const float* gn_beta,
                    const float* dz, float* dw, int N, int Cin, int Di, int Hi, int Wi, int Cout, int KS, int stride, int pad,
                    void* stream);
/* dz = dy * act'(pre-activation), written through the saved OUTPUT y (ReLU, LeakyReLU(slope > 0), tanh, none). */
int rf_act_bwd(const float* dy, const float* y, float* dz, long n, int act, float slope, void* stream);
/* out[c] = sum over n, v of x [N,C,V] (bias gradients; V = 1 for Linear layers). */
int rf_channel_sum(const float* x, int N, int C, long V, float* out, void* stream);
/* GroupNorm backward (model/unet.py:60-64 in front of every conv): g = dL/d(normalised input) of the virtual input
 * concat(x [N,C-C2,...], up2(x2 [N,C2,...])); gn_mu / gn_rstd [N,C] per-channel copies of the group statistics.
 * dx (fine part), dx2 (coarse part: its 8 children summed), dgamma / dbeta [C] (may be NULL). */
size_t rf_gn_bwd_workspace_bytes(int N, int C);
int rf_gn_bwd(const float* x, const float* x2, int C2, const float* g, const float* gn_mu, const float* gn_rstd,
              const float* gamma, int N, int C, int D, int H, int W, int groups, float* dx, float* dx2, float* dgamma,
              float* dbeta, void* workspace, void* stream);
/* Nearest x2 upsampling backward for channels [C1, C) of g [N,C,D,H,W] -> dx2 [N,C-C1,D/2,H/2,W/2] (no GroupNorm). */
int rf_upsample2_bwd(const float* g, int N, int C, int C1, int D, int H, int W, float* dx2, void* stream);
/* MaxPool3d(2) backward: the first maximum of each window (scan order d, h, w) receives dy. */
int rf_maxpool3d_2_bwd(const float* x, const float* dy, float* dx, int N, int C, int D, int H, int W, void* stream);
/* model/attention.py:84-113 on precomputed theta / phi features (rows in Unfold3D order: xf [R,32], xu [R,V]; pf, pu
 * in (b, k, r) order), and its adjoint with respect to the features and the row vectors.  The differentiable path
 * runs theta / phi layer by layer so that autograd keeps the activations; rf_attention_fuse_fwd is the fused
 * inference call. */
int rf_attention_epilogue_fwd(const float* xf, const float* pf, const float* xu, const float* pu, const float* noise, float* orows,
                              long R, int rp3, int K, int V, int normalize, int mode, int blend, float sharp, void* stream);
int rf_attention_epilogue_bwd(const float* xf, const float* pf, const float* xu, const float* pu, const float* noise,
                              const float* dout, float* dxf, float* dpf, float* dxu, float* dpu, long R, int rp3, int K, int V,
                              int normalize, int mode, int blend, float sharp, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RF_B200_H */
