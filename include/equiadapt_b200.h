/* equiadapt_b200 -- C ABI of the B200-native canonicalization hot path.
 *
 * The reference (arnab39/equiadapt) is pure Python/PyTorch and has no FFI of its own; the
 * "interface" each entry point replaces is therefore the reference Python function (or the
 * chain of library kernels it dispatches), cited as file:line relative to /root/reference.
 * INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous row-major float32 unless stated;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is
 *     enqueued on it, nothing synchronises the host;
 *   - return value: 0 = ok, <0 = argument error (EQB_ERR_*), >0 = cudaError_t of the launch;
 *     eqb_last_error() returns a thread-local message for the last non-zero return;
 *   - a discrete group element is an int32 index g in [0,|G|): g <  N : rotation g*360/N degrees,
 *                                                               g >= N : rotation (g-N)*360/N and a reflection
 *     (N = num_rotations, |G| = N or 2N), the order the reference uses for one-hot vectors
 *     (equiadapt/images/canonicalization/discrete_group.py:110-133).
 */
#ifndef EQUIADAPT_B200_H
#define EQUIADAPT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EQB_ABI_VERSION 1
#define EQB_API __attribute__((visibility("default")))

#define EQB_ERR_INVALID (-1)     /* bad shape / size / enum */
#define EQB_ERR_UNSUPPORTED (-2) /* valid in the reference, not covered by this build (message says what) */

#define EQB_REP_SCALAR 0
#define EQB_REP_REGULAR 1

EQB_API int eqb_abi_version(void);
EQB_API const char *eqb_last_error(void);

/* ---- a3  pre-network transform ------------------------------------------------------------
 * CenterCrop + antialiased bilinear Resize (torchvision transforms on tensors -> ATen
 * _upsample_bilinear2d_aa).  Replaces discrete_group.py:174-188 (transforms built at :73-92).
 * x (B,C,H,W) -> y (B,C,out_h,out_w); the crop window is [top,top+crop_h) x [left,left+crop_w). */
EQB_API int eqb_crop_resize_aa(const float *x, float *y, int B, int C, int H, int W, int top, int left,
                       int crop_h, int crop_w, int out_h, int out_w, void *stream);
/* The same transform, additionally leaving y_absmax[b] = max |y[b]| (B floats): the per-IMAGE operand scale of the conv
 * stack (eqb_gconv_stack_run_scaled), reduced inside the resize kernel instead of by a second pass over y. */
EQB_API int eqb_crop_resize_aa_absmax(const float *x, float *y, float *y_absmax, int B, int C, int H, int W, int top,
                                      int left, int crop_h, int crop_w, int out_h, int out_w, void *stream);

/* ---- a4 / a5  filter orbits ---------------------------------------------------------------
 * Lift: w (Cout,Cin,k,k) -> orbit (Cout*|G|, Cin, k, k), channel = o*|G|+g.
 * Replaces RotationEquivariantConvLift.get_rotated_weights /
 * RotoReflectionEquivariantConvLift.get_rotoreflected_weights,
 * custom_group_equivariant_layers.py:62-90, :169-199. */
EQB_API int eqb_lift_filter_orbit(const float *w, float *orbit, int cout, int cin, int k, int num_rotations,
                          int reflect, void *stream);
/* Regular: w (Cout,Cin,|G|,k,k) -> orbit (Cout*|G|, Cin*|G|, k, k).
 * Replaces get_rotated_permuted_weights / get_rotoreflected_permuted_weights, :298-334, :461-507. */
EQB_API int eqb_regular_filter_orbit(const float *w, float *orbit, int cout, int cin, int k, int num_rotations,
                             int reflect, void *stream);

/* ---- a4..a6  fused group-conv stack -> group activations ----------------------------------
 * CustomEquivariantNetwork.forward (custom_equivariant_networks.py:80-93):
 *   lift conv k x k (valid) -> [ReLU -> 1x1 regular group conv] x (L-1) -> mean over (Cout,H',W').
 * x (B,Cin,H,W); lift_w (Cout,Cin,k,k); reg_w[l] (Cout,Cout,|G|,1,1) for l = 0..L-2 given as a
 * HOST array of device pointers; biases (Cout) each, a NULL bias pointer means "no bias".
 * act (B,|G|).  `workspace` is device scratch of at least eqb_gconv_stack_workspace_bytes() bytes. */
EQB_API int64_t eqb_gconv_stack_workspace_bytes(int B, int cin, int H, int W, int cout, int k, int num_rotations,
                                                int reflect, int num_layers);
EQB_API int eqb_gconv_stack_forward(const float *x, int B, int cin, int H, int W, const float *lift_w,
                                    const float *lift_b, const float *const *reg_w, const float *const *reg_b,
                                    int cout, int k, int num_rotations, int reflect, int num_layers, float *act,
                                    void *workspace, int64_t workspace_bytes, void *stream);
/* The same in two steps, for frozen weights: pack ONCE (filter orbits as GEMM operands, expanded biases,
 * folded last layer; the reference rebuilds its orbits on every forward, custom_group_equivariant_layers.py:103,
 * :349-351), then run per batch.  `packed` needs eqb_gconv_stack_packed_bytes(); `scratch` needs
 * eqb_gconv_stack_workspace_bytes() - eqb_gconv_stack_packed_bytes() for the batch at hand.
 * `last_bias` = bias of the last layer (Cout) or NULL (ignored when num_layers == 1). */
EQB_API int64_t eqb_gconv_stack_packed_bytes(int cin, int cout, int k, int num_rotations, int reflect, int num_layers);
EQB_API int eqb_gconv_stack_pack(const float *lift_w, const float *lift_b, const float *const *reg_w,
                                 const float *const *reg_b, int cin, int cout, int k, int num_rotations, int reflect,
                                 int num_layers, void *packed, int64_t packed_bytes, void *stream);
EQB_API int eqb_gconv_stack_run(const float *x, int B, int cin, int H, int W, const void *packed,
                                const float *last_bias, int cout, int k, int num_rotations, int reflect,
                                int num_layers, float *act, void *scratch, int64_t scratch_bytes, void *stream);
/* eqb_gconv_stack_run with x_absmax[b] = max |x[b]| (B floats) supplied by the producer of x (eqb_crop_resize_aa_absmax)
 * instead of being reduced here.  Either way the fp16 operand split is scaled PER IMAGE: the activations of a sample do not
 * depend on the other samples of the batch (custom_equivariant_networks.py:80-93 is per-sample). */
EQB_API int eqb_gconv_stack_run_scaled(const float *x, const float *x_absmax, int B, int cin, int H, int W,
                                       const void *packed, const float *last_bias, int cout, int k, int num_rotations,
                                       int reflect, int num_layers, float *act, void *scratch, int64_t scratch_bytes,
                                       void *stream);

/* eqb_gconv_stack_run_scaled (x_absmax may be NULL: reduced here) with the group pool / select of eqb_group_pool_select
 * (a9 + a13, basecanonicalization.py:221-256, :290-311; discrete_group.py:94-135) fused into its finish kernel: same
 * outputs, bit for bit, as calling eqb_group_pool_select on `act`, without the extra launch.  `packed` is written (a
 * 4-byte ticket the kernel resets itself): one network instance must not run on two streams at once. */
EQB_API int eqb_gconv_stack_run_select(const float *x, const float *x_absmax, int B, int cin, int H, int W, void *packed,
                                       const float *last_bias, int cout, int k, int num_rotations, int reflect,
                                       int num_layers, float *act, int32_t *idx, float *rotation, float *reflection,
                                       float *onehot, float *stats, void *scratch, int64_t scratch_bytes, void *stream);

/* ---- a7  e2cnn-style conv stack with EXPANDED filters -> group activations ------------------
 * ESCNNEquivariantNetwork.forward in eval() (escnn_networks.py:93-117; modules built at :66-91):
 *   [ conv2d k x k (valid) + bias -> scale * . + shift (InnerBatchNorm, eval) -> ReLU ] x (L-1)
 *   -> conv2d k x k + bias -> reshape (B, Cout, |G|, H', W') -> mean over (Cout, H', W').
 * filters[l]: (Cout*|G|, Cin_l, k, k) with Cin_0 = cin and Cin_l = Cout*|G| after (what e2cnn caches as
 * R2Conv.filter); biases[l]: (Cout*|G|) (R2Conv.expanded_bias) or NULL; scales[l] / shifts[l]: per-channel affine of
 * layer l < L-1 or NULL.  All four are HOST arrays of L device pointers (`biases`, `scales`, `shifts` may be NULL).
 * act (B, |G|); `workspace` >= eqb_conv_stack_workspace_bytes() bytes of device scratch, 16-byte aligned. */
EQB_API int64_t eqb_conv_stack_workspace_bytes(int B, int cin, int H, int W, int cout, int k, int num_group,
                                               int num_layers);
EQB_API int eqb_conv_stack_forward(const float *x, int B, int cin, int H, int W, const float *const *filters,
                                   const float *const *biases, const float *const *scales,
                                   const float *const *shifts, int cout, int k, int num_group, int num_layers,
                                   float *act, void *workspace, int64_t workspace_bytes, void *stream);

/* Diagnostics: the tcgen05 stack kernel bounds every pipeline wait (~2 s); if one expires the kernel traps instead of
 * hanging and leaves {flag, block, warp, barrier id, parity} here (host memory).  Returns flag (0 = no stall seen). */
EQB_API int eqb_debug_last_stall(int *out5);
/* Diagnostics: device_buffer (tiles x 64 int64, zeroed by the caller) receives clock64() stamps of the pipeline events of the
 * first `tiles` pair-tiles of cluster 0 in every following CTA-pair stack launch (slot table: csrc/gconv_stack_tc.cu);
 * null switches it off.  The traced kernel is a separate instantiation: production launches carry no instrumentation. */
EQB_API int eqb_debug_stack_trace(void *device_buffer, int tiles);

/* ---- a9 + a13  group pool / select + prior statistic --------------------------------------
 * act (B,|G|) -> idx int32 (B) = first arg-max, rotation (B) in degrees, reflection (B) 0/1 (may be
 * NULL), onehot (B,|G|) (may be NULL), stats[5] = { sum_b CE(act_b, class 0), sum_b [idx_b == 0], B, and this
 * batch's own means stats[0]/B, stats[1]/B } (the first three are what is all-reduced across ranks).
 * Replaces groupactivations_to_groupelementonehot (common/basecanonicalization.py:221-256, eval branch),
 * groupactivations_to_groupelement (discrete_group.py:94-135) and the two reductions of
 * get_prior_regularization_loss / get_identity_metric (basecanonicalization.py:290-311). */
EQB_API int eqb_group_pool_select(const float *act, int B, int num_rotations, int reflect, int32_t *idx,
                          float *rotation, float *reflection, float *onehot, float *stats, void *stream);

/* ---- a10  canonicalize: inverse group action on the input image ---------------------------
 * y = crop(rotate(flip?(pad_replicate(x)), -rotation)) for C != 1, bare zero-fill rotate for C == 1.
 * Replaces discrete_group.py:207-215 (pad :62-66, crop :67-71).  x, y (B,C,H,W). */
EQB_API int eqb_warp_canonicalize(const float *x, float *y, const int32_t *idx, int B, int C, int H, int W,
                          int num_rotations, int reflect, void *stream);

/* ---- a11  invert_canonicalization: forward group action on a feature map ------------------
 * Replaces get_action_on_image_features + roll_by_gather (equiadapt/images/utils.py:8-94).
 * f, out (B,C,H,W); rep = EQB_REP_SCALAR | EQB_REP_REGULAR (C % |G| == 0 required for regular). */
EQB_API int eqb_warp_invert(const float *f, float *out, const int32_t *idx, int B, int C, int H, int W,
                    int num_rotations, int reflect, int rep, void *stream);

/* ---- N3 (partial)  gradient of the two warps above with respect to their image argument --------
 * grad_in = W(g)^T grad_out, the adjoint of the (linear) warp: what autograd needs to train a prediction network whose
 * output goes through invert_canonicalization, or to differentiate canonicalize(x) with respect to x.  The group
 * element is not differentiated.  mode 0: eqb_warp_canonicalize, 1: eqb_warp_invert scalar, 2: eqb_warp_invert regular.
 * (The reference gets this from torch autograd through kornia's grid_sample: discrete_group.py:207-215,
 * images/utils.py:54-79.) */
EQB_API int eqb_warp_adjoint(const float *grad_out, float *grad_in, const int32_t *idx, int B, int C, int H, int W,
                             int num_rotations, int reflect, int mode, void *stream);

/* ---- N3  gradient of the two warps with respect to the GROUP ELEMENT -----------------------------
 * in: the warp's input image (B,C,H,W); grad_out: d loss / d (warp output).  grad_rotation (B): d loss / d rotation in
 * DEGREES (the element the reference feeds kornia rotate, discrete_group.py:110-133,213; images/utils.py:57);
 * grad_reflection (B, may be NULL; ignored unless reflect): d loss / d reflection indicator of the flip blend
 * (discrete_group.py:209-210; images/utils.py:59-64).  mode as in eqb_warp_adjoint.  Both outputs are overwritten.
 * With these two the straight-through one-hot of basecanonicalization.py:239-251 trains the canonicalization network. */
EQB_API int eqb_warp_element_grad(const float *in, const float *grad_out, const int32_t *idx, int B, int C, int H, int W,
                          int num_rotations, int reflect, int mode, float *grad_rotation, float *grad_reflection,
                          void *stream);

/* Host-only, no GPU work: the channel shift the reference derives for rotation index r by
 * `(angle / 360.0 * num_rotations).long()` in float32 (images/utils.py:67,:28); equals r for
 * power-of-two N, may truncate to r-1 otherwise (reference quirk reproduced by eqb_warp_invert). */
EQB_API int eqb_regular_roll_shift(int r, int num_rotations);

/* ---- a12  optimisation-based variant -------------------------------------------------------
 * Orbit expand: x (B,C,h,w) -> out (|G|*B, C, out, out), group-major, each member =
 * crop_out(hflip?(rotate(pad_replicate(x, pad), -deg_g))).  Replaces group_augment /
 * rotate_and_maybe_reflect (discrete_group.py:387-427).  C == 1: no pad/crop (zero fill). */
EQB_API int eqb_orbit_expand(const float *x, float *out, int B, int C, int h, int w, int pad, int out_size,
                     int num_rotations, int reflect, void *stream);
/* ---- N4  evaluation-time group orbit -------------------------------------------------------
 * x (B,C,H,W) -> out (|G|, B, C, H, W): for every group element (rotations by torch.linspace(0,360,N+1)[:-1], then
 * the same rotations of the mirrored image) CenterCrop(H,W)(torchvision rotate NEAREST, zero fill ([hflip](
 * Pad(ceil(0.4 H), edge)(x)))).  Replaces the per-element pad / hflip / rotate / crop loop of
 * GroupInference.get_group_element_wise_logits (examples/images/classification/inference_utils.py:97-122).
 * Bit-exact with torchvision's float32 affine grid + grid_sample(nearest). */
EQB_API int eqb_orbit_rotate_nearest(const float *x, float *out, int B, int C, int H, int W, int num_rotations,
                                     int reflect, void *stream);
/* Cosine similarity to the reference vector + (|G|,B)->(B,|G|) transpose: vec (|G|*B, V), ref (V),
 * act (B,|G|).  Replaces discrete_group.py:475-481. */
EQB_API int eqb_cosine_group_activations(const float *vec, const float *ref, float *act, int B, int num_group,
                                 int V, void *stream);
/* N3: its backward (torch autograd through F.cosine_similarity + transpose in the reference, discrete_group.py:475-481):
 * dact (B,|G|) -> dvec (|G|*B, V) and / or dref (V); either output may be NULL; both are overwritten. */
EQB_API int eqb_cosine_group_activations_backward(const float *vec, const float *ref, const float *dact, float *dvec,
                                          float *dref, int B, int num_group, int V, void *stream);

/* ---- a14..a17  frames ----------------------------------------------------------------------
 * Gram-Schmidt on the three rows of v (B,3,3) -> R (B,3,3).  modified = 0: common/utils.py:22-51;
 * modified = 1: nbody/canonicalization/euclidean_group.py:139-157. */
EQB_API int eqb_gram_schmidt3(const float *v, float *R, int B, int modified, void *stream);
/* y = R x for clouds x (B,3,N): pointcloud/canonicalization/continuous_group.py:66-81. */
EQB_API int eqb_so3_apply(const float *x, const float *R, float *y, int B, int N, void *stream);
/* loc_c = (loc - t) R^T, vel_c = vel R^T, one (R,t) per ROW: euclidean_group.py:108-124.  M rows. */
EQB_API int eqb_e3_apply(const float *loc, const float *vel, const float *R, const float *t, float *loc_c,
                 float *vel_c, int M, void *stream);
/* y = x R + t: euclidean_group.py:126-137. */
EQB_API int eqb_e3_invert(const float *x, const float *R, const float *t, float *y, int M, void *stream);
/* stats[5] = { sum (R - I)^2, B*d*d, 0, mse = stats[0]/stats[1], 1 - mse } for R (B,d,d): the reductions of
 * ContinuousGroupCanonicalization.get_prior_regularization_loss / get_identity_metric
 * (common/basecanonicalization.py:390-430). */
EQB_API int eqb_prior_stats_continuous(const float *R, int B, int d, float *stats, void *stream);

/* ---- N2  continuous rotations / roto-reflections of images -----------------------------------
 * y(dst) = bilinear sample of x at  src = c + A (dst - c), taps replicate-clamped into the image while inside the
 * extent padded by `pad` pixels and zero beyond; refl (B floats 0/1, may be NULL): the source is mirrored
 * horizontally first.  mats (B,2,2): mats_forward = 1 -> M with dst - c = M (src - c) (kornia warp_affine's 2x2 block;
 * A = M^-1), 0 -> A itself (F.affine_grid's theta).  (cx, cy) = c in pixel coordinates of the un-padded image.
 * Replaces ContinuousGroupImageCanonicalization.canonicalize (images/canonicalization/continuous_group.py:162-210:
 * flip blend, Pad(edge), K.geometry.warp_affine, CenterCrop) and the warp of
 * OptimizedSteerableImageCanonicalization.group_augment (:362-412: Pad, affine_grid + grid_sample, CenterCrop). */
EQB_API int eqb_warp_affine(const float *x, float *y, const float *mats, const float *refl, int mats_forward, int B,
                            int C, int H, int W, int pad, double cx, double cy, void *stream);
/* N3: gradient of eqb_warp_affine (canonicalize direction) with respect to the sampling map and the flip blend.
 * theta (B,6): destination -> source affine map in un-padded pixel coordinates, xs = t0 xd + t1 yd + t2,
 * ys = t3 xd + t4 yd + t5 (the inverse of the reference's 2x3 warp_affine matrix, shifted by the pad; built by the
 * caller, whose autograd then carries the result back to the network: continuous_group.py:183-208).  refl (B, may be
 * NULL): 0/1 flip indicator used by the forward; grad_theta (B,6) and grad_refl (B, may be NULL) are overwritten. */
EQB_API int eqb_warp_affine_grad(const float *in, const float *grad_out, const float *theta, const float *refl, int B, int C,
                         int H, int W, int pad, float *grad_theta, float *grad_refl, void *stream);

/* ---- N1  frame-predicting vector-neuron networks (eval mode) --------------------------------
 * VNSmall.forward (pointcloud/canonicalization_networks/equivariant_networks.py:128-150; knn :15-33,
 * get_graph_feature_cross :36-76; VNLinearLeakyReLU / VNBatchNorm vector_neuron_layers.py:210-324), pooling "mean":
 * x (B,3,N) -> out (B,3,3), the three equivariant vectors eqb_gram_schmidt3 turns into a frame.
 * `params`: eqb_vnsmall_param_count() floats = the raw tensors of the reference module, concatenated in the order
 *   conv_pos.{map_to_feat.weight, map_to_dir.weight, batchnorm.bn2d.{weight, bias, running_mean, running_var}},
 *   conv1.{map_to_feat.weight, map_to_dir.weight, batchnorm.bn1d.{...}}, bn1.bn1d.{...},
 *   conv2.{map_to_feat.weight, map_to_dir.weight, batchnorm.bn1d.{...}};   bn_eps = the batch norms' eps.
 * `workspace` >= eqb_vnsmall_workspace_bytes(B, N) bytes of device scratch, 8-byte aligned (a cloud is split over
 * several CTAs when the batch alone cannot fill the GPU). */
EQB_API int eqb_vnsmall_param_count(void);
EQB_API int64_t eqb_vnsmall_workspace_bytes(int B, int N);
EQB_API int eqb_vnsmall_forward(const float *x, int B, int N, const float *params, int n_knn, float bn_eps, float *out,
                                void *workspace, int64_t workspace_bytes, void *stream);
/* VNDeepSets.forward (nbody/canonicalization_networks/custom_equivariant_networks.py:106-172; VNDeepSetLayer
 * :175-252; VNLeakyReLU / VNSoftplus custom_group_equivariant_layers.py:7-99) for S systems of 5 consecutive rows:
 * loc, vel (5S,3), charges (5S) (vel / charges may be NULL when no feature uses them), edges (2,E) int64 = rows
 * (source, destination) as the reference passes them; every edge must stay inside one system (bad_edges, if not
 * NULL, receives the number that do not and are ignored).  Feature channels: canonical location, then velocity
 * (feat_v), angular momentum (feat_a), charge-weighted location (feat_c) - canon_feature "p" / "pv" / "pva" /
 * "pvc" / "pvac".  nonlinearity 0 = relu, 1 = leakyrelu(0.2), 2 = softplus.  -> rot_vectors (5S,3,3), translation (5S,3).
 * `params`: eqb_vndeepsets_param_count() floats: per layer {identity_linear.weight, .bias, pooling_linear.weight,
 * .bias, nonlinear_function.map_to_dir.weight}, then output_layer.{weight (4 x H), bias (4)}.
 * `workspace` >= eqb_vndeepsets_workspace_bytes(S) bytes. */
EQB_API int eqb_vndeepsets_param_count(int in_dim, int hidden, int num_layers);
EQB_API int64_t eqb_vndeepsets_workspace_bytes(int S);
EQB_API int eqb_vndeepsets_forward(const float *loc, const float *vel, const float *charges, const int64_t *edges,
                                   int64_t E, int S, const float *params, int in_dim, int hidden, int num_layers,
                                   int feat_v, int feat_a, int feat_c, int nonlinearity, int layer_pool_mean,
                                   int final_pool_mean, int canon_translation, float *rot_vectors, float *translation,
                                   void *workspace, int64_t workspace_bytes, int32_t *bad_edges, void *stream);

/* ---- N3  training kernels of the group-conv canonicalization network ------------------------------
 * What torch autograd derives in the reference for CustomEquivariantNetwork (custom_equivariant_networks.py:80-93)
 * through F.conv2d / ReLU / mean (custom_group_equivariant_layers.py:104-112, :352-361) and the filter-orbit
 * construction (:62-90, :169-199, :298-334, :461-507).  The fused inference stack keeps no activations; training runs
 * layer by layer on these, saving every post-ReLU feature map.  All tensors NCHW fp32, valid k x k convolution,
 * w (N, cin, k, k).
 *   eqb_conv2d_forward:      y = [relu](conv2d(x, w) + bias) [zeroed where mask <= 0]; bias / mask may be NULL.  With w
 *                            transposed, k = 1 and mask = the saved input feature map it is the data gradient of a 1x1
 *                            layer through the preceding ReLU.
 *   eqb_conv2d_weight_grad:  dw[n, ci, ky, kx] = sum_{b,y,x} dy[b,n,y,x] x[b,ci,y+ky,x+kx]  (dw overwritten).
 *   eqb_plane_sums:          out[row] = sum_p x[row, p] (bias gradients, spatial sums).
 *   eqb_group_mean_backward: dy[b, o*|G|+g, p] = dact[b,g] / (cout * P), the backward of the final mean.
 *   eqb_*_filter_orbit_adjoint: gradient of the base weights from the gradient of the expanded filters (the orbit
 *                            maps are linear: the adjoint scatters the same bilinear taps; dw overwritten). */
EQB_API int eqb_conv2d_forward(const float *x, const float *w, const float *bias, const float *mask, float *y, int B,
                       int cin, int H, int W, int N, int k, int relu, void *stream);
EQB_API int eqb_conv2d_weight_grad(const float *dy, const float *x, float *dw, int B, int cin, int H, int W, int N, int k,
                           void *stream);
/* The same two with the per-image maxima max |x[b]| handed along (B floats each, device; any of them may be NULL).  The
 * 1x1 layers with 256 channels run on the tensor pipe with a per-image power-of-two operand scale (csrc/gconv_stack_tc.cu,
 * namespaces pw / wg); a caller that chains layers passes the maxima the previous call produced (y_absmax) instead of
 * paying a pass over a 555 MB feature map per call.  x_absmax / dy_absmax must be >= the true maxima (an under-estimate
 * overflows the fp16 split) and should be within 2x of them (an over-estimate costs precision).  dy_rowsum (N floats or
 * NULL) receives sum_{b,y,x} dy[b,n,y,x] -- the bias gradient of the layer -- from the same pass over dy. */
EQB_API int eqb_conv2d_forward_scaled(const float *x, const float *w, const float *bias, const float *mask, float *y, int B,
                              int cin, int H, int W, int N, int k, int relu, const float *x_absmax, float *y_absmax,
                              void *stream);
EQB_API int eqb_conv2d_weight_grad_scaled(const float *dy, const float *x, float *dw, int B, int cin, int H, int W, int N,
                                  int k, const float *dy_absmax, const float *x_absmax, float *dy_rowsum, void *stream);
EQB_API int eqb_plane_sums(const float *x, int64_t rows, int64_t P, float *out, void *stream);
EQB_API int eqb_group_mean_backward(const float *dact, float *dy, int B, int cout, int num_group, int64_t P, void *stream);
EQB_API int eqb_lift_filter_orbit_adjoint(const float *dorbit, float *dw, int cout, int cin, int k, int num_rotations,
                                  int reflect, void *stream);
EQB_API int eqb_regular_filter_orbit_adjoint(const float *dorbit, float *dw, int cout, int cin, int k, int num_rotations,
                                     int reflect, void *stream);

/* ---- N3  backward of the frame path (a14..a17) ----------------------------------------------------
 * What torch autograd derives in the reference through gram_schmidt (common/utils.py:22-51) / modified Gram-Schmidt
 * (nbody euclidean_group.py:139-157), the point-cloud bmm (pointcloud continuous_group.py:77-79) and the n-body row
 * products (euclidean_group.py:114-122, :133-136), so that a torch frame-predicting network trains through the
 * canonicalizers.  Output pointers may be NULL where a gradient is not wanted (dv excepted); outputs are overwritten. */
EQB_API int eqb_gram_schmidt3_backward(const float *v, const float *dR, float *dv, int B, int modified, void *stream);
EQB_API int eqb_so3_apply_backward(const float *x, const float *R, const float *dy, float *dx, float *dR, int B, int N,
                           void *stream);
EQB_API int eqb_e3_apply_backward(const float *loc, const float *vel, const float *R, const float *t, const float *dloc_c,
                          const float *dvel_c, float *dloc, float *dvel, float *dR, float *dt, int M, void *stream);
EQB_API int eqb_e3_invert_backward(const float *x, const float *R, const float *dy, float *dx, float *dR, float *dt, int M,
                           void *stream);

#ifdef __cplusplus
}
#endif
#endif /* EQUIADAPT_B200_H */
