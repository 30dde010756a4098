/*
 * libunetb200 — C ABI of the B200 (sm_100a) U-Net training / inference hot path.
 *
 * Drop-in boundary for SaurabhIndi/unet-segmentation: everything below sits UNDER the reference's
 * `UNet` nn.Module (models/unet_model.py:65-146) and `WeightedCrossEntropyLoss`
 * (utils/losses.py:6-57) and under `get_instance_masks` (utils/metrics.py:42-72). The reference has
 * no FFI of its own (it dispatches to ATen/cuDNN through torch.nn); the ctypes binding a
 * maintainer would add is shown in INTEGRATION.md and shipped in unet_segmentation_b200/_lib.py.
 *
 * Conventions
 *   - every function returns 0 on success or a negative ub_status; ub_last_error() gives the text
 *     (thread-local). No C++ exception crosses this boundary. There is NO CPU fallback: without a
 *     CUDA device every compute entry point fails with UB_ERR_CUDA.
 *   - all pointers are DEVICE pointers unless stated otherwise; `stream` is a cudaStream_t passed
 *     as void*; no entry point synchronises the device or allocates after plan creation.
 *   - activations inside the library are NHWC bf16; the tensors that cross the boundary keep the
 *     reference's layouts: images / logits NCHW fp32, parameters and gradients in torch layout fp32,
 *     targets int64, masks uint8, instance labels uint16.
 */
#ifndef UNET_B200_H_
#define UNET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ub_status {
    UB_OK = 0,
    UB_ERR_CUDA = -1,        /* CUDA runtime / driver failure (includes "no device") */
    UB_ERR_ARG = -2,         /* invalid argument */
    UB_ERR_UNSUPPORTED = -3, /* configuration outside the supported envelope (e.g. base_channels = 96) */
    UB_ERR_TMAP = -4,        /* TMA descriptor encoding failed */
    UB_ERR_NOMEM = -5
} ub_status;

/* NHWC bf16 view: element strides, channel stride 1. A centre crop (reference
 * models/unet_model.py:88-102) or a channel slice of a concat buffer (:131-143) is just a view. */
typedef struct ub_view {
    const void* ptr;
    int32_t N, H, W, C;
    int64_t sN, sH, sW;
} ub_view;

const char* ub_last_error(void);
int ub_version(void);
int ub_device_sm_count(void);

/* ------------------------------------------------------------------------------------------------
 * Network executor ("plan"): replaces UNet.forward (models/unet_model.py:105-146) and its autograd
 * backward for a fixed input shape. levels = 5 and base_channels = 64 reproduce the reference
 * (:73-85); other values are the optional trailing keyword arguments of the drop-in constructor.
 *
 * Parameter order (ub_plan_num_params entries) = reference named_parameters() order:
 *   inc.{conv0.w, conv0.b, bn0.w, bn0.b, conv1.w, conv1.b, bn1.w, bn1.b}, down1..down{L-1} (same 8),
 *   up1..up{L-1}.{up.w, up.b, conv0.w, conv0.b, bn0.w, bn0.b, conv1.w, conv1.b, bn1.w, bn1.b},
 *   outc.{w, b}.
 * BN buffer order (ub_plan_num_bn entries): the BatchNorm2d modules in the same traversal order.
 * ---------------------------------------------------------------------------------------------- */
typedef struct ub_plan ub_plan;

int ub_plan_create(ub_plan** plan, int N, int n_channels, int H, int W, int base_channels,
                   int levels, int n_classes, int training);
/* bilinear != 0: the Up blocks of UNet(..., bilinear=True) (models/unet_model.py:40-43,78-81):
 * nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) keeps the channel count, the
 * first convolution of the block then takes C_skip + C_prev channels (1536 -> 512 ...), and the
 * up.{w, b} entries drop out of the parameter order (8 parameters per decoder stage). */
int ub_plan_create_ex(ub_plan** plan, int N, int n_channels, int H, int W, int base_channels,
                      int levels, int n_classes, int training, int bilinear);
int ub_plan_destroy(ub_plan* plan);
int ub_plan_out_hw(const ub_plan* plan, int* out_h, int* out_w);
int ub_plan_num_params(const ub_plan* plan);
int ub_plan_num_bn(const ub_plan* plan);
int64_t ub_plan_param_numel(const ub_plan* plan, int index);
int64_t ub_plan_device_bytes(const ub_plan* plan);

/* Bind the torch-owned fp32 master parameters / BN buffers (device pointers, kept by reference). */
int ub_plan_bind_params(ub_plan* plan, const float* const* params, int count);
int ub_plan_bind_bn_buffers(ub_plan* plan, float* const* running_mean, float* const* running_var,
                            int64_t* const* num_batches_tracked, int count);
/* Per-layer BatchNorm2d hyper-parameters (host arrays of ub_plan_num_bn() floats): `momentum` of the
 * running-statistics update and `eps` (reference models/unet_model.py:12,16 uses the nn.BatchNorm2d
 * defaults 0.1 / 1e-5, which are also the plan's defaults). */
int ub_plan_set_bn_config(ub_plan* plan, const float* momentum, const float* eps, int count);
/* Refresh the packed bf16 operand caches from the bound masters (after optimizer.step() /
 * load_state_dict()). */
int ub_plan_pack_weights(ub_plan* plan, void* stream);

/* x: [N][n_channels][H][W] fp32; logits: [N][n_classes][out_h][out_w] fp32.
 * training plan: batch statistics, running-stat update (momentum 0.1, unbiased variance),
 * activations saved for backward. eval plan: running statistics folded into the conv epilogues;
 * `mask` (optional, may be NULL; [N][out_h][out_w] uint8) receives 255 where logit1 > logit0,
 * i.e. softmax(logits)[:,1] > 0.5 (reference scripts/predict.py:85-92).
 * An eval plan called repeatedly with the SAME x / logits / mask pointers (and unchanged parameter
 * binding) captures its launch sequence into a CUDA graph on the second call and replays it from
 * then on (UB_EVAL_GRAPH=0 disables); ub_plan_graph_replays counts the replays. */
int ub_plan_forward(ub_plan* plan, const float* x, float* logits, uint8_t* mask, void* stream);
int64_t ub_plan_graph_replays(const ub_plan* plan);

/* Backward, split into stages so that a data-parallel caller can all-reduce the gradients of a
 * finished stage while the next one runs. Stage 0 = outc + last up block, then the remaining up
 * blocks deepest-last, then the down blocks bottom-up, last stage = inc.
 * grads[i] receives d(loss)/d(param i) (fp32, torch layout, overwritten) for the parameters of the
 * stage only; ub_plan_stage_params reports which. dlogits: [N][n_classes][out_h][out_w] fp32.
 * x must still be the tensor given to the last ub_plan_forward. */
int ub_plan_num_stages(const ub_plan* plan);
int ub_plan_stage_params(const ub_plan* plan, int stage, int* first_param, int* num_params);
int ub_plan_backward_stage(ub_plan* plan, int stage, const float* dlogits, float* const* grads,
                           void* stream);

/* Fused optimizer step: torch.optim.SGD semantics (scripts/train.py:97,131 use
 * lr 1e-4, momentum 0.99) applied to the bound fp32 parameters in place, with the plan's packed bf16
 * operands refreshed in the same pass (no ub_plan_pack_weights needed afterwards for THIS plan).
 * grads / momentum_bufs: ub_plan_num_params() device pointers in parameter order; momentum_bufs may
 * be NULL when momentum == 0; first_step != 0 initialises the buffers with the gradient. */
int ub_plan_sgd_step(ub_plan* plan, const float* const* grads, float* const* momentum_bufs, float lr,
                     float momentum, float dampening, float weight_decay, int nesterov,
                     int first_step, void* stream);

/* Measurement support (bench.py): total number of kernels the library has launched so far, and
 * optional in-step timing: while enabled, every kernel group the plan launches is bracketed by two
 * CUDA events on the launching stream; collect() synchronises them and returns, per kernel class
 * (ub_plan_profile_class_name), the summed duration [ms], algorithmic FLOPs / bytes and the number
 * of timed groups. Arrays hold ub_plan_profile_classes() entries. */
/* Weight gradients run on an internal low-priority stream, concurrently with the BatchNorm-backward
 * / data-gradient chain of the caller's stream. mode 0: off (everything on the caller's stream);
 * 1 (default): joined at the end of every ub_plan_backward_stage call, so a stage's gradients are
 * complete in stream order when it returns (data-parallel all-reduce hooks); 2: joined only at the
 * end of the last stage (single-GPU training). */
int ub_plan_set_overlap(ub_plan* plan, int mode);
/* Makes `stream` wait for the weight gradients issued so far (mode 2 + a communication stream: the
 * all-reduce of a stage waits for that stage's weight gradients while the backward itself goes on). */
int ub_plan_join_side(ub_plan* plan, void* stream);
int64_t ub_launch_count(void);
int ub_plan_profile_enable(ub_plan* plan, int on);
int ub_plan_profile_classes(void);
const char* ub_plan_profile_class_name(int cls);
int ub_plan_profile_collect(ub_plan* plan, double* ms, double* flops, double* bytes, int* launches);

/* ------------------------------------------------------------------------------------------------
 * WeightedCrossEntropyLoss.forward (utils/losses.py:29-57) + its gradient, one pass.
 * Strides are in elements; inputs may be non-contiguous views (scripts/train.py:118-126).
 * dlogits (optional): contiguous [N][C][H][W], = d(loss)/d(logits) for grad_output == 1.
 * err_flag (device int, caller zeroes it): set to 1 if a target is outside [0, C) and != -100.
 * ---------------------------------------------------------------------------------------------- */
int64_t ub_wce_workspace_floats(void);
int ub_wce_forward(const float* logits, const int64_t logit_strides[4], const int64_t* targets,
                   const int64_t target_strides[3], const float* weight_maps,
                   const int64_t weight_strides[3], int N, int C, int H, int W, float* loss,
                   float* dlogits, float* workspace, int* err_flag, void* stream);
/* out[i] = in[i] * (*scalar) — applies the upstream grad_output (a device scalar). */
int ub_scale_by_device_scalar(const float* in, const float* scalar, float* out, int64_t n,
                              void* stream);

/* ------------------------------------------------------------------------------------------------
 * Device-side input pipeline. One pass over a compactly transferred batch:
 *   image_f32[N][1][H][W] = images_u8 / 255            (ToTensor, utils/dataset.py:92-94)
 *   target[N][out_h][out_w] = (labels > 0) as int64      (utils/dataset.py:96-105), centre-cropped
 *   weight[N][out_h][out_w] = (float)weight_maps         (utils/dataset.py:110), centre-cropped like
 *                             center_crop_tensor (scripts/train.py:39-51,118-126)
 * labels: uint8 (label_bytes 1) or uint16 (2), may be NULL; weight_maps: float32 (4) or float64
 * (8), may be NULL. W must be a multiple of 4. Bit-exact against the torch ops it replaces.
 * ---------------------------------------------------------------------------------------------- */
int ub_prepare_batch(const uint8_t* images_u8, const void* labels, int label_bytes,
                     const void* weight_maps, int weight_bytes, int N, int H, int W, int out_h,
                     int out_w, float* image_f32, int64_t* target, float* weight, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Weight-map generation: calculate_weight_map(mask, w0, sigma) of
 * scripts/preprocess_data.py:17-77 for a batch of instance-label masks, on the device.
 *   labels      [N][H][W] uint8 (label_bytes 1) or uint16 (2) instance ids, 0 = background
 *   weight_maps [N][H][W] float64 (weight_bytes 8: what the reference stores as .npy) or float32
 *               (4: torch.from_numpy(map).float(), utils/dataset.py:110)
 *   counts      N uint32 of device scratch (foreground pixel counts; zeroed by the call)
 * The reference's border term is identically w0 (its distance maps vanish for every mask, SURVEY
 * F6), so the map is float32(1 / class fraction) + w0 per pixel; bit-exact against the stored maps.
 * The result feeds ub_prepare_batch (weight_maps argument).
 * ---------------------------------------------------------------------------------------------- */
int ub_weight_map(const void* labels, int label_bytes, int N, int H, int W, double w0, double sigma,
                  void* weight_maps, int weight_bytes, uint32_t* counts, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Elastic deformation: elastic_deform_image_and_mask(image, mask, alpha, sigma,
 * random_state) of utils/augmentations.py:4-39 for a batch, on the device, bit-exact against the
 * reference (scipy.ndimage.gaussian_filter(mode="constant") + map_coordinates(mode="reflect")).
 *   images      [N][H][W] uint8 (may be NULL)      -> images_out uint8: bilinear, rounded half up
 *   labels      [N][H][W] uint8 / uint16 (may be NULL) -> labels_out, nearest neighbour;
 *               label_out_bytes 1 with uint16 input wraps modulo 256 like the reference's
 *               mask.astype(np.uint8) (utils/dataset.py:93)
 *   noise       [2][N][H][W] float64 uniform [0, 1) draws: the reference's random_state.rand(*shape)
 *               for dx (first draw, augmentations.py:27) then for dy (:28) of every sample
 *   taps        2 * radius + 1 float64 Gaussian weights (only the first radius + 1 are read):
 *               exp(-0.5 / sigma^2 * x^2) / sum for x = -radius … radius, radius = int(4 sigma + .5)
 *               (scipy's _gaussian_kernel1d; computed by the caller so that exp is numpy's)
 *   workspace   ub_elastic_workspace_bytes(N, H, W) bytes of device scratch
 * ---------------------------------------------------------------------------------------------- */
int64_t ub_elastic_workspace_bytes(int N, int H, int W);
int ub_elastic_deform(const uint8_t* images, const void* labels, int label_bytes, int N, int H, int W,
                      const double* noise, const double* taps, int radius, double alpha,
                      uint8_t* images_out, void* labels_out, int label_out_bytes, void* workspace,
                      void* stream);

/* ------------------------------------------------------------------------------------------------
 * get_instance_masks (utils/metrics.py:42-72): 8-connected labelling in raster order, components
 * smaller than min_size zeroed, ids not compacted, uint16 output. Bit-exact.
 * ---------------------------------------------------------------------------------------------- */
int64_t ub_ccl_workspace_bytes(int H, int W);
int ub_ccl_label(const uint8_t* mask, int H, int W, int min_size, uint16_t* labels, void* workspace,
                 void* stream);

/* ------------------------------------------------------------------------------------------------
 * Overlap-tile inference helpers (BASELINE configs[3]; the reference describes the strategy in prose
 * only — README.md:104-106, images/old readme unet.txt:73-87 — semantics in DESIGN.md §5).
 *   origins_yx  T pairs (row, column) of int32 on the DEVICE: top-left corners of the tiles' OUTPUT
 *               windows in image coordinates; a negative row marks an unused slot.
 * ub_extract_tiles: tiles[t][y][x] = image[reflect(row_t - margin + y)][reflect(col_t - margin + x)]
 *   (numpy 'reflect' mirroring, any extension length), image [H][W] fp32, tiles [T][tile_in][tile_in]
 *   fp32 = the (T,1,tile_in,tile_in) network input; tile_in % 4 == 0.
 * ub_stitch_tiles: full[row_t + y][col_t + x] = tiles[t][y][x] clipped to the H x W image; tiles
 *   [T][tile_out][tile_out] uint8 masks as ub_plan_forward writes them; tile_out % 4 == 0.
 * One launch per call for all T tiles; no host work per tile.
 * ---------------------------------------------------------------------------------------------- */
int ub_extract_tiles(const float* image, int H, int W, const int32_t* origins_yx, int T, int tile_in,
                     int margin, float* tiles, void* stream);
int ub_stitch_tiles(const uint8_t* tiles, const int32_t* origins_yx, int T, int tile_out,
                    uint8_t* full, int H, int W, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Operator-level entry points (teacher-forced per-layer parity tests). bf16 NHWC
 * views in, the same kernels the plan launches.
 * ---------------------------------------------------------------------------------------------- */
/* nn.Conv2d(k=3,p=0).weight [Co][Ci][3][3] fp32 -> forward operand wf [9][Co][Ci] bf16 and
 * data-gradient operand wd [9][Ci][Co] bf16 (tap-major; taps rotated by 180 degrees for wd).
 * wd may be NULL. */
int ub_op_pack_conv3x3(const float* w, int Co, int Ci, void* wf, void* wd, void* stream);
/* nn.ConvTranspose2d(k=2,s=2).weight [Ci][Co][2][2] -> wf [4*Co][Ci], wb [4][Ci][Co] (bf16);
 * bias [Co] -> bias4 [4*Co]. wb / bias / bias4 may be NULL. */
int ub_op_pack_convT(const float* w, int Ci, int Co, void* wf, void* wb, const float* bias,
                     float* bias4, void* stream);

int64_t ub_op_conv_stats_floats(int Co);
/* Valid 3x3 convolution over the channel concat [src0, src1] (src1 may be NULL).
 * epilogue 0: y = acc + bias (bf16) and per-channel (sum, sumsq) partials -> stats, launch
 *             geometry -> info[4] = {grid, n_tiles, BN, M} (host ints) for ub_op_bn_finalize;
 * epilogue 1: y = acc + bias;   epilogue 2: y = relu(acc*scale + shift). */
int ub_op_conv3x3_forward(const ub_view* src0, const ub_view* src1, const void* wf,
                          const float* bias, int Co, int epilogue, const float* scale,
                          const float* shift, void* y, float* stats, int* info, void* stream);
/* dx[N][H+2][W+2][Ci] = full correlation of dy[N][H][W][Co] with the rotated weights wd. */
/* Eval path of the LAST conv unit (64 output channels): 3x3 conv + folded BN affine + ReLU with the
 * 1x1 OutConv and the (z1 > z0) mask fused into the epilogue; the activation is never stored
 * (reference models/unet_model.py:11-17,56-63,145; scripts/predict.py:85-92).
 * head_w [n_classes][64], head_b [n_classes] fp32; logits NCHW fp32; mask (optional) u8 [N][H-2][W-2]. */
int ub_op_conv3x3_affine_relu_head(const ub_view* src0, const ub_view* src1, const void* wf,
                                   const float* scale, const float* shift, const float* head_w,
                                   const float* head_b, int n_classes, float* logits,
                                   uint8_t* mask, void* stream);
int ub_op_conv3x3_dgrad(const ub_view* dy, const void* wd, int Ci, void* dx, void* stream);
/* Data gradient whose epilogue also performs the REDUCE pass of the BatchNorm + ReLU backward of the
 * layer that receives dx (nn.BatchNorm2d + nn.ReLU backward, models/unet_model.py:12-13,16-17): y = that
 * layer's stored pre-BN output (bf16, the geometry of dx), scale / shift / mean = its batch-norm
 * statistics. partial: ub_op_conv_stats_floats(Ci) floats; info4 describes the partial rows for
 * ub_op_bn_relu_backward_fused, which then runs only the finalisation and the apply pass. */
int ub_op_conv3x3_dgrad_bnred(const ub_view* dy, const void* wd, int Ci, void* dx, const void* y,
                              const float* scale, const float* shift, const float* mean,
                              float* partial, int* info4, void* stream);
int ub_op_bn_relu_backward_fused(const void* y, int N, int H, int W, int C, const float* scale,
                                 const float* shift, const float* mean, const float* rstd,
                                 const ub_view* g, float* partial, const int* info4, float* dgamma,
                                 float* dbeta, void* dy, void* stream);
/* fp32 workspace (split-K partial tiles) of a weight gradient with `rows` = taps * C_in GEMM rows, `cols`
 * = C_out columns over `pixels` output pixels; covers both forms of the 64-output-channel 3x3 case (tap-major
 * 128 x 64 tiles and the "shifted" 128 x 192 tiles, csrc/igemm.cuh). A smaller buffer makes
 * ub_op_conv3x3_wgrad fall back to the tap-major form or fail with UB_ERR_ARG. */
int64_t ub_op_wgrad_workspace_floats(int rows, int cols, int64_t pixels);
/* dw[Co][C0+C1][3][3] fp32 = sum over pixels of x (concat of src0, src1) * dy[N][H-2][W-2][Co]. */
int ub_op_conv3x3_wgrad(const ub_view* src0, const ub_view* src1, const void* dy, int Co,
                        float* workspace, int64_t workspace_floats, float* dw, void* stream);
/* ConvTranspose2d(k=2,s=2): x[N][H][W][Ci] -> dst view [N][2H][2W][Co] (may be a channel slice). */
int ub_op_convT_forward(const ub_view* x, const void* wf, const float* bias4, int Co,
                        const ub_view* dst, void* stream);
/* dx[N][H][W][Ci] from dup view [N][2H][2W][Co]. */
int ub_op_convT_dgrad(const ub_view* dup, const void* wb, int Ci, void* dx, void* stream);
/* dw[Ci][Co][2][2] fp32. x is contiguous [N][H][W][Ci]. */
int ub_op_convT_wgrad(const ub_view* dup, const void* x, int Ci, float* workspace,
                      int64_t workspace_floats, float* dw, void* stream);

int ub_op_bn_finalize(const float* stats, const int* info, int C, const float* gamma,
                      const float* beta, float* running_mean, float* running_var,
                      int64_t* num_batches_tracked, float momentum, float eps, float* scale,
                      float* shift, float* mean, float* rstd, void* stream);
/* a = relu(y*scale + shift); pooled (optional) = 2x2/2 floor-mode max-pool of a; argmax (optional,
 * uint8 [N][H/2][W/2][C]) = position 0..3 of the first maximum of each window (torch tie rule). */
int ub_op_bn_apply_relu(const void* y, void* a, void* pooled, uint8_t* argmax, int N, int H, int W,
                        int C, const float* scale, const float* shift, void* stream);
/* BN-apply + ReLU of the last conv unit with the 1x1 output convolution fused (reference
 * models/unet_model.py:56-63,145): writes a (bf16 NHWC) and logits (fp32 NCHW [N][n_classes][H][W]).
 * C in {64, 128, 256}, n_classes <= 8. */
int ub_op_bn_apply_relu_head(const void* y, void* a, int N, int H, int W, int C, const float* scale,
                             const float* shift, int n_classes, const float* head_w,
                             const float* head_b, float* logits, void* stream);
int64_t ub_op_bn_bwd_workspace_floats(int C);
/* BN+ReLU backward. Upstream gradient of a: `g` (direct), or — when g == NULL — gathered from the
 * pooled-tensor gradient gp (2x2 max-pool backward, first arg-max) plus the skip-connection
 * gradient gs placed at (crop_h, crop_w) (gs may be NULL). argmax (optional) = the array saved by
 * ub_op_bn_apply_relu; without it the arg-max is recomputed from y. Outputs dgamma, dbeta (fp32) and
 * dy (bf16 [N][H][W][C], gradient of the conv output). */
int ub_op_bn_relu_backward(const void* y, int N, int H, int W, int C, const float* scale,
                           const float* shift, const float* mean, const float* rstd,
                           const ub_view* g, const ub_view* gp, const ub_view* gs, int crop_h,
                           int crop_w, const uint8_t* argmax, float* workspace, float* dgamma,
                           float* dbeta, void* dy, void* stream);

/* First convolution (fp32, C_in = n_channels): training forward = statistics + finalize + apply.
 * The backward takes the forward activation `a` (bf16 [N][H-2][W-2][Co]): for a single-channel
 * input the ReLU mask is read from it and dW / dgamma follow from patch moments in one pass. */
int64_t ub_op_first_conv_workspace_floats(int Co);
int ub_op_first_conv_forward(const float* x, int N, int Ci, int H, int W, const float* w,
                             const float* bias, int Co, const float* gamma, const float* beta,
                             float* running_mean, float* running_var,
                             int64_t* num_batches_tracked, float momentum, float eps,
                             float* workspace, float* scale, float* shift, float* mean,
                             float* rstd, void* a, void* stream);
/* Eval form: a = relu(conv(x, w) * scale + shift) with the caller's per-channel affine (BatchNorm
 * running statistics and the conv bias folded in), bf16 NHWC out. */
int ub_op_first_conv_affine_relu(const float* x, int N, int Ci, int H, int W, const float* w, int Co,
                                 const float* scale, const float* shift, void* a, void* stream);
int ub_op_first_conv_backward(const float* x, int N, int Ci, int H, int W, const float* w,
                              const float* bias, int Co, const float* scale, const float* shift,
                              const float* mean, const float* rstd, const ub_view* g,
                              const void* a, float* workspace, float* dgamma, float* dbeta,
                              float* dw, void* stream);

/* 1x1 head: a[N][H][W][K] bf16 -> logits NCHW fp32 (+ optional 2-class mask). */
int ub_op_head_forward(const void* a, int N, int H, int W, int K, int n_classes, const float* w,
                       const float* b, float* logits, uint8_t* mask, void* stream);
int64_t ub_op_head_bwd_workspace_floats(int K, int n_classes);
int ub_op_head_backward(const float* dlogits, const void* a, int N, int H, int W, int K,
                        int n_classes, const float* w, void* da, float* workspace, float* dw,
                        float* db, void* stream);
/* nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) (Up with bilinear=True,
 * models/unet_model.py:40-43): x view [N,H,W,C] -> out contiguous [N,2H,2W,C]; backward: g view
 * [N,2H,2W,C] -> dx contiguous [N,H,W,C] (exact adjoint, gather form). bf16, C % 8 == 0. */
int ub_op_upsample2x_forward(const ub_view* x, void* out, void* stream);
int ub_op_upsample2x_backward(const ub_view* g, void* dx, void* stream);
int ub_op_maxpool2(const void* a, void* pooled, int N, int H, int W, int C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNET_B200_H_ */
