// Operator-level C ABI (include/unet_b200.h): thin wrappers over the internal launchers, used by
// the teacher-forced per-layer parity tests and by the loss / post-processing modules.
#include "../../include/unet_b200.h"
#include "igemm.cuh"
#include "ub_internal.h"

using namespace ub;

static inline View to_view(const ub_view* v) {
    View r;
    r.ptr = v->ptr; r.N = v->N; r.H = v->H; r.W = v->W; r.C = v->C;
    r.sN = v->sN; r.sH = v->sH; r.sW = v->sW;
    return r;
}
static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
#define UB_REQUIRE(cond, msg)               \
    do {                                    \
        if (!(cond)) {                      \
            ub::set_last_error("%s", msg);  \
            return ub::UB_ERR_ARG;          \
        }                                   \
    } while (0)

extern "C" {

const char* ub_last_error(void) { return ub::last_error(); }
int ub_version(void) { return 100; }
int ub_device_sm_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        ub::set_last_error("no CUDA device available (libunetb200 has no CPU fallback)");
        return ub::UB_ERR_CUDA;
    }
    return ub::num_sms();
}

int64_t ub_wce_workspace_floats(void) { return (int64_t)wce_partial_floats(); }
int ub_wce_forward(const float* logits, const int64_t ls[4], const int64_t* targets,
                   const int64_t ts[3], const float* wm, const int64_t wstr[3], int N, int C, int H,
                   int W, float* loss, float* dlogits, float* workspace, int* err_flag,
                   void* stream) {
    UB_REQUIRE(logits && targets && wm && loss && workspace && err_flag, "wce: null pointer");
    WceDesc d;
    d.z = logits; d.t = (const long long*)targets; d.w = wm;
    for (int i = 0; i < 4; ++i) d.zs[i] = ls[i];
    for (int i = 0; i < 3; ++i) { d.ts[i] = ts[i]; d.ws[i] = wstr[i]; }
    d.N = N; d.C = C; d.H = H; d.W = W;
    return launch_wce(d, loss, dlogits, workspace, err_flag, S(stream));
}
int ub_scale_by_device_scalar(const float* in, const float* scalar, float* out, int64_t n,
                              void* stream) {
    return launch_scale_by_scalar(in, scalar, out, n, S(stream));
}

int ub_prepare_batch(const uint8_t* images_u8, const void* labels, int label_bytes,
                     const void* weight_maps, int weight_bytes, int N, int H, int W, int out_h,
                     int out_w, float* image_f32, int64_t* target, float* weight, void* stream) {
    return launch_prepare_batch(images_u8, labels, label_bytes, weight_maps, weight_bytes, N, H, W,
                                out_h, out_w, image_f32, (long long*)target, weight, S(stream));
}
int ub_weight_map(const void* labels, int label_bytes, int N, int H, int W, double w0, double sigma,
                  void* weight_maps, int weight_bytes, uint32_t* counts, void* stream) {
    return launch_weight_map(labels, label_bytes, N, H, W, w0, sigma, weight_maps, weight_bytes,
                             counts, S(stream));
}
int64_t ub_elastic_workspace_bytes(int N, int H, int W) {
    return N > 0 && H > 0 && W > 0 ? (int64_t)elastic_ws_bytes(N, H, W) : 0;
}
int ub_elastic_deform(const uint8_t* images, const void* labels, int label_bytes, int N, int H, int W,
                      const double* noise, const double* taps, int radius, double alpha,
                      uint8_t* images_out, void* labels_out, int label_out_bytes, void* workspace,
                      void* stream) {
    return launch_elastic(images, labels, label_bytes, N, H, W, noise, taps, radius, alpha, images_out,
                          labels_out, label_out_bytes, workspace, S(stream));
}
int64_t ub_ccl_workspace_bytes(int H, int W) { return (int64_t)ccl_ws_bytes(H, W); }
int ub_ccl_label(const uint8_t* mask, int H, int W, int min_size, uint16_t* labels, void* workspace,
                 void* stream) {
    UB_REQUIRE(mask && labels && workspace, "ccl: null pointer");
    return launch_ccl(mask, H, W, min_size, labels, workspace, S(stream));
}

int ub_extract_tiles(const float* image, int H, int W, const int32_t* origins_yx, int T, int tile_in,
                     int margin, float* tiles, void* stream) {
    return launch_extract_tiles(image, H, W, origins_yx, T, tile_in, margin, tiles, S(stream));
}
int ub_stitch_tiles(const uint8_t* tiles, const int32_t* origins_yx, int T, int tile_out,
                    uint8_t* full, int H, int W, void* stream) {
    return launch_stitch_tiles(tiles, origins_yx, T, tile_out, full, H, W, S(stream));
}

int ub_op_pack_conv3x3(const float* w, int Co, int Ci, void* wf, void* wd, void* stream) {
    return launch_pack_conv3x3(w, Co, Ci, (__nv_bfloat16*)wf, (__nv_bfloat16*)wd, S(stream));
}
int ub_op_pack_convT(const float* w, int Ci, int Co, void* wf, void* wb, const float* bias,
                     float* bias4, void* stream) {
    return launch_pack_convT(w, Ci, Co, (__nv_bfloat16*)wf, (__nv_bfloat16*)wb, bias, bias4,
                             S(stream));
}

int64_t ub_op_conv_stats_floats(int Co) { return (int64_t)igemm_stats_floats(Co); }
int ub_op_conv3x3_forward(const ub_view* src0, const ub_view* src1, const void* wf,
                          const float* bias, int Co, int epilogue, const float* scale,
                          const float* shift, void* y, float* stats, int* info, void* stream) {
    UB_REQUIRE(src0 && wf && y, "conv3x3_forward: null pointer");
    UB_REQUIRE(epilogue >= 0 && epilogue <= 2, "conv3x3_forward: bad epilogue");
    View v0 = to_view(src0), v1;
    if (src1) v1 = to_view(src1);
    IgemmEpilogue e;
    memset(&e, 0, sizeof(e));
    e.kind = epilogue; e.out = (__nv_bfloat16*)y; e.ldo = Co; e.bias = bias; e.scale = scale;
    e.shift = shift; e.stats = stats;
    IgemmLaunchInfo li;
    int r = launch_igemm(v0, src1 ? &v1 : nullptr, 0, -2, 1, 9, 3, (const __nv_bfloat16*)wf, Co, e,
                         &li, S(stream));
    if (r == 0 && info) { info[0] = li.grid; info[1] = li.n_tiles; info[2] = li.BN; info[3] = li.M; }
    return r;
}
int ub_op_conv3x3_affine_relu_head(const ub_view* src0, const ub_view* src1, const void* wf,
                                   const float* scale, const float* shift, const float* head_w,
                                   const float* head_b, int n_classes, float* logits,
                                   uint8_t* mask, void* stream) {
    UB_REQUIRE(src0 && wf && scale && shift && head_w && logits,
               "conv3x3_affine_relu_head: null pointer");
    View v0 = to_view(src0), v1;
    if (src1) v1 = to_view(src1);
    IgemmEpilogue e;
    memset(&e, 0, sizeof(e));
    e.kind = EPI_AFFINE_RELU_HEAD; e.ldo = 64; e.scale = scale; e.shift = shift;
    e.head_w = head_w; e.head_b = head_b; e.head_logits = logits; e.head_mask = mask;
    e.head_nc = n_classes;
    return launch_igemm(v0, src1 ? &v1 : nullptr, 0, -2, 1, 9, 3, (const __nv_bfloat16*)wf, 64, e,
                        nullptr, S(stream));
}
int ub_op_conv3x3_dgrad(const ub_view* dy, const void* wd, int Ci, void* dx, void* stream) {
    UB_REQUIRE(dy && wd && dx, "conv3x3_dgrad: null pointer");
    IgemmEpilogue e;
    memset(&e, 0, sizeof(e));
    e.kind = EPI_STORE; e.out = (__nv_bfloat16*)dx; e.ldo = Ci;
    return launch_igemm(to_view(dy), nullptr, -2, 0, 1, 9, 3, (const __nv_bfloat16*)wd, Ci, e,
                        nullptr, S(stream));
}
int ub_op_conv3x3_dgrad_bnred(const ub_view* dy, const void* wd, int Ci, void* dx, const void* y,
                              const float* scale, const float* shift, const float* mean,
                              float* partial, int* info, void* stream) {
    UB_REQUIRE(dy && wd && dx && y && scale && shift && mean && partial && info,
               "conv3x3_dgrad_bnred: null pointer");
    IgemmEpilogue e;
    memset(&e, 0, sizeof(e));
    e.kind = EPI_STORE_BNRED; e.out = (__nv_bfloat16*)dx; e.ldo = Ci;
    e.red_y = (const __nv_bfloat16*)y; e.scale = scale; e.shift = shift; e.red_mean = mean;
    e.stats = partial;
    IgemmLaunchInfo li;
    memset(&li, 0, sizeof(li));
    const int r = launch_igemm(to_view(dy), nullptr, -2, 0, 1, 9, 3, (const __nv_bfloat16*)wd, Ci, e,
                               &li, S(stream));
    if (r == 0) { info[0] = li.grid; info[1] = li.n_tiles; info[2] = li.BN; info[3] = li.M; }
    return r;
}
int ub_op_bn_relu_backward_fused(const void* y, int N, int H, int W, int C, const float* scale,
                                 const float* shift, const float* mean, const float* rstd,
                                 const ub_view* g, float* partial, const int* info, float* dgamma,
                                 float* dbeta, void* dy, void* stream) {
    UB_REQUIRE(y && g && partial && info && dgamma && dbeta && dy, "bn_relu_backward_fused: null pointer");
    BnBwdDesc d;
    memset(&d, 0, sizeof(d));
    d.y = (const __nv_bfloat16*)y; d.N = N; d.H = H; d.W = W; d.C = C;
    d.scale = scale; d.shift = shift; d.mean = mean; d.rstd = rstd;
    d.pool_skip = false; d.g = to_view(g);
    d.partial = partial; d.dgamma = dgamma; d.dbeta = dbeta; d.dy = (__nv_bfloat16*)dy;
    IgemmLaunchInfo li;
    memset(&li, 0, sizeof(li));
    li.grid = info[0]; li.n_tiles = info[1]; li.BN = info[2]; li.M = info[3];
    return launch_bn_bwd(d, S(stream), &li);
}
int64_t ub_op_wgrad_workspace_floats(int rows, int cols, int64_t pixels) {
    return (int64_t)wgrad_ws_floats(rows, cols, pixels);
}
int ub_op_conv3x3_wgrad(const ub_view* src0, const ub_view* src1, const void* dy, int Co,
                        float* workspace, int64_t workspace_floats, float* dw, void* stream) {
    UB_REQUIRE(src0 && dy && workspace && dw, "conv3x3_wgrad: null pointer");
    View v0 = to_view(src0), v1;
    if (src1) v1 = to_view(src1);
    return launch_wgrad(v0, src1 ? &v1 : nullptr, 0, -2, 1, 9, 3, (const __nv_bfloat16*)dy, Co, Co,
                        workspace, (size_t)workspace_floats, dw, S(stream));
}
int ub_op_convT_forward(const ub_view* x, const void* wf, const float* bias4, int Co,
                        const ub_view* dst, void* stream) {
    UB_REQUIRE(x && wf && dst, "convT_forward: null pointer");
    UB_REQUIRE(Co % 32 == 0, "convT_forward: Co must be a multiple of 32");
    IgemmEpilogue e;
    memset(&e, 0, sizeof(e));
    e.kind = EPI_CONVT; e.bias = bias4; e.ct_dst = to_view(dst);
    return launch_igemm(to_view(x), nullptr, 0, 0, 1, 1, 1, (const __nv_bfloat16*)wf, 4 * Co, e,
                        nullptr, S(stream));
}
int ub_op_convT_dgrad(const ub_view* dup, const void* wb, int Ci, void* dx, void* stream) {
    UB_REQUIRE(dup && wb && dx, "convT_dgrad: null pointer");
    IgemmEpilogue e;
    memset(&e, 0, sizeof(e));
    e.kind = EPI_STORE; e.out = (__nv_bfloat16*)dx; e.ldo = Ci;
    return launch_igemm(to_view(dup), nullptr, 0, -1, 2, 4, 2, (const __nv_bfloat16*)wb, Ci, e,
                        nullptr, S(stream));
}
int ub_op_convT_wgrad(const ub_view* dup, const void* x, int Ci, float* workspace,
                      int64_t workspace_floats, float* dw, void* stream) {
    UB_REQUIRE(dup && x && workspace && dw, "convT_wgrad: null pointer");
    return launch_wgrad(to_view(dup), nullptr, 0, -1, 2, 4, 2, (const __nv_bfloat16*)x, Ci, Ci,
                        workspace, (size_t)workspace_floats, dw, S(stream));
}

int ub_op_bn_finalize(const float* stats, const int* info, int C, const float* gamma,
                      const float* beta, float* rm, float* rv, int64_t* nbt, float momentum,
                      float eps, float* scale, float* shift, float* mean, float* rstd,
                      void* stream) {
    UB_REQUIRE(stats && info, "bn_finalize: null pointer");
    IgemmLaunchInfo li;
    li.grid = info[0]; li.n_tiles = info[1]; li.BN = info[2]; li.M = info[3];
    return launch_bn_finalize(stats, li, C, (double)li.M, gamma, beta, rm, rv, (long long*)nbt,
                              momentum, eps, scale, shift, mean, rstd, S(stream));
}
int ub_op_bn_apply_relu(const void* y, void* a, void* pooled, uint8_t* argmax, int N, int H, int W,
                        int C, const float* scale, const float* shift, void* stream) {
    return launch_bn_apply_relu((const __nv_bfloat16*)y, (__nv_bfloat16*)a, (__nv_bfloat16*)pooled,
                                argmax, N, H, W, C, scale, shift, S(stream));
}
int ub_op_bn_apply_relu_head(const void* y, void* a, int N, int H, int W, int C, const float* scale,
                             const float* shift, int n_classes, const float* head_w,
                             const float* head_b, float* logits, void* stream) {
    UB_REQUIRE(y && a && scale && shift && head_w && logits, "bn_apply_relu_head: null pointer");
    return launch_bn_apply_relu_head((const __nv_bfloat16*)y, (__nv_bfloat16*)a, N, H, W, C, scale,
                                     shift, n_classes, head_w, head_b, logits, S(stream));
}
int64_t ub_op_bn_bwd_workspace_floats(int C) { return (int64_t)bn_bwd_partial_floats(C); }
int ub_op_bn_relu_backward(const void* y, int N, int H, int W, int C, const float* scale,
                           const float* shift, const float* mean, const float* rstd,
                           const ub_view* g, const ub_view* gp, const ub_view* gs, int crop_h,
                           int crop_w, const uint8_t* argmax, float* workspace, float* dgamma,
                           float* dbeta, void* dy, void* stream) {
    UB_REQUIRE(y && workspace && dgamma && dbeta && dy, "bn_relu_backward: null pointer");
    UB_REQUIRE(g || gp, "bn_relu_backward: need g or gp");
    BnBwdDesc d;
    memset(&d, 0, sizeof(d));
    d.y = (const __nv_bfloat16*)y; d.N = N; d.H = H; d.W = W; d.C = C;
    d.scale = scale; d.shift = shift; d.mean = mean; d.rstd = rstd;
    d.pool_skip = (g == nullptr);
    if (g) d.g = to_view(g);
    if (gp) d.gp = to_view(gp);
    if (gs) { d.gs = to_view(gs); d.has_skip = true; }
    d.crop_h = crop_h; d.crop_w = crop_w;
    d.amax = argmax;
    d.partial = workspace; d.dgamma = dgamma; d.dbeta = dbeta; d.dy = (__nv_bfloat16*)dy;
    return launch_bn_bwd(d, S(stream));
}

int64_t ub_op_first_conv_workspace_floats(int Co) { return (int64_t)first_conv_partial_floats(Co); }
int ub_op_first_conv_forward(const float* x, int N, int Ci, int H, int W, const float* w,
                             const float* bias, int Co, const float* gamma, const float* beta,
                             float* rm, float* rv, int64_t* nbt, float momentum, float eps,
                             float* workspace, float* scale, float* shift, float* mean, float* rstd,
                             void* a, void* stream) {
    FirstConvDesc d;
    d.x = x; d.N = N; d.Ci = Ci; d.H = H; d.W = W; d.Co = Co; d.w = w; d.bias = bias;
    UB_TRY(launch_first_conv_train_stats(d, workspace, gamma, beta, rm, rv, (long long*)nbt,
                                         momentum, eps, scale, shift, mean, rstd, nullptr,
                                         S(stream)));
    return launch_first_conv_apply(d, scale, shift, (__nv_bfloat16*)a, S(stream));
}
int ub_op_first_conv_affine_relu(const float* x, int N, int Ci, int H, int W, const float* w, int Co,
                                 const float* scale, const float* shift, void* a, void* stream) {
    UB_REQUIRE(x && w && scale && shift && a, "first_conv_affine_relu: null pointer");
    FirstConvDesc d;
    d.x = x; d.N = N; d.Ci = Ci; d.H = H; d.W = W; d.Co = Co; d.w = w; d.bias = nullptr;
    return launch_first_conv_apply(d, scale, shift, (__nv_bfloat16*)a, S(stream));
}
int ub_op_first_conv_backward(const float* x, int N, int Ci, int H, int W, const float* w,
                              const float* bias, int Co, const float* scale, const float* shift,
                              const float* mean, const float* rstd, const ub_view* g,
                              const void* a, float* workspace, float* dgamma, float* dbeta,
                              float* dw, void* stream) {
    FirstConvDesc d;
    d.x = x; d.N = N; d.Ci = Ci; d.H = H; d.W = W; d.Co = Co; d.w = w; d.bias = bias;
    return launch_first_conv_bwd(d, scale, shift, mean, rstd, to_view(g),
                                 (const __nv_bfloat16*)a, nullptr, workspace, dgamma, dbeta, dw,
                                 S(stream));
}

int ub_op_head_forward(const void* a, int N, int H, int W, int K, int n_classes, const float* w,
                       const float* b, float* logits, uint8_t* mask, void* stream) {
    return launch_head_fwd((const __nv_bfloat16*)a, N, H, W, K, n_classes, w, b, logits, mask,
                           S(stream));
}
int64_t ub_op_head_bwd_workspace_floats(int K, int n_classes) {
    return (int64_t)head_bwd_partial_floats(K, n_classes);
}
int ub_op_head_backward(const float* dlogits, const void* a, int N, int H, int W, int K,
                        int n_classes, const float* w, void* da, float* workspace, float* dw,
                        float* db, void* stream) {
    return launch_head_bwd(dlogits, (const __nv_bfloat16*)a, N, H, W, K, n_classes, w,
                           (__nv_bfloat16*)da, workspace, dw, db, S(stream));
}
int ub_op_upsample2x_forward(const ub_view* x, void* out, void* stream) {
    UB_REQUIRE(x && out, "upsample2x_forward: null pointer");
    return launch_upsample2x_fwd(to_view(x), (__nv_bfloat16*)out, S(stream));
}
int ub_op_upsample2x_backward(const ub_view* g, void* dx, void* stream) {
    UB_REQUIRE(g && dx, "upsample2x_backward: null pointer");
    return launch_upsample2x_bwd(to_view(g), (__nv_bfloat16*)dx, S(stream));
}
int ub_op_maxpool2(const void* a, void* pooled, int N, int H, int W, int C, void* stream) {
    return launch_maxpool2((const __nv_bfloat16*)a, (__nv_bfloat16*)pooled, N, H, W, C, S(stream));
}

}  // extern "C"
