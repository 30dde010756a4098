// Host-side launchers of the HBM-bound kernels (elementwise.cuh, first_conv.cuh, head_loss.cuh,
// ccl.cuh). Grids are sized in multiples of the SM count; no launcher synchronises.
#include <cmath>

#include "ccl.cuh"
#include "elastic.cuh"
#include "elementwise.cuh"
#include "first_conv.cuh"
#include "head_loss.cuh"
#include "tiling.cuh"
#include "input_pipeline.cuh"
#include "igemm.cuh"
#include "ub_internal.h"
#include "upsample.cuh"
#include "weight_map.cuh"

namespace ub {

static inline int ew_blocks(long long items, int per_sm = 8) {
    long long b = (items + 255) / 256;
    const long long cap = (long long)num_sms() * per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}
// Grid for the "fixed channel group per thread" kernels: 4 CTAs per SM.
static inline int red_blocks(long long items, int per_sm = 4) {
    long long b = (items + 255) / 256;
    const long long cap = (long long)num_sms() * per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}
static int check_cg(int C, const char* what) {
    if (C % 8 != 0 || C <= 0 || 256 % (C / 8) != 0) {
        set_last_error("%s: channel count %d unsupported (need C/8 to divide 256)", what, C);
        return UB_ERR_UNSUPPORTED;
    }
    return UB_OK;
}

int launch_pack_conv3x3(const float* w, int Co, int Ci, __nv_bfloat16* wf, __nv_bfloat16* wd,
                        cudaStream_t s) {
    UB_LAUNCH_NC(pack_conv3x3_kernel, ew_blocks((long long)Co * Ci * 9), 256, 0, s, w, Co, Ci, wf, wd);
    UB_POST_LAUNCH();
    return UB_OK;
}
int launch_pack_convT(const float* w, int Ci, int Co, __nv_bfloat16* wf, __nv_bfloat16* wb,
                      const float* bias, float* bias4, cudaStream_t s) {
    UB_LAUNCH_NC(pack_convT2x2_kernel, ew_blocks((long long)Co * Ci * 4), 256, 0, s, w, Ci, Co, wf, wb);
    UB_POST_LAUNCH();
    if (bias && bias4) {
        UB_LAUNCH_NC(tile_bias4_kernel, (4 * Co + 255) / 256, 256, 0, s, bias, Co, bias4);
        UB_POST_LAUNCH();
    }
    return UB_OK;
}

int launch_bn_finalize(const float* stats, const IgemmLaunchInfo& info, int C, double count,
                       const float* gamma, const float* beta, float* rm, float* rv,
                       long long* nbt, float momentum, float eps, float* scale, float* shift,
                       float* mean, float* rstd, cudaStream_t s) {
    UB_LAUNCH_NC(bn_finalize_kernel, (C + 31) / 32, dim3(32, FIN_SLICES), 0, s, stats, info.grid, info.n_tiles, info.BN, C, count, gamma, beta, rm, rv, nbt, momentum, eps, scale, shift, mean, rstd);
    UB_POST_LAUNCH();
    return UB_OK;
}
int launch_bn_finalize_flat(const float* part, int blocks, int C, double count, const float* gamma,
                            const float* beta, float* rm, float* rv, long long* nbt,
                            float momentum, float eps, float* scale, float* shift, float* mean,
                            float* rstd, cudaStream_t s) {
    UB_LAUNCH_NC(bn_finalize_flat_kernel, (C + 31) / 32, dim3(32, FIN_SLICES), 0, s, part, blocks, C, count, gamma, beta, rm, rv, nbt, momentum, eps, scale, shift, mean, rstd);
    UB_POST_LAUNCH();
    return UB_OK;
}
int launch_bn_fold_eval(int C, const float* conv_bias, const float* gamma, const float* beta,
                        const float* rm, const float* rv, float eps, float* scale, float* shift,
                        cudaStream_t s) {
    UB_LAUNCH_NC(bn_fold_eval_kernel, (C + 127) / 128, 128, 0, s, C, conv_bias, gamma, beta, rm, rv, eps, scale, shift);
    UB_POST_LAUNCH();
    return UB_OK;
}

int launch_bn_apply_relu(const __nv_bfloat16* y, __nv_bfloat16* a, __nv_bfloat16* pooled,
                         unsigned char* amax, int N, int H, int W, int C, const float* scale,
                         const float* shift, cudaStream_t s) {
    UB_TRY(check_cg(C, "bn_apply"));
    if ((long long)N * H * W * (C / 8) >= 0x7FFFFFFFLL) {
        set_last_error("bn_apply: tensor too large for 32-bit indexing");
        return UB_ERR_UNSUPPORTED;
    }
    if (pooled) {
        const long long items = (long long)N * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
        UB_LAUNCH_NC((bn_apply_relu_kernel<true>), ew_blocks(items), 256, 0, s, y, a, pooled, amax, N, H, W, C, scale, shift);
    } else {
        const long long items = (long long)N * H * W * (C / 8);
        UB_LAUNCH_NC((bn_apply_relu_kernel<false>), ew_blocks(items), 256, 0, s, y, a, nullptr, nullptr, N, H, W, C, scale, shift);
    }
    UB_POST_LAUNCH();
    return UB_OK;
}
// BN-apply + ReLU of the last unit with the 1x1 head fused (training forward).
bool bn_apply_head_supported(int C, int NC) {
    return C % 8 == 0 && (C == 64 || C == 128 || C == 256) && NC >= 1 && NC <= HEAD_MAX_CLASSES;
}
int launch_bn_apply_relu_head(const __nv_bfloat16* y, __nv_bfloat16* a, int N, int H, int W, int C,
                              const float* scale, const float* shift, int NC, const float* hw,
                              const float* hb, float* logits, cudaStream_t s) {
    if (!bn_apply_head_supported(C, NC)) {
        set_last_error("bn_apply+head: C=%d / n_classes=%d unsupported", C, NC);
        return UB_ERR_UNSUPPORTED;
    }
    const long long items = (long long)N * H * W * (C / 8);
    if (items >= 0x7FFFFFFFLL) {
        set_last_error("bn_apply+head: tensor too large for 32-bit indexing");
        return UB_ERR_UNSUPPORTED;
    }
    const unsigned npix = (unsigned)((long long)N * H * W), HW = (unsigned)(H * W);
    const int blocks = ew_blocks(items);
    if (NC <= 2)
        UB_LAUNCH_NC((bn_apply_relu_head_kernel<2>), blocks, 256, 0, s, y, a, npix, HW, C, scale, shift, NC, hw, hb, logits);
    else if (NC <= 4)
        UB_LAUNCH_NC((bn_apply_relu_head_kernel<4>), blocks, 256, 0, s, y, a, npix, HW, C, scale, shift, NC, hw, hb, logits);
    else
        UB_LAUNCH_NC((bn_apply_relu_head_kernel<8>), blocks, 256, 0, s, y, a, npix, HW, C, scale, shift, NC, hw, hb, logits);
    UB_POST_LAUNCH();
    return UB_OK;
}
int launch_maxpool2(const __nv_bfloat16* a, __nv_bfloat16* pooled, int N, int H, int W, int C,
                    cudaStream_t s) {
    const long long items = (long long)N * (H / 2) * (W / 2) * (C / 8);
    UB_LAUNCH_NC(maxpool2_kernel, ew_blocks(items), 256, 0, s, a, pooled, N, H, W, C);
    UB_POST_LAUNCH();
    return UB_OK;
}

size_t bn_bwd_partial_floats(int C) { return (size_t)num_sms() * 4 * 2 * C; }

int launch_bn_bwd(const BnBwdDesc& d, cudaStream_t s, const IgemmLaunchInfo* fused) {
    UB_TRY(check_cg(d.C, "bn_bwd"));
    if ((long long)d.N * d.H * d.W * (d.C / 8) >= 0x7FFFFFFFLL) {
        set_last_error("bn_bwd: tensor too large for 32-bit indexing");
        return UB_ERR_UNSUPPORTED;
    }
    BnBwdArgs A;
    memset(&A, 0, sizeof(A));
    A.y = d.y; A.N = d.N; A.H = d.H; A.W = d.W; A.C = d.C;
    A.scale = d.scale; A.shift = d.shift; A.mean = d.mean; A.rstd = d.rstd;
    A.g = d.g; A.gp = d.gp; A.gs = d.gs; A.crop_h = d.crop_h; A.crop_w = d.crop_w;
    A.has_skip = d.has_skip ? 1 : 0;
    A.amax = d.amax;
    A.partial = d.partial;
    A.dgamma = d.dgamma; A.dbeta = d.dbeta; A.dy = d.dy;
    const long long count = (long long)d.N * d.H * d.W;
    A.inv_count = (float)(1.0 / (double)count);
    const bool pix = d.pool_skip && d.amax != nullptr;   // saved arg-max available
    const long long items = d.pool_skip
                                ? (long long)d.N * ((d.H + 1) / 2) * ((d.W + 1) / 2) * (d.C / 8)
                                : count * (d.C / 8);
    // one wave of resident CTAs (launch bounds: 3 per SM for the direct variant, 2 for the pooled one)
    const int blocks = red_blocks(items, d.pool_skip ? 2 : 3);
    if (fused) {
        // reduce pass done by the producer of the upstream gradient (EPI_STORE_BNRED epilogue)
        if (d.pool_skip) { set_last_error("bn_bwd: a fused reduction needs a direct upstream gradient"); return UB_ERR_ARG; }
        UB_LAUNCH_NC(bn_bwd_finalize_tiled_kernel, (d.C + 31) / 32, dim3(32, FIN_SLICES), 0, s, d.partial, fused->grid, fused->n_tiles, fused->BN, d.C, d.rstd, d.dgamma, d.dbeta);
        UB_POST_LAUNCH();
    } else {
        if (pix) UB_LAUNCH_NC((bn_bwd_kernel<true, false, true>), blocks, 256, 0, s, A);
        else if (d.pool_skip) UB_LAUNCH_NC((bn_bwd_kernel<true, false>), blocks, 256, 0, s, A);
        else UB_LAUNCH_NC((bn_bwd_kernel<false, false>), blocks, 256, 0, s, A);
        UB_POST_LAUNCH();
        UB_LAUNCH_NC(bn_bwd_finalize_kernel, (d.C + 31) / 32, dim3(32, FIN_SLICES), 0, s, d.partial, blocks, d.C, d.rstd, d.dgamma, d.dbeta);
        UB_POST_LAUNCH();
    }
    // UB_BNBWD_APPLY_CTAS=n: CTAs per SM of the apply pass (default 8 = several waves)
    static const int apply_ctas = [] { const char* e = getenv("UB_BNBWD_APPLY_CTAS"); return e ? atoi(e) : 8; }();
    const int ablocks = ew_blocks(items, apply_ctas);
    if (pix) UB_LAUNCH_NC((bn_bwd_kernel<true, true, true>), ablocks, 256, 0, s, A);
    else if (d.pool_skip) UB_LAUNCH_NC((bn_bwd_kernel<true, true>), ablocks, 256, 0, s, A);
    else UB_LAUNCH_NC((bn_bwd_kernel<false, true>), ablocks, 256, 0, s, A);
    UB_POST_LAUNCH();
    return UB_OK;
}

// ---------------------------------------------------------------------------------------------
size_t first_conv_partial_floats(int Co) {
    const size_t a = (size_t)num_sms() * 4 * 2 * Co;
    const size_t b = (size_t)num_sms() * 4 * Co * 10;
    const size_t c = (size_t)num_sms() * 2 * FC_COV_TERMS * 2;
    size_t m = a > b ? a : b;
    if (c > m) m = c;
    return m + 128;   // + room for the re-computed patch moments (op-level backward)
}
static int fc_fill(const FirstConvDesc& d, FirstConvArgs& A) {
    UB_TRY(check_cg(d.Co, "first conv"));
    if (d.H < 3 || d.W < 3) { set_last_error("first conv: input smaller than 3x3"); return UB_ERR_ARG; }
    if ((size_t)d.Co * d.Ci * 9 * 4 > 96 * 1024) {
        set_last_error("first conv: n_channels=%d too large for the fp32 direct kernel", d.Ci);
        return UB_ERR_UNSUPPORTED;
    }
    memset(&A, 0, sizeof(A));
    A.x = d.x; A.N = d.N; A.Ci = d.Ci; A.H = d.H; A.W = d.W; A.Co = d.Co; A.w = d.w; A.bias = d.bias;
    return UB_OK;
}
// generic n_channels > 1 path: direct fp32 kernel, weights in shared memory
template <int MODE>
static int fc_launch(const FirstConvArgs& A, int blocks, cudaStream_t s) {
    const size_t smem = (size_t)A.Co * A.Ci * 9 * 4;
    if (smem > 48 * 1024)
        UB_CHECK_CUDA(cudaFuncSetAttribute(first_conv_kernel<MODE, false>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    UB_LAUNCH_NC((first_conv_kernel<MODE, false>), blocks, 256, smem, s, A);
    UB_POST_LAUNCH();
    return UB_OK;
}
constexpr int FC1_PX = 4;
// single-channel patch moments -> cov[] (+ BatchNorm statistics when gamma != null)
static int fc1_moments(const FirstConvDesc& d, float* ws, double* cov, const float* gamma,
                       const float* beta, float* rm, float* rv, long long* nbt, float momentum,
                       float eps, float* scale, float* shift, float* mean, float* rstd,
                       cudaStream_t s) {
    const long long groups = (long long)d.N * (d.H - 2) * ((d.W - 2 + FC1_PX - 1) / FC1_PX);
    long long b = (groups + 255) / 256;
    if (b > num_sms() * 2) b = num_sms() * 2;
    double* partial = reinterpret_cast<double*>(ws);
    UB_LAUNCH_NC((fc1_cov_kernel<FC1_PX>), (int)b, 256, 0, s, d.x, d.N, d.H, d.W, partial);
    UB_POST_LAUNCH();
    const double count = (double)d.N * (d.H - 2) * (d.W - 2);
    UB_LAUNCH_NC(fc1_cov_finalize_kernel, 1, 1024, 0, s, partial, (int)b, d.x, count, cov, d.Co, d.w, d.bias, gamma, beta, rm, rv, nbt, momentum, eps, scale, shift, mean, rstd);
    UB_POST_LAUNCH();
    return UB_OK;
}
int launch_first_conv_train_stats(const FirstConvDesc& d, float* ws, const float* gamma,
                                  const float* beta, float* rm, float* rv, long long* nbt,
                                  float momentum, float eps, float* scale, float* shift,
                                  float* mean, float* rstd, double* cov, cudaStream_t s) {
    FirstConvArgs A;
    UB_TRY(fc_fill(d, A));
    if ((reinterpret_cast<uintptr_t>(ws) & 7) != 0) {
        set_last_error("first conv: workspace must be 8-byte aligned");
        return UB_ERR_ARG;
    }
    if (d.Ci == 1)
        return fc1_moments(d, ws, cov, gamma, beta, rm, rv, nbt, momentum, eps, scale, shift, mean,
                           rstd, s);
    A.partial = ws;
    const long long items = (long long)d.N * (d.H - 2) * (d.W - 2) * (d.Co / 8);
    const int blocks = red_blocks(items);
    UB_TRY(fc_launch<FC_STATS>(A, blocks, s));
    return launch_bn_finalize_flat(ws, blocks, d.Co, (double)d.N * (d.H - 2) * (d.W - 2), gamma,
                                   beta, rm, rv, nbt, momentum, eps, scale, shift, mean, rstd, s);
}
int launch_first_conv_apply(const FirstConvDesc& d, const float* scale, const float* shift,
                            __nv_bfloat16* a, cudaStream_t s) {
    FirstConvArgs A;
    UB_TRY(fc_fill(d, A));
    A.scale = scale; A.shift = shift; A.a = a;
    if (d.Ci == 1) {
        const long long groups = (long long)d.N * (d.H - 2) * ((d.W - 2 + FC1_PX - 1) / FC1_PX);
        UB_LAUNCH_NC((fc1_apply_kernel<FC1_PX>), ew_blocks(groups * (d.Co / 8)), 256, 0, s, A);
        UB_POST_LAUNCH();
        return UB_OK;
    }
    const long long items = (long long)d.N * (d.H - 2) * (d.W - 2) * (d.Co / 8);
    return fc_launch<FC_APPLY>(A, ew_blocks(items), s);
}
int launch_first_conv_bwd(const FirstConvDesc& d, const float* scale, const float* shift,
                          const float* mean, const float* rstd, const View& g,
                          const __nv_bfloat16* a, const double* cov, float* ws, float* dgamma,
                          float* dbeta, float* dw, cudaStream_t s) {
    FirstConvArgs A;
    UB_TRY(fc_fill(d, A));
    A.scale = scale; A.shift = shift; A.mean = mean; A.rstd = rstd; A.g = g;
    const long long count = (long long)d.N * (d.H - 2) * (d.W - 2);
    const long long items = count * (d.Co / 8);
    if (d.Ci == 1) {
        if (!a) { set_last_error("first conv backward: the forward activation is required"); return UB_ERR_ARG; }
        if ((reinterpret_cast<uintptr_t>(ws) & 7) != 0) {
            set_last_error("first conv: workspace must be 8-byte aligned");
            return UB_ERR_ARG;
        }
        float* partial = ws + 128;
        if (!cov) {   // op-level call without plan state: recompute the patch moments
            double* c = reinterpret_cast<double*>(ws);
            UB_TRY(fc1_moments(d, partial, c, nullptr, nullptr, nullptr, nullptr, nullptr, 0.f, 0.f,
                               nullptr, nullptr, nullptr, nullptr, s));
            cov = c;
        }
        A.a = const_cast<__nv_bfloat16*>(a);
        A.partial = partial;
        constexpr int PX = 2;
        const long long groups = (long long)d.N * (d.H - 2) * ((d.W - 2 + PX - 1) / PX);
        long long b = (groups * (d.Co / 8) + 255) / 256;
        if (b > num_sms() * 2) b = num_sms() * 2;
        UB_LAUNCH_NC((fc1_bwd_kernel<PX>), (int)b, 256, 0, s, A, cov);
        UB_POST_LAUNCH();
        UB_LAUNCH_NC(fc1_bwd_finalize_kernel, d.Co, 320, 0, s, partial, (int)b, d.Co, cov, d.w, d.bias, scale, mean, rstd, dgamma, dbeta, dw);
        UB_POST_LAUNCH();
        return UB_OK;
    }
    A.partial = ws;
    int blocks = red_blocks(items);
    UB_TRY(fc_launch<FC_BWD_REDUCE>(A, blocks, s));
    UB_LAUNCH_NC(bn_bwd_finalize_kernel, (d.Co + 31) / 32, dim3(32, FIN_SLICES), 0, s, ws, blocks, d.Co, (const float*)nullptr, dgamma, dbeta);
    UB_POST_LAUNCH();
    A.dgamma = dgamma; A.dbeta = dbeta; A.inv_count = (float)(1.0 / (double)count);
    A.wpartial = ws;
    for (int ci = 0; ci < d.Ci; ++ci) {
        A.ci_sel = ci;
        const int wblocks = red_blocks(items);
        UB_TRY(fc_launch<FC_BWD_WGRAD>(A, wblocks, s));
        UB_LAUNCH_NC(first_wgrad_finalize_kernel, (d.Co * 9 + 127) / 128, 128, 0, s, ws, wblocks, d.Co, d.Ci, ci, dw);
        UB_POST_LAUNCH();
    }
    return UB_OK;
}

// ---------------------------------------------------------------------------------------------
int launch_head_fwd(const __nv_bfloat16* a, int N, int H, int W, int K, int NC, const float* w,
                    const float* b, float* logits, unsigned char* mask, cudaStream_t s) {
    if (NC < 1 || NC > HEAD_MAX_CLASSES || K % 8) {
        set_last_error("head: n_classes=%d (max %d) / K=%d unsupported", NC, HEAD_MAX_CLASSES, K);
        return UB_ERR_UNSUPPORTED;
    }
    const long long P = (long long)N * H * W;
    if (P >= 0x7FFFFFF0LL / 8) { set_last_error("head: too many pixels"); return UB_ERR_UNSUPPORTED; }
    if (K == 64 || K == 128 || K == 256) {
        const int blocks = ew_blocks(P * (K / 8));
        const unsigned npix = (unsigned)P, HW = (unsigned)(H * W);
        if (NC <= 2)
            UB_LAUNCH_NC((head_fwd_vec_kernel<2>), blocks, 256, 0, s, a, npix, HW, K, NC, w, b, logits, mask);
        else if (NC <= 4)
            UB_LAUNCH_NC((head_fwd_vec_kernel<4>), blocks, 256, 0, s, a, npix, HW, K, NC, w, b, logits, mask);
        else
            UB_LAUNCH_NC((head_fwd_vec_kernel<8>), blocks, 256, 0, s, a, npix, HW, K, NC, w, b, logits, mask);
        UB_POST_LAUNCH();
        return UB_OK;
    }
    UB_LAUNCH_NC(head_fwd_kernel, ew_blocks(P), 256, (size_t)(NC * K + NC) * 4, s, a, P, (long long)H * W, K, NC, w, b, logits, mask);
    UB_POST_LAUNCH();
    return UB_OK;
}
size_t head_bwd_partial_floats(int K, int NC) { return (size_t)num_sms() * 4 * (NC * K + NC); }
int launch_head_bwd(const float* dlogits, const __nv_bfloat16* a, int N, int H, int W, int K,
                    int NC, const float* w, __nv_bfloat16* da, float* partial, float* dw, float* db,
                    cudaStream_t s) {
    if (NC < 1 || NC > HEAD_MAX_CLASSES) {
        set_last_error("head: n_classes=%d unsupported", NC);
        return UB_ERR_UNSUPPORTED;
    }
    UB_TRY(check_cg(K, "head_bwd"));
    const long long P = (long long)N * H * W;
    const int blocks = red_blocks(P * (K / 8));
    if (NC <= 2)
        UB_LAUNCH_NC((head_bwd_kernel<2>), blocks, 256, 0, s, dlogits, a, P, (long long)H * W, K, NC, w, da, partial);
    else if (NC <= 4)
        UB_LAUNCH_NC((head_bwd_kernel<4>), blocks, 256, 0, s, dlogits, a, P, (long long)H * W, K, NC, w, da, partial);
    else
        UB_LAUNCH_NC((head_bwd_kernel<8>), blocks, 256, 0, s, dlogits, a, P, (long long)H * W, K, NC, w, da, partial);
    UB_POST_LAUNCH();
    const int len = NC * K + NC;
    UB_LAUNCH_NC(reduce_partials_kernel, (len + 31) / 32, dim3(32, FIN_SLICES), 0, s, partial, blocks, len, dw, NC * K, db);
    UB_POST_LAUNCH();
    return UB_OK;
}

size_t wce_partial_floats() { return (size_t)num_sms() * 8; }
int launch_wce(const WceDesc& d, float* loss, float* dz, float* partial, int* err,
               cudaStream_t s) {
    WceArgs A;
    A.z = d.z; A.zN = d.zs[0]; A.zC = d.zs[1]; A.zH = d.zs[2]; A.zW = d.zs[3];
    A.t = d.t; A.tN = d.ts[0]; A.tH = d.ts[1]; A.tW = d.ts[2];
    A.wm = d.w; A.wN = d.ws[0]; A.wH = d.ws[1]; A.wW = d.ws[2];
    A.N = d.N; A.C = d.C; A.H = d.H; A.W = d.W;
    A.dz = dz; A.partial = partial; A.err = err;
    const long long P = (long long)d.N * d.H * d.W;
    if (P <= 0 || d.C < 1) { set_last_error("wce: empty input"); return UB_ERR_ARG; }
    const int blocks = ew_blocks(P);
    UB_LAUNCH_NC(wce_fwd_bwd_kernel, blocks, 256, 0, s, A);
    UB_POST_LAUNCH();
    UB_LAUNCH_NC(wce_finalize_kernel, 1, 256, 0, s, partial, blocks, (double)P, loss);
    UB_POST_LAUNCH();
    return UB_OK;
}
int launch_scale_by_scalar(const float* in, const float* scalar, float* out, long long n,
                           cudaStream_t s) {
    UB_LAUNCH_NC(scale_by_scalar_kernel, ew_blocks(n), 256, 0, s, in, scalar, out, n);
    UB_POST_LAUNCH();
    return UB_OK;
}
int launch_fill_zero(float* p, long long n, cudaStream_t s) {
    if (n <= 0) return UB_OK;
    UB_LAUNCH_NC(fill_zero_kernel, ew_blocks(n), 256, 0, s, p, n);
    UB_POST_LAUNCH();
    return UB_OK;
}

// ---------------------------------------------------------------------------------------------
int launch_prepare_batch(const unsigned char* img, const void* labels, int label_bytes,
                         const void* wmap, int wmap_bytes, int N, int H, int W, int oh, int ow,
                         float* image, long long* target, float* weight, cudaStream_t s) {
    if (!img || !image || N < 1 || H < 1 || W < 4 || W % 4 != 0) {
        set_last_error("prepare_batch: need uint8 images with a width that is a multiple of 4");
        return UB_ERR_ARG;
    }
    if (oh < 0 || ow < 0 || oh > H || ow > W || (labels && !target) || (wmap && !weight) ||
        (labels && label_bytes != 1 && label_bytes != 2) || (wmap && wmap_bytes != 4 && wmap_bytes != 8)) {
        set_last_error("prepare_batch: bad crop size or label / weight-map element size");
        return UB_ERR_ARG;
    }
    if ((long long)N * H * W >= 0x7FFFFFFFLL) {
        set_last_error("prepare_batch: batch too large for 32-bit indexing");
        return UB_ERR_UNSUPPORTED;
    }
    PrepArgs A;
    A.img = img; A.labels = labels; A.wmap = wmap; A.label_bytes = label_bytes; A.wmap_bytes = wmap_bytes;
    A.N = N; A.H = H; A.W = W; A.oh = oh; A.ow = ow;
    A.h0 = (H - oh) / 2; A.w0 = (W - ow) / 2;   // center_crop_tensor, scripts/train.py:39-51
    A.image = image; A.target = target; A.weight = weight;
    UB_LAUNCH_NC(prepare_batch_kernel, ew_blocks((long long)N * H * (W / 4)), 256, 0, s, A);
    UB_POST_LAUNCH();
    return UB_OK;
}

// ---------------------------------------------------------------------------------------------
static int check_up_view(const View& v, const char* what) {
    if (!v.ptr || v.N < 1 || v.H < 1 || v.W < 1 || v.C < 8 || v.C % 8 != 0 || v.sW % 8 != 0 ||
        v.sH % 8 != 0 || v.sN % 8 != 0 || ((uintptr_t)v.ptr & 15) != 0) {
        set_last_error("%s: need an NHWC bf16 view with C %% 8 == 0 and 16-byte aligned pixels", what);
        return UB_ERR_ARG;
    }
    return UB_OK;
}
int launch_upsample2x_fwd(const View& x, __nv_bfloat16* out, cudaStream_t s) {
    UB_TRY(check_up_view(x, "upsample2x forward"));
    if (!out) { set_last_error("upsample2x forward: null output"); return UB_ERR_ARG; }
    const long long items = (long long)x.N * 4 * x.H * x.W * (x.C / 8);
    UB_LAUNCH_NC(upsample2x_fwd_kernel, ew_blocks(items), 256, 0, s, x, out);
    UB_POST_LAUNCH();
    return UB_OK;
}
int launch_upsample2x_bwd(const View& g, __nv_bfloat16* dx, cudaStream_t s) {
    UB_TRY(check_up_view(g, "upsample2x backward"));
    if (!dx || g.H % 2 != 0 || g.W % 2 != 0) {
        set_last_error("upsample2x backward: null output or odd gradient size %dx%d", g.H, g.W);
        return UB_ERR_ARG;
    }
    const long long items = (long long)g.N * (g.H / 2) * (g.W / 2) * (g.C / 8);
    UB_LAUNCH_NC(upsample2x_bwd_kernel, ew_blocks(items), 256, 0, s, g, dx, g.H / 2, g.W / 2);
    UB_POST_LAUNCH();
    return UB_OK;
}

// ---------------------------------------------------------------------------------------------
template <typename LabelT, int VEC>
static int weight_map_typed(const LabelT* labels, int N, unsigned P, double border, void* out,
                            int out_bytes, unsigned* counts, cudaStream_t s) {
    // one image per grid row; enough blocks per image to fill the machine with the whole batch
    int per_image = ew_blocks((long long)(P / VEC));
    const int want = (num_sms() * 8 + N - 1) / N;
    if (per_image > want) per_image = want;
    const dim3 grid((unsigned)per_image, (unsigned)N);
    UB_LAUNCH_NC((wmap_count_kernel<LabelT, VEC>), grid, 256, 0, s, labels, P, counts);
    UB_POST_LAUNCH();
    if (out_bytes == 8)
        UB_LAUNCH_NC((wmap_emit_kernel<LabelT, double, VEC>), grid, 256, 0, s, labels, P, counts, border, (double*)out);
    else
        UB_LAUNCH_NC((wmap_emit_kernel<LabelT, float, VEC>), grid, 256, 0, s, labels, P, counts, border, (float*)out);
    UB_POST_LAUNCH();
    return UB_OK;
}
int launch_weight_map(const void* labels, int label_bytes, int N, int H, int W, double w0,
                      double sigma, void* out, int out_bytes, unsigned* counts, cudaStream_t s) {
    if (!labels || !out || !counts || N < 1 || H < 1 || W < 1 || N > 65535) {
        set_last_error("weight_map: null pointer or bad shape (N=%d H=%d W=%d)", N, H, W);
        return UB_ERR_ARG;
    }
    if ((label_bytes != 1 && label_bytes != 2) || (out_bytes != 4 && out_bytes != 8)) {
        set_last_error("weight_map: labels must be uint8 / uint16, output float32 / float64");
        return UB_ERR_ARG;
    }
    if (!(w0 == w0) || !(sigma == sigma) || !(sigma * sigma + 1e-8 > 0.0)) {
        set_last_error("weight_map: w0 / sigma must be finite numbers");
        return UB_ERR_ARG;
    }
    const long long P = (long long)H * W;
    if (P >= 0x7FFFFFFFLL) {
        set_last_error("weight_map: image too large for 32-bit indexing");
        return UB_ERR_UNSUPPORTED;
    }
    // w0 * exp(-((d1 + d2)^2) / (2 (sigma^2 + 1e-8))) with d1 = d2 = 0 (see weight_map.cuh)
    const double border = w0 * std::exp(-0.0 / (2.0 * (sigma * sigma + 1e-8)));
    UB_CHECK_CUDA(cudaMemsetAsync(counts, 0, (size_t)N * sizeof(unsigned), s));
    const bool vec = P % 4 == 0;
    if (label_bytes == 1)
        return vec ? weight_map_typed<unsigned char, 4>((const unsigned char*)labels, N, (unsigned)P, border, out, out_bytes, counts, s)
                   : weight_map_typed<unsigned char, 1>((const unsigned char*)labels, N, (unsigned)P, border, out, out_bytes, counts, s);
    return vec ? weight_map_typed<unsigned short, 4>((const unsigned short*)labels, N, (unsigned)P, border, out, out_bytes, counts, s)
               : weight_map_typed<unsigned short, 1>((const unsigned short*)labels, N, (unsigned)P, border, out, out_bytes, counts, s);
}

// ---------------------------------------------------------------------------------------------
// Elastic deformation: noise [2][N][H][W] (dx draws, then dy draws) -> workspace halves
// [0] = after the row pass, [1] = displacement fields [2][N][H][W]; then one sampling pass.
size_t elastic_ws_bytes(int N, int H, int W) { return (size_t)4 * N * H * W * sizeof(double); }
template <typename LI, typename LO>
static int elastic_sample_typed(const unsigned char* img, const void* labels, const double* dx,
                                const double* dy, int N, int H, int W, unsigned char* img_out,
                                void* labels_out, cudaStream_t s) {
    UB_LAUNCH_NC((elastic_sample_kernel<LI, LO>), ew_blocks((long long)N * H * W), 256, 0, s, img,
                 (const LI*)labels, dx, dy, N, H, W, img_out, (LO*)labels_out);
    UB_POST_LAUNCH();
    return UB_OK;
}
int launch_elastic(const unsigned char* img, const void* labels, int label_bytes, int N, int H, int W,
                   const double* noise, const double* taps, int radius, double alpha,
                   unsigned char* img_out, void* labels_out, int label_out_bytes, void* ws,
                   cudaStream_t s) {
    if (!noise || !taps || !ws || N < 1 || H < 1 || W < 1 || radius < 0) {
        set_last_error("elastic_deform: null pointer or bad shape (N=%d H=%d W=%d radius=%d)", N, H, W, radius);
        return UB_ERR_ARG;
    }
    if ((img && !img_out) || (labels && !labels_out) || (!img && !labels) ||
        (labels && ((label_bytes != 1 && label_bytes != 2) || (label_out_bytes != 1 && label_out_bytes != 2)))) {
        set_last_error("elastic_deform: need images and / or labels (uint8 / uint16) with their outputs");
        return UB_ERR_ARG;
    }
    if ((long long)N * H * W >= 0x7FFFFFFFLL || 2LL * N > 65535) {
        set_last_error("elastic_deform: batch too large for 32-bit indexing");
        return UB_ERR_UNSUPPORTED;
    }
    const size_t smem_v = elastic_v_smem(radius), smem_h = elastic_h_smem(radius);
    if (smem_v > 200 * 1024 || smem_h > 200 * 1024) {
        set_last_error("elastic_deform: Gaussian radius %d (sigma too large) exceeds the shared-memory tile", radius);
        return UB_ERR_UNSUPPORTED;
    }
    // opt-in to > 48 KB of dynamic shared memory (per device, so not cached in a static)
    UB_CHECK_CUDA(cudaFuncSetAttribute(elastic_blur_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    UB_CHECK_CUDA(cudaFuncSetAttribute(elastic_blur_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const size_t field = (size_t)N * H * W;
    double* tmp = reinterpret_cast<double*>(ws);
    double* disp = tmp + 2 * field;
    const dim3 gv((W + 31) / 32, (H + EL_V_ROWS - 1) / EL_V_ROWS, 2 * N);
    const dim3 gh((W + EL_H_COLS - 1) / EL_H_COLS, (H + EL_H_ROWS - 1) / EL_H_ROWS, 2 * N);
    if (gv.y > 65535 || gh.y > 65535) {
        set_last_error("elastic_deform: image too tall");
        return UB_ERR_UNSUPPORTED;
    }
    UB_LAUNCH_NC(elastic_blur_rows_kernel, gv, dim3(32, 8), smem_v, s, noise, tmp, H, W, taps, radius, 1);
    UB_POST_LAUNCH();
    UB_LAUNCH_NC(elastic_blur_cols_kernel, gh, dim3(32, 8), smem_h, s, (const double*)tmp, disp, H, W, taps, radius, alpha);
    UB_POST_LAUNCH();
    const double* dx = disp;
    const double* dy = disp + field;
    if (!labels) return elastic_sample_typed<unsigned char, unsigned char>(img, nullptr, dx, dy, N, H, W, img_out, nullptr, s);
    if (label_bytes == 1)
        return label_out_bytes == 1
                   ? elastic_sample_typed<unsigned char, unsigned char>(img, labels, dx, dy, N, H, W, img_out, labels_out, s)
                   : elastic_sample_typed<unsigned char, unsigned short>(img, labels, dx, dy, N, H, W, img_out, labels_out, s);
    return label_out_bytes == 1
               ? elastic_sample_typed<unsigned short, unsigned char>(img, labels, dx, dy, N, H, W, img_out, labels_out, s)
               : elastic_sample_typed<unsigned short, unsigned short>(img, labels, dx, dy, N, H, W, img_out, labels_out, s);
}

// ---------------------------------------------------------------------------------------------
size_t ccl_ws_bytes(int H, int W) {
    const size_t n = (size_t)H * W;
    const size_t nb = (n + 1023) / 1024;
    const size_t nc = (nb + 1023) / 1024;
    return (3 * n + nb + nc + 16) * sizeof(int);
}
int launch_ccl(const unsigned char* mask, int H, int W, int min_size, unsigned short* out, void* ws,
               cudaStream_t s) {
    const long long n = (long long)H * W;
    if (n <= 0) return UB_OK;
    if (n > 0x7FFFFFFFLL) { set_last_error("ccl: image too large"); return UB_ERR_UNSUPPORTED; }
    const int nb = (int)((n + 1023) / 1024);
    const int nc = (nb + 1023) / 1024;
    int* L = reinterpret_cast<int*>(ws);
    int* area = L + n;
    int* rank = area + n;
    int* block_roots = rank + n;
    int* chunk_tot = block_roots + nb;
    ccl_init_kernel<<<ew_blocks(n), 256, 0, s>>>(mask, L, area, n, W);
    ccl_merge_kernel<<<ew_blocks(n), 256, 0, s>>>(mask, L, H, W);
    ccl_flatten_kernel<<<nb, 1024, 0, s>>>(L, area, block_roots, n);
    ccl_scan_chunks_kernel<<<nc, 1024, 0, s>>>(block_roots, nb, chunk_tot);
    int extra = 0;
    if (nc > 1) {     // more than 1024 blocks of 1024 pixels (images beyond 1 Mpix)
        ccl_scan_kernel<<<1, 1024, 0, s>>>(chunk_tot, nc);
        ccl_scan_add_kernel<<<nc, 1024, 0, s>>>(block_roots, nb, chunk_tot);
        extra = 2;
    }
    ccl_rank_kernel<<<nb, 1024, 0, s>>>(L, block_roots, rank, n);
    ccl_emit_kernel<<<ew_blocks(n), 256, 0, s>>>(L, area, rank, min_size, out, n);
    count_launch(5 + extra);
    UB_POST_LAUNCH();
    return UB_OK;
}

// ---------------------------------------------------------------------------------------------
int launch_extract_tiles(const float* image, int H, int W, const int* origins_yx, int T, int S,
                         int margin, float* tiles, cudaStream_t s) {
    if (!image || !origins_yx || !tiles || H < 1 || W < 1 || T < 1 || S < 4 || S % 4 != 0 || margin < 0) {
        set_last_error("extract_tiles: null pointer or bad shape (H=%d W=%d T=%d tile=%d)", H, W, T, S);
        return UB_ERR_ARG;
    }
    UB_LAUNCH_NC(extract_tiles_kernel, ew_blocks((long long)T * S * (S / 4)), 256, 0, s, image, H, W,
                 reinterpret_cast<const int2*>(origins_yx), T, S, margin, tiles);
    UB_POST_LAUNCH();
    return UB_OK;
}
int launch_stitch_tiles(const unsigned char* tiles, const int* origins_yx, int T, int TO,
                        unsigned char* full, int H, int W, cudaStream_t s) {
    if (!tiles || !origins_yx || !full || H < 1 || W < 1 || T < 1 || TO < 4 || TO % 4 != 0) {
        set_last_error("stitch_tiles: null pointer or bad shape (H=%d W=%d T=%d tile_out=%d)", H, W, T, TO);
        return UB_ERR_ARG;
    }
    UB_LAUNCH_NC(stitch_tiles_kernel, ew_blocks((long long)T * TO * (TO / 4)), 256, 0, s, tiles,
                 reinterpret_cast<const int2*>(origins_yx), T, TO, full, H, W);
    UB_POST_LAUNCH();
    return UB_OK;
}

}  // namespace ub
