// Network executor: the whole UNet.forward (reference models/unet_model.py:105-146) and its
// backward as a fixed sequence of kernel launches over a pre-allocated NHWC bf16 activation arena.
// No allocation, no host synchronisation and no tensor-map-independent host work happens after
// ub_plan_create, so a step can be captured into a CUDA graph.
#include <stdlib.h>

#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/unet_b200.h"
#include "igemm.cuh"
#include "sgd.cuh"
#include "ub_internal.h"

using namespace ub;
typedef __nv_bfloat16 bf16;

namespace {

struct ConvUnit {
    int Ci = 0, Co = 0, Hin = 0, Win = 0;  // output is (Hin-2) x (Win-2)
    int p_w = 0, p_b = 0, p_g = 0, p_be = 0, bn = 0;
    bool first = false;
    bf16 *wf = nullptr, *wd = nullptr;
    bf16 *y = nullptr, *a = nullptr, *dy = nullptr;
    float *scale = nullptr, *shift = nullptr, *mean = nullptr, *rstd = nullptr;
    IgemmLaunchInfo info{};
    int Ho() const { return Hin - 2; }
    int Wo() const { return Win - 2; }
};
struct Block {
    ConvUnit u[2];
    View in0{}, in1{};
    bool two = false;
    bool pool = false;
    bf16* pooled = nullptr;  // [N, Ho/2, Wo/2, Co]
    unsigned char* amax = nullptr;  // pool arg-max per pooled element (training)
    bf16* da0 = nullptr;     // gradient of u[0].a
    bf16* din = nullptr;     // gradient of the block input [N, Hin, Win, Ci] (contiguous)
};
struct UpT {
    int Ci = 0, Co = 0, Hin = 0, Win = 0;
    int p_w = 0, p_b = 0;
    bf16 *wf = nullptr, *wb = nullptr;
    float* bias4 = nullptr;
    bf16* out = nullptr;  // [N, 2Hin, 2Win, Co]
    bf16* dx = nullptr;   // [N, Hin, Win, Ci]
};

}  // namespace

struct ub_plan {
    int N = 0, Cin = 0, H = 0, W = 0, base = 0, L = 0, NC = 0;
    bool training = false;
    bool bilinear = false;   // Up blocks use nn.Upsample(bilinear, align_corners) instead of ConvTranspose2d
    int outH = 0, outW = 0;
    std::vector<Block> enc, dec;
    std::vector<UpT> ups;
    std::vector<int> crop;  // crop start of the skip of dec[j]
    std::vector<const float*> params;
    std::vector<long long> param_numel;
    std::vector<float*> rm, rv;
    std::vector<long long*> nbt;
    std::vector<void*> allocs;
    size_t bytes = 0;
    float* scratch = nullptr;   // stats / reduction partials
    // logits written by the last unit itself: the last BN-apply kernel (training) or the last conv's
    // epilogue (eval, 64 channels) — the separate head kernel is skipped
    float* fused_logits = nullptr;
    unsigned char* fused_mask = nullptr;
    double* fc_cov = nullptr;   // patch moments of the single-channel first conv (forward -> backward)
    float* wgrad_ws = nullptr;
    size_t wgrad_ws_floats = 0;
    bf16* head_da = nullptr;
    const float* x = nullptr;   // input of the last forward (kept by the caller)
    bool packed = false;
    // per BatchNorm layer (ub_plan_set_bn_config): nn.BatchNorm2d defaults unless the module says otherwise
    std::vector<float> bn_mom, bn_eps;
    // Weight gradients run on a low-priority side stream: they are off the backward critical path
    // (BN backward -> data gradient -> BN backward ...), tensor-bound, and small enough in registers
    // (256 threads x 74) to share an SM with the HBM-bound BN-backward CTAs, so the two overlap.
    //   overlap: 0 = everything on the caller's stream, 1 = join at the end of every backward stage
    //   (a stage's gradients are complete when the call returns its work to the stream: what the
    //   data-parallel all-reduce hook needs), 2 = join only at the end of the last stage.
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int overlap = 1;
    bool side_used = false;
    // Eval plans: the launch sequence of a forward pass (~60 kernels, each with freshly encoded tensor
    // maps) is captured into a CUDA graph the second time ub_plan_forward sees the same (x, logits,
    // mask) pointers and parameter binding, and replayed afterwards: per-frame prediction
    // (scripts/predict.py:73-112) is otherwise bound by host launch work, not by the GPU.
    // BN-backward reduce pass fused into the epilogue of the data-gradient kernel that produces the
    // upstream gradient (EPI_STORE_BNRED): layout of the partial rows left in `scratch` by the transposed
    // conv's data gradient at the end of a decoder stage, consumed by the first BN backward of the next
    bool dx_fused = false;
    IgemmLaunchInfo dx_info{};
    cudaGraphExec_t fwd_graph = nullptr;
    cudaStream_t cap_stream = nullptr;   // capture happens here (the caller's may be the legacy stream)
    const float* g_x = nullptr;
    float* g_logits = nullptr;
    uint8_t* g_mask = nullptr;
    unsigned long long bind_epoch = 0, g_epoch = 0;
    int g_calls = 0;
    bool g_disabled = false;             // a capture attempt failed: stay on plain launches
    long long g_nodes = 0;
    long long g_replays = 0;
    // optional in-step kernel timing (CUDA events on the launching stream)
    struct ProfRec { int cls; double flops, bytes; cudaEvent_t e0, e1; };
    bool prof_on = false;
    std::vector<ProfRec> prof;

    template <typename T>
    int alloc(T** p, size_t count) {
        void* q = nullptr;
        const size_t b = ((count * sizeof(T) + 255) / 256) * 256;
        cudaError_t e = cudaMalloc(&q, b ? b : 256);
        if (e != cudaSuccess) {
            set_last_error("cudaMalloc of %zu bytes failed: %s", b, cudaGetErrorString(e));
            cudaGetLastError();
            return ub::UB_ERR_NOMEM;
        }
        allocs.push_back(q);
        bytes += b;
        *p = reinterpret_cast<T*>(q);
        return 0;
    }
    void drop_graph() {
        if (fwd_graph) cudaGraphExecDestroy(fwd_graph);
        fwd_graph = nullptr;
        g_calls = 0;
    }
    ~ub_plan() {
        drop_graph();
        if (cap_stream) cudaStreamDestroy(cap_stream);
        if (side) cudaStreamDestroy(side);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        for (void* q : allocs) cudaFree(q);
        for (auto& r : prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    }
};

enum : int { CLS_FPROP = 0, CLS_DGRAD, CLS_WGRAD, CLS_CT_FPROP, CLS_CT_DGRAD, CLS_CT_WGRAD,
             CLS_BN_APPLY, CLS_BN_BWD, CLS_FIRST, CLS_HEAD, CLS_SGD, CLS_COUNT };

static const char* const kClassNames[CLS_COUNT] = {
    "conv3x3_fprop", "conv3x3_dgrad", "conv3x3_wgrad", "convT_fprop", "convT_dgrad", "convT_wgrad",
    "bn_apply_relu_pool", "bn_relu_backward", "first_conv_fp32", "head_1x1", "sgd_update"};
static bool nvtx_on() {   // UB_NVTX=1: name every kernel class for `ncu --nvtx --print-nvtx-rename kernel`
    static const bool on = [] { const char* e = getenv("UB_NVTX"); return e && e[0] == '1'; }();
    return on;
}

// Brackets the launches issued inside its scope with two CUDA events when profiling is enabled
// (and with an NVTX range named after the kernel class under UB_NVTX=1).
struct ProfScope {
    ub_plan* P; cudaStream_t s; cudaEvent_t e1 = nullptr; bool range = false;
    ProfScope(ub_plan* P_, int cls, double flops, double bytes, cudaStream_t s_) : P(P_), s(s_) {
        if (nvtx_on()) { nvtxRangePushA(kClassNames[cls]); range = true; }
        if (!P->prof_on) return;
        ub_plan::ProfRec r;
        r.cls = cls; r.flops = flops; r.bytes = bytes;
        if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
        cudaEventRecord(r.e0, s);
        e1 = r.e1;
        P->prof.push_back(r);
    }
    ~ProfScope() {
        if (e1) cudaEventRecord(e1, s);
        if (range) nvtxRangePop();
    }
};
static inline double gemm_bytes(double rows_in, double cin, double cout, double taps, double rows_out) {
    return 2.0 * (rows_in * cin + cout * taps * cin + rows_out * cout);
}

static int alloc_unit(ub_plan* P, ConvUnit& u) {
    const size_t out = (size_t)P->N * u.Ho() * u.Wo() * u.Co;
    UB_TRY(P->alloc(&u.a, out));
    UB_TRY(P->alloc(&u.scale, u.Co));
    UB_TRY(P->alloc(&u.shift, u.Co));
    UB_TRY(P->alloc(&u.mean, u.Co));
    UB_TRY(P->alloc(&u.rstd, u.Co));
    if (!u.first) {
        UB_TRY(P->alloc(&u.wf, (size_t)u.Co * 9 * u.Ci));
        if (P->training) UB_TRY(P->alloc(&u.wd, (size_t)u.Co * 9 * u.Ci));
    }
    if (P->training) {
        if (!u.first) {
            UB_TRY(P->alloc(&u.y, out));
            UB_TRY(P->alloc(&u.dy, out));
        }
    }
    return 0;
}

extern "C" {

int ub_plan_create(ub_plan** out, int N, int n_channels, int H, int W, int base, int levels,
                   int n_classes, int training) {
    return ub_plan_create_ex(out, N, n_channels, H, W, base, levels, n_classes, training, 0);
}
int ub_plan_create_ex(ub_plan** out, int N, int n_channels, int H, int W, int base, int levels,
                      int n_classes, int training, int bilinear) {
    if (!out) return ub::UB_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        set_last_error("no CUDA device available (libunetb200 has no CPU fallback)");
        return ub::UB_ERR_CUDA;
    }
    if (N < 1 || n_channels < 1 || levels < 2 || levels > 7 || n_classes < 1 ||
        n_classes > 8 || base < 64 || base % 64 != 0) {
        set_last_error("plan: unsupported configuration (N=%d n_channels=%d base=%d levels=%d "
                       "n_classes=%d); base must be a multiple of 64, n_classes <= 8",
                       N, n_channels, base, levels, n_classes);
        return ub::UB_ERR_UNSUPPORTED;
    }
    if ((base & (base - 1)) != 0 || (base << (levels - 1)) > 2048) {
        set_last_error("plan: base_channels=%d must be a power of two and the widest layer (%d) "
                       "at most 2048 channels", base, base << (levels - 1));
        return ub::UB_ERR_UNSUPPORTED;
    }
    ub_plan* P = new ub_plan();
    P->N = N; P->Cin = n_channels; P->H = H; P->W = W; P->base = base; P->L = levels;
    P->NC = n_classes; P->training = training != 0; P->bilinear = bilinear != 0;
    const int L = levels;
    const int UPP = P->bilinear ? 0 : 2;          // parameters of the up-sampling op of a decoder stage
    P->enc.resize(L);
    P->dec.resize(L - 1);
    P->ups.resize(L - 1);
    P->crop.resize(L - 1);
    const int nparams = 8 * L + (8 + UPP) * (L - 1) + 2;
    P->params.assign(nparams, nullptr);
    P->param_numel.assign(nparams, 0);
    P->rm.assign(2 * L + 2 * (L - 1), nullptr);
    P->rv.assign(2 * L + 2 * (L - 1), nullptr);
    P->nbt.assign(2 * L + 2 * (L - 1), nullptr);
    P->bn_mom.assign(2 * L + 2 * (L - 1), 0.1f);
    P->bn_eps.assign(2 * L + 2 * (L - 1), 1e-5f);

    auto fail = [&](int code) { delete P; return code; };
    auto set_unit_params = [&](ConvUnit& u, int pbase, int bn) {
        u.p_w = pbase; u.p_b = pbase + 1; u.p_g = pbase + 2; u.p_be = pbase + 3; u.bn = bn;
        P->param_numel[u.p_w] = (long long)u.Co * u.Ci * 9;
        P->param_numel[u.p_b] = u.Co; P->param_numel[u.p_g] = u.Co; P->param_numel[u.p_be] = u.Co;
    };

    // ---- geometry: encoder ----
    int h = H, w = W;
    for (int i = 0; i < L; ++i) {
        Block& b = P->enc[i];
        const int co = base << i;
        b.u[0].Ci = i == 0 ? n_channels : base << (i - 1);
        b.u[0].Co = co; b.u[0].Hin = h; b.u[0].Win = w; b.u[0].first = (i == 0);
        b.u[1].Ci = co; b.u[1].Co = co; b.u[1].Hin = h - 2; b.u[1].Win = w - 2;
        set_unit_params(b.u[0], 8 * i, 2 * i);
        set_unit_params(b.u[1], 8 * i + 4, 2 * i + 1);
        h -= 4; w -= 4;
        if (h < 1 || w < 1) {
            set_last_error("plan: input %dx%d too small for %d levels", H, W, L);
            return fail(ub::UB_ERR_ARG);
        }
        b.pool = i < L - 1;
        if (b.pool) { h /= 2; w /= 2; }
        if (b.pool && (h < 5 || w < 5)) {
            set_last_error("plan: input %dx%d too small for %d levels", H, W, L);
            return fail(ub::UB_ERR_ARG);
        }
    }
    // ---- geometry: decoder ----
    for (int j = 0; j < L - 1; ++j) {
        UpT& t = P->ups[j];
        Block& b = P->dec[j];
        const int cp = base << (L - 1 - j);
        // ConvTranspose2d(cp, cp / 2, 2, 2)  |  Upsample: channels unchanged (models/unet_model.py:40-46)
        t.Ci = cp; t.Co = P->bilinear ? cp : cp / 2; t.Hin = h; t.Win = w;
        const int pb = 8 * L + (8 + UPP) * j;
        if (!P->bilinear) {
            t.p_w = pb; t.p_b = pb + 1;
            P->param_numel[t.p_w] = (long long)t.Ci * t.Co * 4;
            P->param_numel[t.p_b] = t.Co;
        }
        h *= 2; w *= 2;
        const Block& sk = P->enc[L - 2 - j];
        const int sh = sk.u[1].Ho(), sw = sk.u[1].Wo();
        if (sh < h || sw < w) {
            set_last_error("plan: skip connection (%dx%d) smaller than up-sampled map (%dx%d)", sh,
                           sw, h, w);
            return fail(ub::UB_ERR_ARG);
        }
        b.two = true;
        b.u[0].Ci = cp / 2 + t.Co; b.u[0].Co = cp / 2; b.u[0].Hin = h; b.u[0].Win = w;
        b.u[1].Ci = cp / 2; b.u[1].Co = cp / 2; b.u[1].Hin = h - 2; b.u[1].Win = w - 2;
        set_unit_params(b.u[0], pb + UPP, 2 * L + 2 * j);
        set_unit_params(b.u[1], pb + UPP + 4, 2 * L + 2 * j + 1);
        h -= 4; w -= 4;
        if (h < 1 || w < 1) {
            set_last_error("plan: input %dx%d too small for %d levels", H, W, L);
            return fail(ub::UB_ERR_ARG);
        }
    }
    P->outH = h; P->outW = w;
    P->param_numel[nparams - 2] = (long long)n_classes * base;
    P->param_numel[nparams - 1] = n_classes;

    // ---- allocation ----
    size_t scratch = wce_partial_floats();
    size_t wws = 0;
    auto upd = [](size_t& a, size_t b) { if (b > a) a = b; };
    for (int i = 0; i < L; ++i) {
        Block& b = P->enc[i];
        for (int k = 0; k < 2; ++k) {
            ConvUnit& u = b.u[k];
            if (int r = alloc_unit(P, u)) return fail(r);
            if (u.first) upd(scratch, first_conv_partial_floats(u.Co));
            else upd(scratch, igemm_stats_floats(u.Co));
            upd(scratch, bn_bwd_partial_floats(u.Co));
            if (!u.first)
                upd(wws, wgrad_ws_floats(9 * u.Ci, u.Co, (long long)N * u.Ho() * u.Wo()));
        }
        if (b.pool) {
            const int ph = b.u[1].Ho() / 2, pw = b.u[1].Wo() / 2;
            if (int r = P->alloc(&b.pooled, (size_t)N * ph * pw * b.u[1].Co)) return fail(r);
            if (P->training)
                if (int r = P->alloc(&b.amax, (size_t)N * ph * pw * b.u[1].Co)) return fail(r);
        }
        if (P->training) {
            if (int r = P->alloc(&b.da0, (size_t)N * b.u[0].Ho() * b.u[0].Wo() * b.u[0].Co))
                return fail(r);
            if (i > 0)
                if (int r = P->alloc(&b.din, (size_t)N * b.u[0].Hin * b.u[0].Win * b.u[0].Ci))
                    return fail(r);
        }
    }
    for (int j = 0; j < L - 1; ++j) {
        UpT& t = P->ups[j];
        Block& b = P->dec[j];
        if (!P->bilinear) {
            if (int r = P->alloc(&t.wf, (size_t)4 * t.Co * t.Ci)) return fail(r);
            if (int r = P->alloc(&t.bias4, (size_t)4 * t.Co)) return fail(r);
        }
        if (int r = P->alloc(&t.out, (size_t)N * 4 * t.Hin * t.Win * t.Co)) return fail(r);
        if (P->training) {
            if (!P->bilinear) {
                if (int r = P->alloc(&t.wb, (size_t)4 * t.Co * t.Ci)) return fail(r);
                upd(wws, wgrad_ws_floats(4 * t.Co, t.Ci, (long long)N * t.Hin * t.Win));
            }
            if (int r = P->alloc(&t.dx, (size_t)N * t.Hin * t.Win * t.Ci)) return fail(r);
        }
        for (int k = 0; k < 2; ++k) {
            ConvUnit& u = b.u[k];
            if (int r = alloc_unit(P, u)) return fail(r);
            upd(scratch, igemm_stats_floats(u.Co));
            upd(scratch, bn_bwd_partial_floats(u.Co));
            upd(wws, wgrad_ws_floats(9 * u.Ci, u.Co, (long long)N * u.Ho() * u.Wo()));
        }
        if (P->training) {
            if (int r = P->alloc(&b.da0, (size_t)N * b.u[0].Ho() * b.u[0].Wo() * b.u[0].Co))
                return fail(r);
            if (int r = P->alloc(&b.din, (size_t)N * b.u[0].Hin * b.u[0].Win * b.u[0].Ci))
                return fail(r);
        }
        // sources of the zero-copy concat: [cropped skip, up-sampled]
        const Block& sk = P->enc[L - 2 - j];
        const int sh = sk.u[1].Ho(), sw = sk.u[1].Wo(), sc = sk.u[1].Co;
        const int uh = b.u[0].Hin, uw = b.u[0].Win;
        const int ch = (sh - uh) / 2, cw = (sw - uw) / 2;
        P->crop[j] = ch;
        View full = make_view(sk.u[1].a, N, sh, sw, sc);
        View v = full;
        v.ptr = sk.u[1].a + ((long long)ch * sw + cw) * sc;
        v.H = uh; v.W = uw;
        b.in0 = v;
        b.in1 = make_view(t.out, N, uh, uw, t.Co);
        // remember the column crop in the view itself (crop_w may differ from crop_h)
        (void)cw;
    }
    for (int i = 1; i < L; ++i) {
        Block& b = P->enc[i];
        const Block& pr = P->enc[i - 1];
        b.in0 = make_view(pr.pooled, N, b.u[0].Hin, b.u[0].Win, b.u[0].Ci);
    }
    upd(scratch, head_bwd_partial_floats(base, n_classes));
    if (int r = P->alloc(&P->scratch, scratch)) return fail(r);
    if (int r = P->alloc(&P->fc_cov, (size_t)FIRST_CONV_COV_DOUBLES)) return fail(r);
    if (P->training) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);   // lo = lowest priority (numerically greatest)
        if (cudaStreamCreateWithPriority(&P->side, cudaStreamNonBlocking, lo) != cudaSuccess ||
            cudaEventCreateWithFlags(&P->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&P->ev_join, cudaEventDisableTiming) != cudaSuccess) {
            set_last_error("plan: could not create the weight-gradient side stream");
            cudaGetLastError();
            return fail(ub::UB_ERR_CUDA);
        }
        const char* ov = getenv("UB_WGRAD_OVERLAP");
        if (ov) P->overlap = atoi(ov);
        P->wgrad_ws_floats = wws;
        if (int r = P->alloc(&P->wgrad_ws, wws)) return fail(r);
        if (int r = P->alloc(&P->head_da, (size_t)N * P->outH * P->outW * base)) return fail(r);
    }
    *out = P;
    return 0;
}

int ub_plan_destroy(ub_plan* P) {
    delete P;
    return 0;
}
int ub_plan_out_hw(const ub_plan* P, int* oh, int* ow) {
    if (!P) return ub::UB_ERR_ARG;
    if (oh) *oh = P->outH;
    if (ow) *ow = P->outW;
    return 0;
}
int ub_plan_num_params(const ub_plan* P) { return P ? (int)P->params.size() : ub::UB_ERR_ARG; }
int ub_plan_num_bn(const ub_plan* P) { return P ? (int)P->rm.size() : ub::UB_ERR_ARG; }
int64_t ub_plan_param_numel(const ub_plan* P, int i) {
    if (!P || i < 0 || i >= (int)P->param_numel.size()) return ub::UB_ERR_ARG;
    return P->param_numel[i];
}
int64_t ub_plan_device_bytes(const ub_plan* P) { return P ? (int64_t)P->bytes : 0; }

int ub_plan_bind_params(ub_plan* P, const float* const* params, int count) {
    if (!P || !params || count != (int)P->params.size()) {
        set_last_error("bind_params: expected %d parameters, got %d",
                       P ? (int)P->params.size() : -1, count);
        return ub::UB_ERR_ARG;
    }
    for (int i = 0; i < count; ++i) {
        if (!params[i]) { set_last_error("bind_params: parameter %d is null", i); return ub::UB_ERR_ARG; }
        P->params[i] = params[i];
    }
    P->packed = false;
    ++P->bind_epoch;
    return 0;
}
int ub_plan_bind_bn_buffers(ub_plan* P, float* const* rm, float* const* rv,
                            int64_t* const* nbt, int count) {
    if (!P || !rm || !rv || count != (int)P->rm.size()) {
        set_last_error("bind_bn_buffers: expected %d BatchNorm layers, got %d",
                       P ? (int)P->rm.size() : -1, count);
        return ub::UB_ERR_ARG;
    }
    for (int i = 0; i < count; ++i) {
        P->rm[i] = rm[i]; P->rv[i] = rv[i];
        P->nbt[i] = nbt ? (long long*)nbt[i] : nullptr;
    }
    ++P->bind_epoch;
    return 0;
}

int ub_plan_set_bn_config(ub_plan* P, const float* momentum, const float* eps, int count) {
    if (!P || !momentum || !eps || count != (int)P->rm.size()) {
        set_last_error("set_bn_config: expected %d BatchNorm layers, got %d",
                       P ? (int)P->rm.size() : -1, count);
        return ub::UB_ERR_ARG;
    }
    for (int i = 0; i < count; ++i) {
        if (!(momentum[i] >= 0.f && momentum[i] <= 1.f) || !(eps[i] > 0.f)) {
            set_last_error("set_bn_config: layer %d: momentum %g must lie in [0, 1] and eps %g be "
                           "positive", i, (double)momentum[i], (double)eps[i]);
            return ub::UB_ERR_ARG;
        }
    }
    for (int i = 0; i < count; ++i) { P->bn_mom[i] = momentum[i]; P->bn_eps[i] = eps[i]; }
    ++P->bind_epoch;
    return 0;
}

int ub_plan_pack_weights(ub_plan* P, void* stream) {
    if (!P) return ub::UB_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    for (size_t i = 0; i < P->params.size(); ++i)
        if (!P->params[i]) { set_last_error("pack_weights: parameters not bound"); return ub::UB_ERR_ARG; }
    auto pack_block = [&](Block& b) -> int {
        for (int k = 0; k < 2; ++k) {
            ConvUnit& u = b.u[k];
            if (u.first) continue;
            UB_TRY(launch_pack_conv3x3(P->params[u.p_w], u.Co, u.Ci, u.wf, u.wd, s));
        }
        return 0;
    };
    for (auto& b : P->enc) UB_TRY(pack_block(b));
    for (auto& b : P->dec) UB_TRY(pack_block(b));
    if (!P->bilinear)
        for (auto& t : P->ups)
            UB_TRY(launch_pack_convT(P->params[t.p_w], t.Ci, t.Co, t.wf, t.wb, P->params[t.p_b],
                                     t.bias4, s));
    P->packed = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------
static int conv_unit_forward(ub_plan* P, ConvUnit& u, const View& in0, const View* in1,
                             bf16* pooled, unsigned char* amax, cudaStream_t s) {
    const int N = P->N;
    const double pix_out = (double)N * u.Ho() * u.Wo(), pix_in = (double)N * u.Hin * u.Win;
    if (u.first) {
        ProfScope ps(P, CLS_FIRST, 2.0 * pix_out * u.Co * 9 * u.Ci * (P->training ? 2 : 1),
                     4.0 * pix_in * u.Ci * (P->training ? 2 : 1) + 2.0 * pix_out * u.Co, s);
        FirstConvDesc d;
        d.x = P->x; d.N = N; d.Ci = u.Ci; d.H = u.Hin; d.W = u.Win; d.Co = u.Co;
        d.w = P->params[u.p_w];
        if (P->training) {
            d.bias = P->params[u.p_b];
            UB_TRY(launch_first_conv_train_stats(d, P->scratch, P->params[u.p_g], P->params[u.p_be],
                                                 P->rm[u.bn], P->rv[u.bn], P->nbt[u.bn],
                                                 P->bn_mom[u.bn], P->bn_eps[u.bn], u.scale, u.shift,
                                                 u.mean, u.rstd, P->fc_cov, s));
        } else {
            d.bias = nullptr;
            UB_TRY(launch_bn_fold_eval(u.Co, P->params[u.p_b], P->params[u.p_g], P->params[u.p_be],
                                       P->rm[u.bn], P->rv[u.bn], P->bn_eps[u.bn], u.scale, u.shift, s));
        }
        return launch_first_conv_apply(d, u.scale, u.shift, u.a, s);
    }
    IgemmEpilogue e;
    memset(&e, 0, sizeof(e));
    e.ldo = u.Co;
    const double fl = 2.0 * pix_out * u.Co * 9.0 * u.Ci;
    const double by = gemm_bytes(pix_in, u.Ci, u.Co, 9, pix_out);
    if (P->training) {
        e.kind = EPI_CONV_STATS; e.out = u.y; e.bias = P->params[u.p_b]; e.stats = P->scratch;
        {
            ProfScope ps(P, CLS_FPROP, fl, by, s);
            UB_TRY(launch_igemm(in0, in1, 0, -2, 1, 9, 3, u.wf, u.Co, e, &u.info, s));
        }
        ProfScope ps(P, CLS_BN_APPLY, 0, pix_out * u.Co * (pooled ? 4.5 : 4.0), s);
        UB_TRY(launch_bn_finalize(P->scratch, u.info, u.Co, (double)u.info.M, P->params[u.p_g],
                                  P->params[u.p_be], P->rm[u.bn], P->rv[u.bn], P->nbt[u.bn],
                                  P->bn_mom[u.bn], P->bn_eps[u.bn], u.scale, u.shift, u.mean, u.rstd, s));
        if (P->fused_logits && &u == &P->dec[P->L - 2].u[1]) {   // last unit: 1x1 head fused in
            const int np = (int)P->params.size();
            return launch_bn_apply_relu_head(u.y, u.a, N, u.Ho(), u.Wo(), u.Co, u.scale, u.shift,
                                             P->NC, P->params[np - 2], P->params[np - 1],
                                             P->fused_logits, s);
        }
        return launch_bn_apply_relu(u.y, u.a, pooled, amax, N, u.Ho(), u.Wo(), u.Co, u.scale,
                                    u.shift, s);
    }
    UB_TRY(launch_bn_fold_eval(u.Co, P->params[u.p_b], P->params[u.p_g], P->params[u.p_be],
                               P->rm[u.bn], P->rv[u.bn], P->bn_eps[u.bn], u.scale, u.shift, s));
    e.kind = EPI_AFFINE_RELU; e.out = u.a; e.scale = u.scale; e.shift = u.shift;
    if (P->fused_logits && &u == &P->dec[P->L - 2].u[1]) {
        // eval, last unit (64 channels): the 1x1 head and the mask come out of the conv epilogue
        const int np = (int)P->params.size();
        e.kind = EPI_AFFINE_RELU_HEAD;
        e.head_w = P->params[np - 2]; e.head_b = P->params[np - 1];
        e.head_logits = P->fused_logits; e.head_mask = P->fused_mask; e.head_nc = P->NC;
    }
    if (e.kind == EPI_AFFINE_RELU) e.pooled = pooled;   // fused where the selected kernel can (info.pool_fused)
    {
        ProfScope ps(P, CLS_FPROP, fl, by, s);
        UB_TRY(launch_igemm(in0, in1, 0, -2, 1, 9, 3, u.wf, u.Co, e, &u.info, s));
    }
    if (pooled && !u.info.pool_fused) {
        ProfScope ps(P, CLS_BN_APPLY, 0, pix_out * u.Co * 2.5, s);
        return launch_maxpool2(u.a, pooled, N, u.Ho(), u.Wo(), u.Co, s);
    }
    return 0;
}

static int block_forward(ub_plan* P, Block& b, cudaStream_t s) {
    UB_TRY(conv_unit_forward(P, b.u[0], b.in0, b.two ? &b.in1 : nullptr, nullptr, nullptr, s));
    View mid = make_view(b.u[0].a, P->N, b.u[0].Ho(), b.u[0].Wo(), b.u[0].Co);
    return conv_unit_forward(P, b.u[1], mid, nullptr, b.pool ? b.pooled : nullptr,
                             b.pool ? b.amax : nullptr, s);
}

static int forward_impl(ub_plan* P, const float* x, float* logits, uint8_t* mask, cudaStream_t s) {
    P->x = x;
    const int L = P->L;
    static int fuse_eval = -1;
    if (fuse_eval < 0) { const char* e = getenv("UB_FUSE_EVAL_HEAD"); fuse_eval = (e && !atoi(e)) ? 0 : 1; }
    if (P->training)
        P->fused_logits = bn_apply_head_supported(P->base, P->NC) ? logits : nullptr;
    else
        P->fused_logits = (fuse_eval && P->base == 64 && P->NC <= HEAD_EPI_MAX_CLASSES) ? logits : nullptr;
    P->fused_mask = P->training ? nullptr : mask;
    for (int i = 0; i < L; ++i) UB_TRY(block_forward(P, P->enc[i], s));
    for (int j = 0; j < L - 1; ++j) {
        UpT& t = P->ups[j];
        const ConvUnit& prev = j == 0 ? P->enc[L - 1].u[1] : P->dec[j - 1].u[1];
        View xin = make_view(prev.a, P->N, t.Hin, t.Win, t.Ci);
        if (P->bilinear) {   // nn.Upsample(scale_factor=2, bilinear, align_corners=True), :41
            const double m = (double)P->N * t.Hin * t.Win * t.Ci;
            ProfScope ps(P, CLS_CT_FPROP, 0, 2.0 * m * 5, s);
            UB_TRY(launch_upsample2x_fwd(xin, t.out, s));
            UB_TRY(block_forward(P, P->dec[j], s));
            continue;
        }
        IgemmEpilogue e;
        memset(&e, 0, sizeof(e));
        e.kind = EPI_CONVT; e.bias = t.bias4;
        e.ct_dst = make_view(t.out, P->N, 2 * t.Hin, 2 * t.Win, t.Co);
        {
            const double m = (double)P->N * t.Hin * t.Win;
            ProfScope ps(P, CLS_CT_FPROP, 2.0 * m * 4 * t.Co * t.Ci,
                         gemm_bytes(m, t.Ci, 4.0 * t.Co, 1, m), s);
            UB_TRY(launch_igemm(xin, nullptr, 0, 0, 1, 1, 1, t.wf, 4 * t.Co, e, nullptr, s));
        }
        UB_TRY(block_forward(P, P->dec[j], s));
    }
    if (P->fused_logits) return 0;
    const ConvUnit& last = P->dec[L - 2].u[1];
    const int np = (int)P->params.size();
    ProfScope ps(P, CLS_HEAD, 2.0 * P->N * P->outH * P->outW * P->base * P->NC,
                 (double)P->N * P->outH * P->outW * (2.0 * P->base + 4.0 * P->NC), s);
    return launch_head_fwd(last.a, P->N, P->outH, P->outW, P->base, P->NC, P->params[np - 2],
                           P->params[np - 1], logits, P->training ? nullptr : mask, s);
}

static bool eval_graph_enabled() {   // UB_EVAL_GRAPH=0 disables the captured eval forward
    static const bool on = [] { const char* e = getenv("UB_EVAL_GRAPH"); return !(e && e[0] == '0'); }();
    return on;
}

int ub_plan_forward(ub_plan* P, const float* x, float* logits, uint8_t* mask, void* stream) {
    if (!P || !x || !logits) { set_last_error("forward: null pointer"); return ub::UB_ERR_ARG; }
    if (!P->packed) { set_last_error("forward: call ub_plan_pack_weights first"); return ub::UB_ERR_ARG; }
    for (size_t i = 0; i < P->rm.size(); ++i)
        if (!P->rm[i] || !P->rv[i]) { set_last_error("forward: BN buffers not bound"); return ub::UB_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    bool graphable = !P->training && !P->g_disabled && eval_graph_enabled() && !P->prof_on && !nvtx_on();
    if (graphable) {
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(s, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) {
            cudaGetLastError();
            graphable = false;       // the caller is capturing a graph of its own
        }
    }
    if (!graphable) return forward_impl(P, x, logits, mask, s);
    const bool same = x == P->g_x && logits == P->g_logits && mask == P->g_mask &&
                      P->bind_epoch == P->g_epoch;
    if (same && P->fwd_graph) {
        UB_CHECK_CUDA(cudaGraphLaunch(P->fwd_graph, s));
        ub::count_launch((int)P->g_nodes);
        ++P->g_replays;
        return 0;
    }
    if (!same) {
        P->drop_graph();
        P->g_x = x; P->g_logits = logits; P->g_mask = mask; P->g_epoch = P->bind_epoch;
    }
    if (++P->g_calls < 2) return forward_impl(P, x, logits, mask, s);   // first sight: plain launches
    // Capture on an internal stream: nothing executes during capture, and the caller's stream may be
    // the legacy default stream, which cannot be captured. The graph is launched on the caller's.
    if (!P->cap_stream &&
        cudaStreamCreateWithFlags(&P->cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        P->cap_stream = nullptr;
        P->g_disabled = true;
        return forward_impl(P, x, logits, mask, s);
    }
    if (cudaStreamBeginCapture(P->cap_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        P->g_disabled = true;
        return forward_impl(P, x, logits, mask, s);
    }
    const long long before = ub::launch_count();
    const int rc = forward_impl(P, x, logits, mask, P->cap_stream);
    cudaGraph_t g = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(P->cap_stream, &g);
    if (rc != 0 || ce != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        P->drop_graph();
        P->g_disabled = true;
        if (rc != 0) return rc;
        return forward_impl(P, x, logits, mask, s);      // capture refused: stay on plain launches
    }
    P->g_nodes = ub::launch_count() - before;
    const cudaError_t ie = cudaGraphInstantiate(&P->fwd_graph, g, 0);
    cudaGraphDestroy(g);
    if (ie != cudaSuccess) {
        cudaGetLastError();
        P->fwd_graph = nullptr;
        P->g_disabled = true;
        ub::count_launch(-(int)P->g_nodes);
        return forward_impl(P, x, logits, mask, s);
    }
    UB_CHECK_CUDA(cudaGraphLaunch(P->fwd_graph, s));
    ++P->g_replays;
    return 0;
}
int64_t ub_plan_graph_replays(const ub_plan* P) { return P ? (int64_t)P->g_replays : 0; }

// ------------------------------------------------------------------------------------------------
int ub_plan_num_stages(const ub_plan* P) { return P ? 2 * P->L - 1 : ub::UB_ERR_ARG; }

int ub_plan_stage_params(const ub_plan* P, int stage, int* first, int* count) {
    if (!P || stage < 0 || stage >= 2 * P->L - 1) return ub::UB_ERR_ARG;
    const int L = P->L;
    int f, c;
    if (stage < L - 1) {
        const int j = L - 2 - stage;  // decoder block index
        const int per = P->bilinear ? 8 : 10;
        f = 8 * L + per * j;
        c = per + (stage == 0 ? 2 : 0);  // the head's two parameters follow the last up block
    } else {
        const int i = 2 * L - 2 - stage;  // encoder block index
        f = 8 * i;
        c = 8;
    }
    if (first) *first = f;
    if (count) *count = c;
    return 0;
}

// Stream for a weight-gradient launch: the side stream, ordered after everything enqueued on `s` so
// far (its inputs dy / d(concat) were just produced there), or `s` itself when overlap is off.
static cudaStream_t wgrad_stream(ub_plan* P, cudaStream_t s) {
    if (P->overlap == 0 || P->prof_on || !P->side) return s;
    cudaEventRecord(P->ev_fork, s);
    cudaStreamWaitEvent(P->side, P->ev_fork, 0);
    P->side_used = true;
    return P->side;
}
static void wgrad_join(ub_plan* P, cudaStream_t s) {
    if (!P->side_used) return;
    cudaEventRecord(P->ev_join, P->side);
    cudaStreamWaitEvent(s, P->ev_join, 0);
    P->side_used = false;
}

// The fused reduction makes the data-gradient epilogue ~5x longer (y prefetch, mask, two column sums per
// 32-column chunk). It is free only where a tile's MMAs outlast it, i.e. for long K: measured on the
// N = 16 step, fusing EVERY eligible layer moved 0.48 ms out of the BN-backward class but added 0.80 ms
// to the data gradients (the 128 / 64-column two-row tiles of the shallow layers become epilogue-bound),
// so a layer is fused only from UB_BNRED_MINKB k-blocks of 64 per tile on (UB_FUSE_BNRED=0: never).
static int bnred_min_kblocks() {
    static const int v = [] {
        const char* off = getenv("UB_FUSE_BNRED");
        if (off && off[0] == '0') return 1 << 30;
        const char* e = getenv("UB_BNRED_MINKB");
        return e ? atoi(e) : 36;
    }();
    return v;
}
static bool fuse_bnred(int kblocks) { return kblocks >= bnred_min_kblocks(); }

// UB_WGRAD_DEFER=1: enqueue a unit's weight gradient AFTER its data gradient, so that on the side
// stream it becomes runnable together with the next (HBM-bound) BN-backward instead of competing with
// the tensor-bound data gradient of the same unit.
static bool wgrad_defer() {
    static const bool v = [] { const char* e = getenv("UB_WGRAD_DEFER"); return e && e[0] == '1'; }();
    return v;
}

struct Upstream {
    const IgemmLaunchInfo* fused = nullptr;   // reduce pass of the block's second unit already done
    bool pool_skip = false;
    View g{}, gp{}, gs{};
    int crop_h = 0, crop_w = 0;
    bool has_skip = false;
    const unsigned char* amax = nullptr;
};

static int block_backward(ub_plan* P, Block& b, const Upstream& up, float* const* grads,
                          cudaStream_t s) {
    const int N = P->N;
    ConvUnit& u1 = b.u[1];
    ConvUnit& u0 = b.u[0];
    // ---- second conv unit ----
    BnBwdDesc d;
    memset(&d, 0, sizeof(d));
    d.y = u1.y; d.N = N; d.H = u1.Ho(); d.W = u1.Wo(); d.C = u1.Co;
    d.scale = u1.scale; d.shift = u1.shift; d.mean = u1.mean; d.rstd = u1.rstd;
    d.pool_skip = up.pool_skip; d.g = up.g; d.gp = up.gp; d.gs = up.gs;
    d.crop_h = up.crop_h; d.crop_w = up.crop_w; d.has_skip = up.has_skip;
    d.amax = up.amax;
    d.partial = P->scratch; d.dgamma = grads[u1.p_g]; d.dbeta = grads[u1.p_be]; d.dy = u1.dy;
    const double po1 = (double)N * u1.Ho() * u1.Wo(), pi1 = (double)N * u1.Hin * u1.Win;
    const double fl1 = 2.0 * po1 * u1.Co * 9.0 * u1.Ci;
    {
        ProfScope ps(P, CLS_BN_BWD, 0, po1 * u1.Co * (up.fused ? 6.0 : 10.0), s);
        UB_TRY(launch_bn_bwd(d, s, up.fused));
    }
    View a0 = make_view(u0.a, N, u0.Ho(), u0.Wo(), u0.Co);
    auto wgrad1 = [&]() -> int {
        cudaStream_t ws = wgrad_stream(P, s);
        ProfScope ps(P, CLS_WGRAD, fl1, 2.0 * (pi1 * u1.Ci + po1 * u1.Co) + 4.0 * 9 * u1.Ci * u1.Co, ws);
        // both conv biases of the block sit ahead of a BatchNorm: analytically zero gradients
        return launch_wgrad(a0, nullptr, 0, -2, 1, 9, 3, u1.dy, u1.Co, u1.Co, P->wgrad_ws,
                            P->wgrad_ws_floats, grads[u1.p_w], ws, grads[u1.p_b], u1.Co,
                            grads[u0.p_b], u0.Co);
    };
    const bool defer = wgrad_defer();
    if (!defer) UB_TRY(wgrad1());
    // The data gradient of the second unit IS the upstream gradient of the first unit's BN + ReLU: its
    // epilogue also does that layer's reduce pass (sum dyh, sum dyh (y - mean)) from the stored y.
    const bool fuse0 = !u0.first && fuse_bnred(9 * u1.Co / 64);
    IgemmLaunchInfo info0{};
    {
        IgemmEpilogue e;
        memset(&e, 0, sizeof(e));
        e.kind = EPI_STORE; e.out = b.da0; e.ldo = u1.Ci;
        if (fuse0) {
            e.kind = EPI_STORE_BNRED;
            e.red_y = u0.y; e.scale = u0.scale; e.shift = u0.shift; e.red_mean = u0.mean;
            e.stats = P->scratch;
        }
        View dyv = make_view(u1.dy, N, u1.Ho(), u1.Wo(), u1.Co);
        ProfScope ps(P, CLS_DGRAD, fl1, gemm_bytes(po1, u1.Co, u1.Ci, 9, pi1) + (fuse0 ? 2.0 * pi1 * u1.Ci : 0.0), s);
        UB_TRY(launch_igemm(dyv, nullptr, -2, 0, 1, 9, 3, u1.wd, u1.Ci, e, &info0, s));
    }
    if (defer) UB_TRY(wgrad1());
    // ---- first conv unit ----
    View g0 = make_view(b.da0, N, u0.Ho(), u0.Wo(), u0.Co);
    if (u0.first) {
        FirstConvDesc f;
        f.x = P->x; f.N = N; f.Ci = u0.Ci; f.H = u0.Hin; f.W = u0.Win; f.Co = u0.Co;
        f.w = P->params[u0.p_w]; f.bias = P->params[u0.p_b];
        const double po = (double)N * u0.Ho() * u0.Wo();
        ProfScope ps(P, CLS_FIRST, 2.0 * po * u0.Co * 9 * u0.Ci,
                     4.0 * N * u0.Hin * u0.Win * u0.Ci + 4.0 * po * u0.Co, s);
        return launch_first_conv_bwd(f, u0.scale, u0.shift, u0.mean, u0.rstd, g0, u0.a, P->fc_cov,
                                     P->scratch, grads[u0.p_g], grads[u0.p_be], grads[u0.p_w], s);
    }
    memset(&d, 0, sizeof(d));
    d.y = u0.y; d.N = N; d.H = u0.Ho(); d.W = u0.Wo(); d.C = u0.Co;
    d.scale = u0.scale; d.shift = u0.shift; d.mean = u0.mean; d.rstd = u0.rstd;
    d.pool_skip = false; d.g = g0;
    d.partial = P->scratch; d.dgamma = grads[u0.p_g]; d.dbeta = grads[u0.p_be]; d.dy = u0.dy;
    const double po0 = (double)N * u0.Ho() * u0.Wo(), pi0 = (double)N * u0.Hin * u0.Win;
    const double fl0 = 2.0 * po0 * u0.Co * 9.0 * u0.Ci;
    {
        ProfScope ps(P, CLS_BN_BWD, 0, po0 * u0.Co * (fuse0 ? 6.0 : 10.0), s);
        UB_TRY(launch_bn_bwd(d, s, fuse0 ? &info0 : nullptr));
    }
    auto wgrad0 = [&]() -> int {
        cudaStream_t ws = wgrad_stream(P, s);
        ProfScope ps(P, CLS_WGRAD, fl0, 2.0 * (pi0 * u0.Ci + po0 * u0.Co) + 4.0 * 9 * u0.Ci * u0.Co, ws);
        return launch_wgrad(b.in0, b.two ? &b.in1 : nullptr, 0, -2, 1, 9, 3, u0.dy, u0.Co, u0.Co,
                            P->wgrad_ws, P->wgrad_ws_floats, grads[u0.p_w], ws);
    };
    if (!defer) UB_TRY(wgrad0());
    {
        IgemmEpilogue e;
        memset(&e, 0, sizeof(e));
        e.kind = EPI_STORE; e.out = b.din; e.ldo = u0.Ci;
        View dyv = make_view(u0.dy, N, u0.Ho(), u0.Wo(), u0.Co);
        ProfScope ps(P, CLS_DGRAD, fl0, gemm_bytes(po0, u0.Co, u0.Ci, 9, pi0), s);
        UB_TRY(launch_igemm(dyv, nullptr, -2, 0, 1, 9, 3, u0.wd, u0.Ci, e, nullptr, s));
    }
    if (defer) UB_TRY(wgrad0());
    return 0;
}

static int backward_stage_impl(ub_plan* P, int stage, const float* dlogits, float* const* grads,
                               void* stream) {
    if (!P || !grads) { set_last_error("backward: null pointer"); return ub::UB_ERR_ARG; }
    if (!P->training) {
        set_last_error("backward: plan was created for inference (training=0)");
        return ub::UB_ERR_ARG;
    }
    if (stage < 0 || stage >= 2 * P->L - 1) { set_last_error("backward: bad stage"); return ub::UB_ERR_ARG; }
    if (!P->x) { set_last_error("backward: no forward pass recorded"); return ub::UB_ERR_ARG; }
    int first = 0, cnt = 0;
    ub_plan_stage_params(P, stage, &first, &cnt);
    for (int i = first; i < first + cnt; ++i)
        if (!grads[i]) { set_last_error("backward: gradient buffer %d is null", i); return ub::UB_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    const int L = P->L, N = P->N;

    if (stage < L - 1) {
        const int j = L - 2 - stage;
        Block& b = P->dec[j];
        UpT& t = P->ups[j];
        Upstream up;
        if (stage == 0) {
            if (!dlogits) { set_last_error("backward: dlogits is null"); return ub::UB_ERR_ARG; }
            const int np = (int)P->params.size();
            ProfScope ps(P, CLS_HEAD, 4.0 * N * P->outH * P->outW * P->base * P->NC,
                         (double)N * P->outH * P->outW * (4.0 * P->base + 4.0 * P->NC), s);
            UB_TRY(launch_head_bwd(dlogits, b.u[1].a, N, P->outH, P->outW, P->base, P->NC,
                                   P->params[np - 2], P->head_da, P->scratch, grads[np - 2],
                                   grads[np - 1], s));
            up.g = make_view(P->head_da, N, P->outH, P->outW, P->base);
        } else {
            const UpT& nt = P->ups[j + 1];
            up.g = make_view(nt.dx, N, nt.Hin, nt.Win, nt.Ci);
            if (P->dx_fused) up.fused = &P->dx_info;    // left by the previous stage's transposed conv
        }
        P->dx_fused = false;
        UB_TRY(block_backward(P, b, up, grads, s));
        // transposed conv: the second channel range of d(concat) is d(up)
        const int cs = b.u[0].Ci - t.Co;  // skip channels
        View dup = make_view(b.din, N, b.u[0].Hin, b.u[0].Win, b.u[0].Ci);
        dup.ptr = b.din + cs;
        dup.C = t.Co;
        if (P->bilinear) {   // adjoint of the bilinear up-sampling; no parameters
            const double m = (double)N * t.Hin * t.Win * t.Ci;
            ProfScope ps(P, CLS_CT_DGRAD, 0, 2.0 * m * 5, s);
            return launch_upsample2x_bwd(dup, t.dx, s);
        }
        const ConvUnit& prev = j == 0 ? P->enc[L - 1].u[1] : P->dec[j - 1].u[1];
        IgemmEpilogue e;
        memset(&e, 0, sizeof(e));
        e.kind = EPI_STORE; e.out = t.dx; e.ldo = t.Ci;
        const bool fuse_dx = fuse_bnred(4 * t.Co / 64);
        if (fuse_dx) {
            // t.dx is the upstream gradient of the previous block's second BN + ReLU (no pool, no skip
            // in between): reduce pass in this epilogue, consumed by the next backward stage
            e.kind = EPI_STORE_BNRED;
            e.red_y = prev.y; e.scale = prev.scale; e.shift = prev.shift; e.red_mean = prev.mean;
            e.stats = P->scratch;
        }
        const double mt = (double)N * t.Hin * t.Win, flt = 2.0 * mt * 4 * t.Co * t.Ci;
        {
            ProfScope ps(P, CLS_CT_DGRAD, flt, gemm_bytes(mt, 4.0 * t.Co, t.Ci, 1, mt), s);
            UB_TRY(launch_igemm(dup, nullptr, 0, -1, 2, 4, 2, t.wb, t.Ci, e, &P->dx_info, s));
        }
        P->dx_fused = fuse_dx;
        cudaStream_t ws = wgrad_stream(P, s);   // after the data gradient that produced d(concat)
        ProfScope ps(P, CLS_CT_WGRAD, flt, 2.0 * mt * (4.0 * t.Co + t.Ci) + 16.0 * t.Co * t.Ci, ws);
        // the transposed-conv bias is removed by the following BatchNorm: zero gradient
        return launch_wgrad(dup, nullptr, 0, -1, 2, 4, 2, prev.a, t.Ci, t.Ci, P->wgrad_ws,
                            P->wgrad_ws_floats, grads[t.p_w], ws, grads[t.p_b], t.Co);
    }
    const int i = 2 * L - 2 - stage;
    Block& b = P->enc[i];
    Upstream up;
    if (i == L - 1) {
        const UpT& t0 = P->ups[0];
        up.g = make_view(t0.dx, N, t0.Hin, t0.Win, t0.Ci);
        if (P->dx_fused) up.fused = &P->dx_info;
        P->dx_fused = false;
    } else {
        const Block& nx = P->enc[i + 1];
        const int j = L - 2 - i;
        const Block& db = P->dec[j];
        up.pool_skip = true;
        up.gp = make_view(nx.din, N, nx.u[0].Hin, nx.u[0].Win, nx.u[0].Ci);
        View gs = make_view(db.din, N, db.u[0].Hin, db.u[0].Win, db.u[0].Ci);
        gs.C = b.u[1].Co;  // first channel range of d(concat) = gradient of the cropped skip
        up.gs = gs;
        up.has_skip = true;
        up.amax = b.amax;
        up.crop_h = (b.u[1].Ho() - db.u[0].Hin) / 2;
        up.crop_w = (b.u[1].Wo() - db.u[0].Win) / 2;
    }
    return block_backward(P, b, up, grads, s);
}


int ub_plan_backward_stage(ub_plan* P, int stage, const float* dlogits, float* const* grads,
                           void* stream) {
    const int rc = backward_stage_impl(P, stage, dlogits, grads, stream);
    // weight gradients of this stage ran on the side stream: join where the caller needs them
    if (P && (rc != 0 || P->overlap == 1 || stage == 2 * P->L - 2)) wgrad_join(P, (cudaStream_t)stream);
    return rc;
}
// Makes `stream` (e.g. a communication stream) wait for the weight gradients enqueued so far on the
// internal stream, without stalling the stream the backward runs on.
int ub_plan_join_side(ub_plan* P, void* stream) {
    if (!P) return ub::UB_ERR_ARG;
    if (!P->side || P->overlap == 0) return 0;
    UB_CHECK_CUDA(cudaEventRecord(P->ev_join, P->side));
    UB_CHECK_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, P->ev_join, 0));
    return 0;
}
int ub_plan_set_overlap(ub_plan* P, int mode) {
    if (!P || mode < 0 || mode > 2) { set_last_error("set_overlap: mode must be 0, 1 or 2"); return ub::UB_ERR_ARG; }
    if (!getenv("UB_WGRAD_OVERLAP")) P->overlap = mode;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Fused SGD step over all parameters + refresh of the packed bf16 operands (sgd.cuh).
int ub_plan_sgd_step(ub_plan* P, const float* const* grads, float* const* momentum_bufs, float lr,
                     float momentum, float dampening, float weight_decay, int nesterov,
                     int first_step, void* stream) {
    if (!P || !grads) { set_last_error("sgd_step: null pointer"); return ub::UB_ERR_ARG; }
    const int np = (int)P->params.size();
    if (momentum != 0.f && !momentum_bufs) {
        set_last_error("sgd_step: momentum buffers required");
        return ub::UB_ERR_ARG;
    }
    for (int i = 0; i < np; ++i)
        if (!P->params[i] || !grads[i] || (momentum != 0.f && !momentum_bufs[i])) {
            set_last_error("sgd_step: parameter, gradient or momentum buffer %d is null", i);
            return ub::UB_ERR_ARG;
        }
    std::vector<SgdTensor> all(np);
    auto base = [&](int i) -> SgdTensor& {
        SgdTensor& t = all[i];
        memset(&t, 0, sizeof(t));
        t.p = const_cast<float*>(P->params[i]);
        t.g = grads[i];
        t.buf = momentum != 0.f ? momentum_bufs[i] : nullptr;
        t.kind = SGD_PLAIN;
        t.n = (int)P->param_numel[i];
        return t;
    };
    for (int i = 0; i < np; ++i) base(i);
    auto conv_units = [&](Block& b) {
        for (int k = 0; k < 2; ++k) {
            ConvUnit& u = b.u[k];
            if (u.first) continue;
            SgdTensor& t = all[u.p_w];
            t.kind = SGD_CONV3; t.d0 = u.Co; t.d1 = u.Ci; t.T = 9; t.outA = u.wf; t.outB = u.wd;
        }
    };
    for (auto& b : P->enc) conv_units(b);
    for (auto& b : P->dec) conv_units(b);
    for (auto& u : P->ups) {
        if (P->bilinear) break;
        SgdTensor& t = all[u.p_w];
        t.kind = SGD_CONVT; t.d0 = u.Ci; t.d1 = u.Co; t.T = 4; t.outA = u.wb; t.outB = u.wf;
        SgdTensor& tb = all[u.p_b];
        tb.kind = SGD_CONVT_BIAS; tb.d0 = u.Co; tb.bias4 = u.bias4;
    }
    cudaStream_t s = (cudaStream_t)stream;
    // algorithmic bytes: p, grad, momentum read + p, momentum written (fp32) + two bf16 operand layouts
    double total = 0.0, packed = 0.0;
    for (int i = 0; i < np; ++i) {
        total += (double)all[i].n;
        if (all[i].kind == SGD_CONV3 || all[i].kind == SGD_CONVT) packed += (double)all[i].n;
    }
    static const int sgd_vec = [] { const char* e = getenv("UB_SGD_VEC"); return (e && e[0] == '0') ? 0 : 1; }();
    ProfScope ps(P, CLS_SGD, 0.0, total * 4.0 * (momentum != 0.f ? 5.0 : 3.0) + packed * 4.0, s);
    for (int i0 = 0; i0 < np; i0 += SGD_MAX_TENSORS) {
        SgdBatch B;
        memset(&B, 0, sizeof(B));
        B.count = np - i0 < SGD_MAX_TENSORS ? np - i0 : SGD_MAX_TENSORS;
        B.lr = lr; B.momentum = momentum; B.dampening = dampening; B.weight_decay = weight_decay;
        B.nesterov = nesterov; B.first_step = first_step;
        B.vec = sgd_vec;
        int blocks = 0;
        for (int j = 0; j < B.count; ++j) {
            B.t[j] = all[i0 + j];
            B.t[j].block_begin = blocks;
            const SgdTensor& t = B.t[j];
            blocks += (t.kind == SGD_CONV3 || t.kind == SGD_CONVT) ? (t.d0 / 32) * (t.d1 / 32)
                                                                     : (t.n + 2047) / 2048;
        }
        UB_LAUNCH_NC(sgd_fused_kernel, blocks, 256, 0, s, B);
        UB_POST_LAUNCH();
    }
    P->packed = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------
int64_t ub_launch_count(void) { return (int64_t)ub::launch_count(); }

int ub_plan_profile_enable(ub_plan* P, int on) {
    if (!P) return ub::UB_ERR_ARG;
    for (auto& r : P->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    P->prof.clear();
    P->prof_on = on != 0;
    return 0;
}
int ub_plan_profile_classes(void) { return CLS_COUNT; }
const char* ub_plan_profile_class_name(int cls) {
    return (cls >= 0 && cls < CLS_COUNT) ? kClassNames[cls] : "";
}
int ub_plan_profile_collect(ub_plan* P, double* ms, double* flops, double* bytes, int* launches) {
    if (!P || !ms || !flops || !bytes || !launches) return ub::UB_ERR_ARG;
    for (int c = 0; c < CLS_COUNT; ++c) { ms[c] = 0; flops[c] = 0; bytes[c] = 0; launches[c] = 0; }
    for (auto& r : P->prof) {
        UB_CHECK_CUDA(cudaEventSynchronize(r.e1));
        float t = 0.f;
        UB_CHECK_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
        ms[r.cls] += t; flops[r.cls] += r.flops; bytes[r.cls] += r.bytes; launches[r.cls] += 1;
        cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
    }
    P->prof.clear();
    return 0;
}

}  // extern "C"
