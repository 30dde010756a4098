// Host-side launchers of the tcgen05 implicit-GEMM kernels (see igemm.cuh).
#include "igemm.cuh"
#include "igemm_rr.cuh"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <utility>

#include "elementwise.cuh"
#include "tmaps.cuh"
#include "ub_internal.h"

namespace ub {

static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

static int pick_bn(int ncols) {
    if (ncols % 256 == 0) return 256;
    if (ncols % 128 == 0) return 128;
    if (ncols % 64 == 0) return 64;
    return 0;
}

size_t igemm_stats_floats(int ncols) {
    int bn = pick_bn(ncols);
    if (!bn) return 0;
    return (size_t)num_sms() * 2 * bn;   // one [2][BN] row per CTA
}

bool pdl_allow(bool gemm, unsigned blocks) {
    // Default 0: with every edge of the step programmatic (mode 1), early-resident CTAs of the next
    // kernel take registers / thread slots from the HBM-bound kernels and the step got 1 % SLOWER
    // (15.55 -> 15.75 ms, same box).
    static int mode = -1;
    static thread_local int last_kind = 0;   // 0 tiny, 1 large elementwise, 2 tensor-core kernel
    if (mode < 0) { const char* e = getenv("UB_PDL"); mode = e ? atoi(e) : 0; }
    const int kind = gemm ? 2 : (blocks >= (unsigned)num_sms() ? 1 : 0);
    const int prev = last_kind;
    last_kind = kind;
    if (mode == 1) return true;
    if (mode == 2) return !(kind == 2 && prev == 1);
    if (mode == 3) return kind == 0 || prev == 0;   // only the edges into / out of the tiny finalisation kernels
    return false;
}
// cta_group::2 pairs: on unless UB_PAIR=0
static bool pairs_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("UB_PAIR"); on = (e && !atoi(e)) ? 0 : 1; }
    return on != 0;
}

// Pairs pay off for BN >= 128 with at least 16 k-blocks per tile (short-K tiles are epilogue-bound) and
// two waves of tiles. BN = 64: an M128 x N64 MMA reads 6 KB of shared memory for 32 cycles of math; a
// pair reads 5 KB per CTA (each CTA stages half of the weight rows), which wins once the weights are
// streamed per tile, i.e. for C_in >= 128 (18+ k-blocks: up4.a fprop 0.392 -> 0.330 ms, d1.a dgrad
// 0.175 -> 0.157 ms at N = 16); the C_in = 64 layers keep the single-CTA resident-weight kernel, which
// the pair form only equals. UB_PAIR_MINBN / UB_PAIR64_MINKB override.
static int pick_cg(int BN, int kblocks, long long tiles) {
    static int minkb = -1, minbn = -1, minkb64 = -1;
    if (minkb < 0) { const char* e = getenv("UB_PAIR_MINKB"); minkb = e ? atoi(e) : 16; }
    if (minbn < 0) { const char* e = getenv("UB_PAIR_MINBN"); minbn = e ? atoi(e) : 64; }
    if (minkb64 < 0) { const char* e = getenv("UB_PAIR64_MINKB"); minkb64 = e ? atoi(e) : 18; }
    if (!pairs_enabled() || BN < minbn || tiles < 2LL * num_sms()) return 1;
    return kblocks >= (BN == 64 ? minkb64 : minkb) ? 2 : 1;
}

template <int BN, int EPI, int CG>
static int launch_igemm_t(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                          const IgemmParams& p, int grid, cudaStream_t stream) {
    using Cfg = IgemmCfg<BN, CG>;
    static bool attr_set = false;
    if (!attr_set) {
        UB_CHECK_CUDA(cudaFuncSetAttribute(igemm_kmajor_kernel<BN, EPI, CG>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Cfg::SMEM_BYTES));
        attr_set = true;
    }
    UB_CHECK_CUDA(ub_launch(igemm_kmajor_kernel<BN, EPI, CG>, dim3(grid), dim3(igemm_threads(BN)),
                                   Cfg::SMEM_BYTES, stream, CG, a0, a1, b, p));
    UB_POST_LAUNCH();
    return UB_OK;
}

template <int BN, int CG>
static int launch_igemm_bn(int epi, const CUtensorMap& a0, const CUtensorMap& a1,
                           const CUtensorMap& b, const IgemmParams& p, int grid,
                           cudaStream_t stream) {
    switch (epi) {
        case EPI_CONV_STATS: return launch_igemm_t<BN, EPI_CONV_STATS, CG>(a0, a1, b, p, grid, stream);
        case EPI_STORE: return launch_igemm_t<BN, EPI_STORE, CG>(a0, a1, b, p, grid, stream);
        case EPI_STORE_BNRED: return launch_igemm_t<BN, EPI_STORE_BNRED, CG>(a0, a1, b, p, grid, stream);
        case EPI_AFFINE_RELU: return launch_igemm_t<BN, EPI_AFFINE_RELU, CG>(a0, a1, b, p, grid, stream);
        case EPI_CONVT: return launch_igemm_t<BN, EPI_CONVT, CG>(a0, a1, b, p, grid, stream);
        case EPI_AFFINE_RELU_HEAD:
            if (BN == 64 && CG == 1)
                return launch_igemm_t<64, EPI_AFFINE_RELU_HEAD, 1>(a0, a1, b, p, grid, stream);
            break;
    }
    set_last_error("unknown epilogue kind %d", epi);
    return UB_ERR_ARG;
}

static int check_view(const View& v, const char* what) {
    if (v.C % 64 != 0 || v.C <= 0) {
        set_last_error("%s: channel count %d is not a positive multiple of 64", what, v.C);
        return UB_ERR_UNSUPPORTED;
    }
    if ((reinterpret_cast<uintptr_t>(v.ptr) & 15) || (v.sW % 8) || (v.sH % 8) || (v.sN % 8)) {
        set_last_error("%s: view is not 16-byte aligned", what);
        return UB_ERR_ARG;
    }
    return UB_OK;
}

// ---- row-run variant (igemm_rr.cuh) -----------------------------------------------------------
template <int BN, int EPI, int CG, bool WRES, int ROWS>
static int launch_rowrun_t(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                           const RowRunParams& p, int grid, cudaStream_t stream) {
    using Cfg = RowRunCfg<BN, CG, WRES, ROWS>;
    static bool attr_set = false;
    if (!attr_set) {
        UB_CHECK_CUDA(cudaFuncSetAttribute(igemm_rowrun_kernel<BN, EPI, CG, WRES, ROWS>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Cfg::SMEM_BYTES));
        attr_set = true;
    }
    UB_CHECK_CUDA(ub_launch(igemm_rowrun_kernel<BN, EPI, CG, WRES, ROWS>, dim3(grid),
                                   dim3(igemm_threads(BN)), Cfg::SMEM_BYTES, stream, CG, a0, a1, b, p));
    UB_POST_LAUNCH();
    return UB_OK;
}
template <int BN, int CG, bool WRES = false, int ROWS = 1>
static int launch_rowrun_bn(int epi, const CUtensorMap& a0, const CUtensorMap& a1,
                            const CUtensorMap& b, const RowRunParams& p, int grid,
                            cudaStream_t stream) {
    switch (epi) {
        case EPI_CONV_STATS:
            return launch_rowrun_t<BN, EPI_CONV_STATS, CG, WRES, ROWS>(a0, a1, b, p, grid, stream);
        case EPI_STORE: return launch_rowrun_t<BN, EPI_STORE, CG, WRES, ROWS>(a0, a1, b, p, grid, stream);
        case EPI_STORE_BNRED:
            return launch_rowrun_t<BN, EPI_STORE_BNRED, CG, WRES, ROWS>(a0, a1, b, p, grid, stream);
        case EPI_AFFINE_RELU:
            return launch_rowrun_t<BN, EPI_AFFINE_RELU, CG, WRES, ROWS>(a0, a1, b, p, grid, stream);
        case EPI_AFFINE_RELU_HEAD:
            if (BN == 64 && CG == 1)
                return launch_rowrun_t<64, EPI_AFFINE_RELU_HEAD, 1, WRES, ROWS>(a0, a1, b, p, grid, stream);
            break;
    }
    set_last_error("row-run: unsupported epilogue kind %d", epi);
    return UB_ERR_ARG;
}
// Row-run pays off when a 128-pixel tile of one output row is mostly valid pixels.
static bool rowrun_eligible(int Wo, int taps, int tstride, int epi_kind) {
    if (taps != 9 || tstride != 1 || epi_kind == EPI_CONVT) return false;
    static int disabled = -1;
    if (disabled < 0) { const char* e = getenv("UB_NO_ROWRUN"); disabled = (e && atoi(e)) ? 1 : 0; }
    if (disabled) return false;
    const int tiles = (Wo + 127) / 128;
    return Wo * 100 >= tiles * 128 * 80;
}

int launch_igemm(const View& src0, const View* src1, int lower, int upper, int tstride, int taps,
                 int tapw, const __nv_bfloat16* wB, int ncols, const IgemmEpilogue& epi,
                 IgemmLaunchInfo* info, cudaStream_t stream) {
    UB_TRY(check_view(src0, "igemm source 0"));
    if (src1) {
        UB_TRY(check_view(*src1, "igemm source 1"));
        if (src1->N != src0.N || src1->H != src0.H || src1->W != src0.W) {
            set_last_error("igemm: concat sources differ in shape");
            return UB_ERR_ARG;
        }
    }
    const int BN = pick_bn(ncols);
    if (!BN) {
        set_last_error("igemm: %d output columns is not a multiple of 64", ncols);
        return UB_ERR_UNSUPPORTED;
    }
    if (epi.kind == EPI_STORE_BNRED && (!epi.red_y || !epi.red_mean || !epi.scale || !epi.shift ||
                                        !epi.stats || epi.ldo != ncols)) {
        set_last_error("igemm: the fused BN-backward reduction needs y, mean, scale, shift, a partial "
                       "buffer and a dense output (ldo == columns)");
        return UB_ERR_ARG;
    }
    if (epi.kind == EPI_AFFINE_RELU_HEAD &&
        (ncols != 64 || epi.head_nc < 1 || epi.head_nc > HEAD_EPI_MAX_CLASSES || !epi.head_w ||
         !epi.head_logits)) {
        set_last_error("igemm: the fused head epilogue needs 64 output channels and <= %d classes",
                       HEAD_EPI_MAX_CLASSES);
        return UB_ERR_UNSUPPORTED;
    }
    const int Wo = (src0.W + upper - lower - 1) / tstride + 1;
    const int Ho = (src0.H + upper - lower - 1) / tstride + 1;
    if (Wo <= 0 || Ho <= 0) {
        set_last_error("igemm: empty output (%d x %d)", Ho, Wo);
        return UB_ERR_ARG;
    }
    const long long M = (long long)src0.N * Ho * Wo;
    if (M > 0x7FFFFF00LL) {
        set_last_error("igemm: too many rows");
        return UB_ERR_UNSUPPORTED;
    }
    const int ctot = src0.C + (src1 ? src1->C : 0);
    const long long K = (long long)taps * ctot;

    CUtensorMap mA0, mA1, mB;
    if (rowrun_eligible(Wo, taps, tstride, epi.kind)) {
        RowRunParams q;
        memset(&q, 0, sizeof(q));
        q.N = src0.N; q.Ho = Ho; q.Wo = Wo; q.lower = lower;
        q.cchunks0 = src0.C / 64; q.cchunks1 = src1 ? src1->C / 64 : 0;
        q.qtiles = (Wo + 127) / 128;
        q.n_tiles = ncols / BN;
        // two output rows per tile for BN <= 128 (igemm_rr.cuh: halves the weight traffic through the
        // shared-memory port and stages 4 input rows per 2 output rows); UB_RR_ROWS2=0 disables
        static int rows2_on = -1;
        if (rows2_on < 0) { const char* e = getenv("UB_RR_ROWS2"); rows2_on = (e && !atoi(e)) ? 0 : 1; }
        const int ROWS = (rows2_on && BN <= 128 && Ho >= 2) ? 2 : 1;
        q.HoT = (Ho + ROWS - 1) / ROWS;
        q.m_tiles = src0.N * q.HoT * q.qtiles;
        // (a two-row tile issues twice the MMAs per staged weight tile: it counts as twice as deep for
        // the "pairs need long tiles" rule, which makes d1.a fprop and up4.a dgrad pairs: their
        // single-CTA two-row form must stage weights one tap at a time to fit shared memory)
        int CG = pick_cg(BN, 9 * (q.cchunks0 + q.cchunks1) * ROWS, (long long)q.m_tiles * q.n_tiles);
        // Cin = Cout = 64: weights resident in shared memory (UB_WRES=0 disables). The CTA-pair form of
        // it (each CTA keeps half of the weight rows: 5 KB instead of 6 KB of operand reads per MMA) is
        // built but OFF: measured 46 % slower (inc.b fprop 0.324 -> 0.472 ms, up4.b 0.173 -> 0.249 ms at
        // N = 16) — a 256 x 64 pair MMA issues slower than two 128 x 64 ones (UB_WRES_PAIR=1 enables).
        // The fused-head epilogue exists for CG = 1 only.
        static int wres_on = -1, wres_pair = -1;
        if (wres_on < 0) { const char* e = getenv("UB_WRES"); wres_on = (e && !atoi(e)) ? 0 : 1; }
        if (wres_pair < 0) { const char* e = getenv("UB_WRES_PAIR"); wres_pair = (e && atoi(e)) ? 1 : 0; }
        const bool wres = wres_on && BN == 64 && q.cchunks0 + q.cchunks1 == 1 && q.n_tiles == 1;
        if (epi.kind == EPI_AFFINE_RELU_HEAD) CG = 1;
        else if (wres)
            CG = (wres_pair && pairs_enabled() && (long long)q.m_tiles >= 2LL * num_sms()) ? 2 : 1;
        int rr = make_tmap_rows(&mA0, src0, 130, (unsigned)(ROWS + 2));
        if (!rr && src1) rr = make_tmap_rows(&mA1, *src1, 130, (unsigned)(ROWS + 2));
        if (!src1) mA1 = mA0;
        if (!rr) rr = make_tmap_weights(&mB, wB, (unsigned long long)ctot, (unsigned long long)ncols,
                                        (unsigned long long)taps, (unsigned)(BN / CG),
                                        (unsigned)RowRunCfg<64>::btaps(BN, CG, ROWS));
        if (rr) { set_last_error("row-run: tensor map encoding failed: %d", rr); return UB_ERR_TMAP; }
        q.epi.M = (int)M; q.epi.out = epi.out; q.epi.ldo = epi.ldo; q.epi.bias = epi.bias;
        q.epi.scale = epi.scale; q.epi.shift = epi.shift; q.epi.stats = epi.stats;
        q.epi.red_y = epi.red_y; q.epi.red_mean = epi.red_mean;
        q.epi.head_w = epi.head_w; q.epi.head_b = epi.head_b; q.epi.head_logits = epi.head_logits;
        q.epi.head_mask = epi.head_mask; q.epi.head_nc = epi.head_nc; q.epi.head_hw = Ho * Wo;
        static int fuse_pool = -1;
        if (fuse_pool < 0) { const char* e = getenv("UB_FUSE_EVAL_POOL"); fuse_pool = (e && !atoi(e)) ? 0 : 1; }
        const bool pool_fused = fuse_pool && ROWS == 2 && epi.kind == EPI_AFFINE_RELU && epi.pooled != nullptr;
        q.epi.pooled = pool_fused ? epi.pooled : nullptr;
        q.epi.pool_Ho = Ho; q.epi.pool_Wo = Wo;
        int units = (num_sms() / CG / q.n_tiles) * q.n_tiles;
        if (units <= 0) units = q.n_tiles;
        const long long tiles = (long long)((q.m_tiles + CG - 1) / CG) * q.n_tiles;
        if (units > tiles) units = (int)tiles;
        const int grid = units * CG;
        if (info) {
            info->grid = grid; info->n_tiles = q.n_tiles; info->BN = BN; info->M = (int)M;
            info->pool_fused = pool_fused ? 1 : 0;
        }
        int rc;
#define UB_RR(BN_, CG_, WRES_, ROWS_) launch_rowrun_bn<BN_, CG_, WRES_, ROWS_>(epi.kind, mA0, mA1, mB, q, grid, stream)
        if (BN == 256) rc = CG == 2 ? UB_RR(256, 2, false, 1) : UB_RR(256, 1, false, 1);
        else if (BN == 128) {
            if (ROWS == 2) rc = CG == 2 ? UB_RR(128, 2, false, 2) : UB_RR(128, 1, false, 2);
            else rc = CG == 2 ? UB_RR(128, 2, false, 1) : UB_RR(128, 1, false, 1);
        } else if (wres) {
            if (ROWS == 2) rc = CG == 2 ? UB_RR(64, 2, true, 2) : UB_RR(64, 1, true, 2);
            else rc = CG == 2 ? UB_RR(64, 2, true, 1) : UB_RR(64, 1, true, 1);
        } else {
            if (ROWS == 2) rc = CG == 2 ? UB_RR(64, 2, false, 2) : UB_RR(64, 1, false, 2);
            else rc = CG == 2 ? UB_RR(64, 2, false, 1) : UB_RR(64, 1, false, 1);
        }
#undef UB_RR
        return rc;
    }
    int r = make_tmap_im2col(&mA0, src0, lower, upper, tstride, 128);
    if (r) { set_last_error("igemm: im2col tensor map (source 0) failed: %d", r); return UB_ERR_TMAP; }
    if (src1) {
        r = make_tmap_im2col(&mA1, *src1, lower, upper, tstride, 128);
        if (r) { set_last_error("igemm: im2col tensor map (source 1) failed: %d", r); return UB_ERR_TMAP; }
    } else {
        mA1 = mA0;
    }
    IgemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = (int)M; p.Wo = Wo; p.Ho = Ho; p.lower = lower; p.tstride = tstride;
    p.taps = taps; p.tapw = tapw;
    p.cchunks0 = src0.C / 64; p.cchunks1 = src1 ? src1->C / 64 : 0;
    p.m_tiles = (int)((M + 127) / 128);
    p.n_tiles = ncols / BN;
    // CTA pairs (256 x BN tiles) whenever there are at least two full waves of pair tiles
    const int CG = epi.kind == EPI_AFFINE_RELU_HEAD
                       ? 1 : pick_cg(BN, taps * (p.cchunks0 + p.cchunks1), (long long)p.m_tiles * p.n_tiles);
    r = make_tmap_weights(&mB, wB, (unsigned long long)ctot, (unsigned long long)ncols,
                          (unsigned long long)taps, (unsigned)(BN / CG), 1);
    if (r) { set_last_error("igemm: weight tensor map failed: %d", r); return UB_ERR_TMAP; }
    (void)K;
    p.out = epi.out; p.ldo = epi.ldo; p.bias = epi.bias; p.scale = epi.scale; p.shift = epi.shift;
    p.stats = epi.stats;
    p.red_y = epi.red_y; p.red_mean = epi.red_mean;
    p.head_w = epi.head_w; p.head_b = epi.head_b; p.head_logits = epi.head_logits;
    p.head_mask = epi.head_mask; p.head_nc = epi.head_nc; p.head_hw = Ho * Wo;
    if (epi.kind == EPI_CONVT) {
        p.out = (__nv_bfloat16*)epi.ct_dst.ptr;
        p.ct_cout = ncols / 4; p.ct_H = src0.H; p.ct_W = src0.W;
        p.ct_sN = epi.ct_dst.sN; p.ct_sH = epi.ct_dst.sH; p.ct_sW = epi.ct_dst.sW;
    }
    int units = (num_sms() / CG / p.n_tiles) * p.n_tiles;
    if (units <= 0) units = p.n_tiles;
    const long long tiles = (long long)((p.m_tiles + CG - 1) / CG) * p.n_tiles;
    if (units > tiles) units = (int)tiles;
    const int grid = units * CG;
    if (info) { info->grid = grid; info->n_tiles = p.n_tiles; info->BN = BN; info->M = (int)M; info->pool_fused = 0; }
    if (CG == 2) {
        switch (BN) {
            case 256: return launch_igemm_bn<256, 2>(epi.kind, mA0, mA1, mB, p, grid, stream);
            case 128: return launch_igemm_bn<128, 2>(epi.kind, mA0, mA1, mB, p, grid, stream);
            default: return launch_igemm_bn<64, 2>(epi.kind, mA0, mA1, mB, p, grid, stream);
        }
    }
    switch (BN) {
        case 256: return launch_igemm_bn<256, 1>(epi.kind, mA0, mA1, mB, p, grid, stream);
        case 128: return launch_igemm_bn<128, 1>(epi.kind, mA0, mA1, mB, p, grid, stream);
        default: return launch_igemm_bn<64, 1>(epi.kind, mA0, mA1, mB, p, grid, stream);
    }
}

// ---------------------------------------------------------------------------------------------
static int wgrad_cg(int cols) {
    const int bn = pick_bn(cols);
    return (pairs_enabled() && bn >= 128) ? 2 : 1;
}
static int wgrad_pick_splits(int units, long long kblocks, int cg);
static int wgrad_shift_splits(int ctot, long long mpix, int cg);
static int wgrad_splits(int rows, int cols, long long mpix) {
    const int bn = pick_bn(cols);
    if (!bn) return 1;
    const int cg = wgrad_cg(cols);
    const int m_tiles = (rows + 128 * cg - 1) / (128 * cg), n_tiles = cols / bn;
    const int kpix = bn == 256 ? 64 : 128;
    return wgrad_pick_splits(m_tiles * n_tiles, (mpix + kpix - 1) / kpix, cg);
}
// split count of the shifted form, from the OUTPUT pixel count so that the workspace query (which does
// not know the row width) and the launch agree
// The 128-output-channel pair form is correct but not faster than the tap-major pair form on the reference
// net (profiles/r02_wgrad_shifted.txt: d1.b 0.260 -> 0.272 ms, up3.b 0.119 -> 0.151 ms, up3.a 0.204 -> 0.197 ms;
// the wide U-Net step gains 1 %): opt-in with UB_WGRAD_SHIFT128=1.
static bool wgrad_shift_pair_enabled() {
    static const bool v = [] { const char* e = getenv("UB_WGRAD_SHIFT128"); return e && e[0] == '1'; }();
    return v && pairs_enabled();
}
static int wgrad_shift_splits(int ctot, long long mpix, int cg) {
    return wgrad_pick_splits((3 * ctot + 128 * cg - 1) / (128 * cg), (mpix + 63) / 64, cg);
}
static int wgrad_pick_splits(int units, long long kblocks, int cg) {
    const long long smax = kblocks / 8 > 1 ? kblocks / 8 : 1;
    // UB_WGRAD_MAXKB=n: at most n k-blocks per CTA (short-lived CTAs release their SM sooner to the
    // data-gradient kernel they overlap with); 0 = only the wave rule below
    static const int maxkb = [] { const char* e = getenv("UB_WGRAD_MAXKB"); return e ? atoi(e) : 0; }();
    const long long smin = maxkb > 0 ? (kblocks + maxkb - 1) / maxkb : 1;
    // Pick the split count that minimises (number of CTA waves) x (k-blocks per CTA + fixed cost):
    // avoids e.g. 300 CTAs on 148 SMs (a third, nearly empty, wave).
    const double fixed = 24.0;  // prologue + fp32 epilogue of one CTA, in k-block units
    const int slots = num_sms() / cg;   // CTAs (or CTA pairs) resident at once
    long long best = 1;
    double best_cost = 1e30;
    for (long long s = smin < smax ? smin : smax; s <= smax && s <= 1024; ++s) {
        const long long ctas = (long long)units * s;
        const long long waves = (ctas + slots - 1) / slots;
        const double cost = (double)waves * ((double)((kblocks + s - 1) / s) + fixed);
        if (cost < best_cost * 0.999) { best_cost = cost; best = s; }
    }
    return (int)best;
}
size_t wgrad_ws_floats(int rows, int cols, long long mpix) {
    // sized for either tiling (UB_PAIR may differ between the size query and the launch only in tests)
    size_t splits = (size_t)wgrad_splits(rows, cols, mpix);
    if ((cols == 64 || cols == 128) && rows % (9 * 64) == 0) {   // 3x3 layer eligible for a shifted form
        const size_t s2 = (size_t)wgrad_shift_splits(rows / 9, mpix, cols == 128 ? 2 : 1);
        if (s2 > splits) splits = s2;
    }
    return splits * (size_t)rows * (size_t)cols;
}

// "Shifted dY" form for 3x3 layers with 64 output channels (UB_WGRAD_SHIFT=0 disables): GEMM rows =
// (filter row dy, input channel), GEMM columns = (filter column dx, output channel) = 192, reduction
// over the INPUT-wide pixel grid; dW[dy][dx] = sum_u X[u + (dy, 0)] dY[u - (0, dx)].
static bool wgrad_shift_enabled() {
    static const bool v = [] { const char* e = getenv("UB_WGRAD_SHIFT"); return !(e && e[0] == '0'); }();
    return v;
}

template <int BN, int CG>
static int launch_wgrad_t(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                          const WgradParams& p, dim3 grid, cudaStream_t stream) {
    using Cfg = WgradCfg<BN, CG>;
    static bool attr_set = false;
    if (!attr_set) {
        UB_CHECK_CUDA(cudaFuncSetAttribute(igemm_wgrad_kernel<BN, CG>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Cfg::SMEM_BYTES));
        attr_set = true;
    }
    UB_CHECK_CUDA(ub_launch(igemm_wgrad_kernel<BN, CG>, grid, dim3(256), Cfg::SMEM_BYTES,
                                   stream, CG, a0, a1, b, p));
    UB_POST_LAUNCH();
    return UB_OK;
}

int launch_wgrad(const View& src0, const View* src1, int lower, int upper, int tstride, int taps,
                 int tapw, const __nv_bfloat16* B, long long ldb, int cols, float* ws,
                 size_t ws_floats, float* out, cudaStream_t stream, float* zero0, int nzero0,
                 float* zero1, int nzero1) {
    UB_TRY(check_view(src0, "wgrad source 0"));
    if (src1) UB_TRY(check_view(*src1, "wgrad source 1"));
    const int BN = pick_bn(cols);
    if (!BN) {
        set_last_error("wgrad: %d columns is not a multiple of 64", cols);
        return UB_ERR_UNSUPPORTED;
    }
    const int Wo = (src0.W + upper - lower - 1) / tstride + 1;
    const int Ho = (src0.H + upper - lower - 1) / tstride + 1;
    const long long mpix = (long long)src0.N * Ho * Wo;
    const int ctot = src0.C + (src1 ? src1->C : 0);
    const int rows = taps * ctot;
    if (taps == 9 && tapw == 3 && tstride == 1 && lower == 0 && upper == -2 && wgrad_shift_enabled() &&
        (cols == 64 || (cols == 128 && wgrad_shift_pair_enabled()))) {
        // ---- shifted forms: a CTA computes a 128 x 192 tile (a pair 256 x 384), A staged once for
        //      three filter columns ----
        constexpr int KP = WgradCfg<192, 1>::KPIX;
        static_assert(WgradCfg<384, 2>::KPIX == KP, "one k-block size for both shifted forms");
        const int scg = cols == 128 ? 2 : 1;
        const long long tpix = (long long)src0.N * Ho * src0.W;          // base pixels: input-wide rows
        const int srows = 3 * ctot, m_tiles = (srows + 128 * scg - 1) / (128 * scg);
        const long long kblocks = (tpix + KP - 1) / KP;
        const int splits = wgrad_shift_splits(ctot, mpix, scg);
        if ((size_t)splits * rows * cols <= ws_floats && tpix < 0x7FFFFFFFLL) {
            CUtensorMap mA0, mA1, mB;
            int r = make_tmap_im2col_wh(&mA0, src0, 0, 0, 0, -2, 1, (unsigned)KP);
            if (!r && src1) r = make_tmap_im2col_wh(&mA1, *src1, 0, 0, 0, -2, 1, (unsigned)KP);
            if (!src1) mA1 = mA0;
            View dyv{};
            dyv.ptr = B;
            dyv.N = src0.N; dyv.H = Ho; dyv.W = Wo; dyv.C = cols;
            dyv.sW = ldb; dyv.sH = (long long)Wo * ldb; dyv.sN = (long long)Ho * Wo * ldb;
            if (!r) r = make_tmap_im2col_wh(&mB, dyv, -2, 0, 0, 0, 1, (unsigned)KP);
            if (r) { set_last_error("wgrad: im2col tensor map (shifted form) failed: %d", r); return UB_ERR_TMAP; }
            WgradParams p;
            memset(&p, 0, sizeof(p));
            p.Mpix = (int)tpix; p.Wo = src0.W; p.Ho = Ho; p.lower = 0; p.tstride = 1;
            p.taps = 3; p.tapw = 1;   // A-side taps: filter rows only
            p.cchunks0 = src0.C / 64; p.cchunks1 = src1 ? src1->C / 64 : 0;
            p.a_chunks_total = srows / 64;
            p.n_tiles = 1; p.splits = splits;
            p.kblocks_total = (int)kblocks;
            p.ws = ws; p.ldw = 3 * cols; p.split_stride = (long long)rows * cols;
            if (scg == 2) UB_TRY((launch_wgrad_t<384, 2>(mA0, mA1, mB, p, dim3(m_tiles * 2, splits), stream)));
            else UB_TRY((launch_wgrad_t<192, 1>(mA0, mA1, mB, p, dim3(m_tiles, splits), stream)));
            const dim3 rgrid(cols / 32, ctot / 8);
            UB_LAUNCH_NC((wgrad_reduce_kernel<9>), rgrid, dim3(32, 8, WGR_SLICES), 0, stream, ws, splits, p.split_stride, cols, ctot, out, zero0, nzero0, zero1, nzero1, 3);
            UB_POST_LAUNCH();
            return UB_OK;
        }
        // workspace sized by an older query: fall through to the tap-major form
    }
    const int CG = wgrad_cg(cols);
    const int splits = wgrad_splits(rows, cols, mpix);
    if ((size_t)splits * rows * cols > ws_floats) {
        set_last_error("wgrad: workspace too small");
        return UB_ERR_ARG;
    }
    CUtensorMap mA0, mA1, mB;
    const int kpix = BN == 256 ? 64 : 128;
    int r = make_tmap_im2col(&mA0, src0, lower, upper, tstride, (unsigned)kpix);
    if (r) { set_last_error("wgrad: im2col tensor map failed: %d", r); return UB_ERR_TMAP; }
    if (src1) {
        r = make_tmap_im2col(&mA1, *src1, lower, upper, tstride, (unsigned)kpix);
        if (r) { set_last_error("wgrad: im2col tensor map (source 1) failed: %d", r); return UB_ERR_TMAP; }
    } else {
        mA1 = mA0;
    }
    r = make_tmap_chunked(&mB, B, (unsigned long long)cols, (unsigned long long)mpix,
                          (unsigned long long)ldb * 2, (unsigned)kpix, (unsigned)(BN / 64 / CG));
    if (r) { set_last_error("wgrad: matrix tensor map failed: %d", r); return UB_ERR_TMAP; }

    WgradParams p;
    memset(&p, 0, sizeof(p));
    p.Mpix = (int)mpix; p.Wo = Wo; p.Ho = Ho; p.lower = lower; p.tstride = tstride;
    p.taps = taps; p.tapw = tapw;
    p.cchunks0 = src0.C / 64; p.cchunks1 = src1 ? src1->C / 64 : 0;
    p.a_chunks_total = rows / 64;
    p.n_tiles = cols / BN; p.splits = splits;
    p.kblocks_total = (int)((mpix + kpix - 1) / kpix);
    p.ws = ws; p.ldw = cols; p.split_stride = (long long)rows * cols;
    const int m_tiles = (rows + 128 * CG - 1) / (128 * CG);
    dim3 grid(m_tiles * p.n_tiles * CG, splits);
    int rc;
    if (CG == 2) {
        if (BN == 256) rc = launch_wgrad_t<256, 2>(mA0, mA1, mB, p, grid, stream);
        else rc = launch_wgrad_t<128, 2>(mA0, mA1, mB, p, grid, stream);
    } else {
        switch (BN) {
            case 256: rc = launch_wgrad_t<256, 1>(mA0, mA1, mB, p, grid, stream); break;
            case 128: rc = launch_wgrad_t<128, 1>(mA0, mA1, mB, p, grid, stream); break;
            default: rc = launch_wgrad_t<64, 1>(mA0, mA1, mB, p, grid, stream); break;
        }
    }
    UB_TRY(rc);
    const dim3 rgrid(cols / 32, ctot / 8);
    if (taps == 9)
        UB_LAUNCH_NC((wgrad_reduce_kernel<9>), rgrid, dim3(32, 8, WGR_SLICES), 0, stream, ws, splits, p.split_stride, cols, ctot, out, zero0, nzero0, zero1, nzero1, 1);
    else if (taps == 4)
        UB_LAUNCH_NC((wgrad_reduce_kernel<4>), rgrid, dim3(32, 8, WGR_SLICES), 0, stream, ws, splits, p.split_stride, cols, ctot, out, zero0, nzero0, zero1, nzero1, 1);
    else {
        set_last_error("wgrad: unsupported tap count %d", taps);
        return UB_ERR_UNSUPPORTED;
    }
    UB_POST_LAUNCH();
    return UB_OK;
}

}  // namespace ub
