// "Row-run" implicit-GEMM 3x3 convolution (forward and data gradient) for wide feature maps.
//
// The im2col TMA path (igemm.cuh) re-reads every input pixel nine times, once per filter tap, and is
// bound by the ~55 B/cycle an SM can ingest from L2. Here a tile is 128 consecutive output pixels of
// ONE output row: for each 64-channel chunk the three input rows it needs (130 pixels each) are
// loaded ONCE with tiled-mode TMA into 128-byte-swizzled shared memory, and the nine taps are
// addressed as shifted windows of those rows: the UMMA shared-memory descriptor of tap (r, s) starts
// at row-buffer r + s*128 bytes. (Measured on B200: the 128B swizzle is a function of the absolute
// shared-memory address, so a descriptor start shifted by whole 128-byte rows, with base_offset 0,
// reads exactly the rows TMA wrote — scripts_dev/shift_probe.py.) A-operand ingest drops 2.9x.
//
// Pipelines: A ring (SA stages of 3 row buffers) fed by warp 0, B ring (SB stages of one
// (tap, chunk) weight tile) fed by warp 6, MMA issue by warp 1, epilogue warps 2-5 (and 7-10) on a
// double-buffered TMEM accumulator — same epilogues as igemm_kmajor_kernel.
#pragma once
#include "igemm.cuh"

namespace ub {

struct RowRunParams {
    int N, Ho, Wo;        // output extents; the tile (n, p, qt) covers q in [128*qt, 128*qt + 128)
    int lower;            // input coordinate of tap (0,0) relative to the output pixel (0 / -2)
    int cchunks0, cchunks1;
    int qtiles;           // ceil(Wo / 128)
    int HoT;              // ceil(Ho / ROWS): a tile covers ROWS consecutive output rows
    int m_tiles, n_tiles; // m_tiles = N * HoT * qtiles
    IgemmParams epi;      // epilogue fields (out, ldo, bias, scale, shift, stats); M = N*Ho*Wo
};

// One TMA operation costs the issuing thread a few hundred cycles whatever its size, and one
// barrier wait + tcgen05 fence costs the MMA thread ~230 cycles, so operations are made as large as
// possible: ONE box (64 ch, 130 px, 3 rows) per A stage, and BTAPS filter taps of weights per B stage
// (a whole filter row for BN <= 128), i.e. 12 MMAs per B wait and 36 per A wait.
// WRES (BN = 64, one 64-channel chunk, i.e. Cin = Cout = 64): the nine weight taps (72 KB) stay
// resident in shared memory for the whole kernel instead of being re-staged for every tile — these
// tiles are bound by shared-memory bandwidth (TMA writes + MMA operand reads), and the weights were
// 21 % of it.
// ROWS = 2 (BN <= 128): a tile is TWO consecutive output rows of the same 128-pixel column range. One
// A stage then holds four input rows (66.5 KB for two output rows instead of 2 x 50 KB) and every
// staged weight tile feeds the MMAs of both rows, so the weights cross the shared-memory port once
// per two output rows; the two rows accumulate side by side in tensor memory (2 x 2 x BN columns with
// the double buffering). These layers are bound by that port (TMA writes + MMA operand reads share
// 128 B/clk): per output row d1.a fprop moves 393 KB instead of 482 KB, up4.a 462 instead of 532.
template <int BN, int CG = 1, bool WRES = false, int ROWS = 1>
struct RowRunCfg {
    static constexpr int btaps(int bn, int cg, int rows = 1) {
        return bn <= 128 ? ((rows == 2 && bn == 128 && cg == 1) ? 1 : 3) : 1;
    }
    static constexpr int BTAPS = btaps(BN, CG, ROWS);
    static constexpr int ROW_BYTES = 130 * 128;            // input rows are packed back to back
    static constexpr int A_ROWS = ROWS + 2;
    static constexpr int A_TX = A_ROWS * ROW_BYTES;        // bytes one A stage receives (49920 / 66560)
    static constexpr int A_STAGE = (A_TX + 1023) / 1024 * 1024;   // 1024-aligned stage pitch
    static constexpr int B_ROWS = BN / CG;                 // weight rows staged by one CTA
    static constexpr int B_TILE = B_ROWS * 128;            // one tap: [B_ROWS][64] K-major
    static constexpr int B_STAGE = BTAPS * B_TILE;
    // Ring depths. A B stage of a BN = 128 pair is only 12 MMAs x 64 cycles = 768 cycles of cover
    // against ~1 us of TMA latency: with two stages the MMA warp waited for weights a third of the
    // time (role profile), so that ring gets four stages and the A ring, which had slack, two.
    static constexpr int SA = ROWS == 2 ? 2 : ((CG == 2) ? (BN == 64 ? 3 : 2) : ((BN == 64) ? 3 : 2));
    static constexpr int SB = WRES ? 1
                              : ROWS == 2 ? ((BN == 128 && CG == 2) || (BN == 64 && CG == 1) ? 3 : 4)
                                          : ((CG == 2) ? (BN == 256 ? 6 : 4) : ((BN == 64) ? 3 : (BN == 128 ? 2 : 3)));
    static constexpr int B_AREA = WRES ? 9 * B_TILE : SB * B_STAGE;
    static constexpr int BAR_BYTES = 256;
    // epilogue scratch: BN-statistics rows of the 4 lane quadrants, or the fused head's weights
    static constexpr int STAT_BYTES = (BN == 64) ? 2304 : 4 * 2 * BN * 4;
    static constexpr int CONST_BYTES = 3 * BN * 4;   // per-column epilogue constants of the n tile
    static constexpr int SMEM_BYTES = SA * A_STAGE + B_AREA + BAR_BYTES + STAT_BYTES + CONST_BYTES + 1024;
    static_assert(!WRES || BN == 64, "resident weights: BN = 64 (one 64-channel chunk)");
    static_assert(ROWS == 1 || (ROWS == 2 && BN <= 128), "two-row tiles: 2 x 2 x BN TMEM columns");
    static constexpr uint32_t TMEM_COLS = 2 * ROWS * BN;
    static_assert(SMEM_BYTES <= 227 * 1024, "row-run configuration exceeds shared memory");
};

template <int BN, int EPI, int CG, bool WRES, int ROWS>
__global__ void __launch_bounds__(igemm_threads(BN), 1)
igemm_rowrun_kernel(const __grid_constant__ CUtensorMap mapA0,
                    const __grid_constant__ CUtensorMap mapA1,
                    const __grid_constant__ CUtensorMap mapB, const RowRunParams p) {
    pdl_trigger();   // let the next kernel's CTAs be scheduled while this grid drains
    using Cfg = RowRunCfg<BN, CG, WRES, ROWS>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - raw);
    const uint32_t b_base = base + Cfg::SA * Cfg::A_STAGE;
    const uint32_t bar_base = b_base + Cfg::B_AREA;
    auto fullA = [&](int s) { return bar_base + 8u * s; };
    auto emptyA = [&](int s) { return bar_base + 8u * (Cfg::SA + s); };
    auto fullB = [&](int s) { return bar_base + 8u * (2 * Cfg::SA + s); };
    auto emptyB = [&](int s) { return bar_base + 8u * (2 * Cfg::SA + Cfg::SB + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::SA + 2 * Cfg::SB + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::SA + 2 * Cfg::SB + 2 + s); };
    constexpr int kSlot = 2 * Cfg::SA + 2 * Cfg::SB + 4;
    const uint32_t tmem_slot = bar_base + 8u * kSlot;
    volatile uint32_t* tmem_slot_g = reinterpret_cast<volatile uint32_t*>(
        gbase + Cfg::SA * Cfg::A_STAGE + Cfg::B_AREA + 8 * kSlot);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // CTA pair geometry (see common.cuh): unit = CTA or CTA pair, rank 0 issues the MMAs
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const int unit = blockIdx.x / CG;
    const int nunits = gridDim.x / CG;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA0);
        tma_prefetch_desc(&mapA1);
        tma_prefetch_desc(&mapB);
        for (int s = 0; s < Cfg::SA; ++s) { mbar_init(fullA(s), 1); mbar_init(emptyA(s), 1); }
        for (int s = 0; s < Cfg::SB; ++s) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4 * EpiCfg<BN>::HALVES * CG); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_cg<Cfg::TMEM_COLS, CG>(tmem_slot);
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_g;
    pdl_wait();   // prologue done; from here on global memory of the preceding kernels is read
    float* hs = reinterpret_cast<float*>(gbase + Cfg::SA * Cfg::A_STAGE + Cfg::B_AREA + Cfg::BAR_BYTES);
    float* cs = reinterpret_cast<float*>(gbase + Cfg::SA * Cfg::A_STAGE + Cfg::B_AREA + Cfg::BAR_BYTES +
                                         Cfg::STAT_BYTES);
    stage_head_weights<EPI>(p.epi, hs);
    stage_epilogue_consts<BN, EPI>(p.epi, (unit % p.n_tiles) * BN, cs);

    const int cchunks = p.cchunks0 + p.cchunks1;
    const int cta_n = unit % p.n_tiles;
    const int m_first = unit / p.n_tiles;
    const int m_step = nunits / p.n_tiles;
    const int m_units = (p.m_tiles + CG - 1) / CG;
    const int n0 = cta_n * BN;

    // (whole warps run the role loops; see the note on elect_one() in igemm_kmajor_kernel)
    if (warp == 0) {
        // ---------------- A producer: ROWS + 2 input rows x 130 pixels x 64 channels per chunk ------
        int stage = 0;
        uint32_t phase = 0;
        for (int mu = m_first; mu < m_units; mu += m_step) {
            int mt = mu * CG + (int)rank;
            if (mt >= p.m_tiles) mt = p.m_tiles - 1;   // odd tail: reload a valid tile, rows masked
            const int qt = mt % p.qtiles;
            const int t = mt / p.qtiles;
            const int pr = (t % p.HoT) * ROWS;
            const int n = t / p.HoT;
            const int w0 = qt * 128 + p.lower;
            const int h0 = pr + p.lower;   // rows past the tensor (odd Ho, last tile) are zero filled
            for (int cc = 0; cc < cchunks; ++cc) {
                mbar_wait(emptyA(stage), phase ^ 1u);
                const uint32_t sa = base + stage * Cfg::A_STAGE;
                const uint32_t fb = (CG == 2) ? mapa_rank(fullA(stage), 0) : fullA(stage);
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(fullA(stage), CG * Cfg::A_TX);
                    if (cc < p.cchunks0)   // box (64, 130, ROWS + 2, 1)
                        tma_load_4d_cg<CG>(sa, &mapA0, fb, cc * 64, w0, h0, n);
                    else
                        tma_load_4d_cg<CG>(sa, &mapA1, fb, (cc - p.cchunks0) * 64, w0, h0, n);
                }
                __syncwarp();
                if (++stage == Cfg::SA) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 6) {
        // ---------------- B producer: BTAPS (tap, chunk) weight tiles per stage --------------------
        int stage = 0;
        uint32_t phase = 0;
        const int nrow0 = n0 + (int)rank * Cfg::B_ROWS;
        if (WRES) {   // all nine taps once: three boxes of (64 ch, B_ROWS rows, 3 taps); in a pair each
                      // CTA keeps its half of the weight rows and both complete on the leader's barrier
            const uint32_t fb = (CG == 2) ? mapa_rank(fullB(0), 0) : fullB(0);
            if (elect_one()) {
                if (rank == 0) mbar_expect_tx(fullB(0), CG * 9 * Cfg::B_TILE);
#pragma unroll
                for (int tap = 0; tap < 9; tap += 3)
                    tma_load_3d_cg<CG>(b_base + tap * Cfg::B_TILE, &mapB, fb, 0, nrow0, tap);
            }
            __syncwarp();
        } else
        for (int mu = m_first; mu < m_units; mu += m_step) {
            for (int cc = 0; cc < cchunks; ++cc) {
#pragma unroll 1
                for (int tap = 0; tap < 9; tap += Cfg::BTAPS) {
                    mbar_wait(emptyB(stage), phase ^ 1u);
                    const uint32_t fb = (CG == 2) ? mapa_rank(fullB(stage), 0) : fullB(stage);
                    if (elect_one()) {
                        if (rank == 0) mbar_expect_tx(fullB(stage), CG * Cfg::B_STAGE);
                        tma_load_3d_cg<CG>(b_base + stage * Cfg::B_STAGE, &mapB, fb, cc * 64, nrow0,
                                           tap);
                    }
                    __syncwarp();
                    if (++stage == Cfg::SB) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ---------------- MMA issuer (leader CTA) ---------------------------------------------------
        constexpr uint32_t idesc = make_idesc_bf16(128 * CG, BN, 0, 0);
        int sa_i = 0, sb_i = 0;
        uint32_t pa = 0, pb = 0;
        int as = 0;
        uint32_t aphase = 0;
        if (WRES) mbar_wait(fullB(0), 0);
        for (int mu = m_first; mu < m_units; mu += m_step) {
            mbar_wait(tempty_bar(as), aphase ^ 1u);
            tc_fence_after();
            // accumulator of output row j of the tile: columns [(as * ROWS + j) * BN, + BN)
            const uint32_t tmem_d = tmem_base + (uint32_t)(as * ROWS * BN);
            if (WRES) {
                // one chunk, resident weights: all 36 * ROWS MMAs of the tile behind a single wait
                mbar_wait(fullA(sa_i), pa);
                tc_fence_after();
                const uint32_t sa = base + sa_i * Cfg::A_STAGE;
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < ROWS; ++j) {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const uint64_t da = make_smem_desc(
                                sa + (tap / 3 + j) * Cfg::ROW_BYTES + (tap % 3) * 128, 0, 1024);
                            const uint64_t db = make_smem_desc(b_base + tap * Cfg::B_TILE, 0, 1024);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16_cg<CG>(tmem_d + (uint32_t)(j * BN), da + (uint64_t)(2 * k),
                                                 db + (uint64_t)(2 * k), idesc, (uint32_t)((tap | k) != 0));
                        }
                    }
                    umma_commit_cg<CG>(emptyA(sa_i));
                    umma_commit_cg<CG>(tfull_bar(as));
                }
                __syncwarp();
                if (++sa_i == Cfg::SA) { sa_i = 0; pa ^= 1u; }
            } else
            for (int cc = 0; cc < cchunks; ++cc) {
                mbar_wait(fullA(sa_i), pa);
                const uint32_t sa = base + sa_i * Cfg::A_STAGE;
#pragma unroll 1
                for (int tap = 0; tap < 9; tap += Cfg::BTAPS) {
                    mbar_wait(fullB(sb_i), pb);
                    tc_fence_after();
                    const uint32_t first = (cc == 0 && tap == 0) ? 0u : 1u;   // 0: overwrite the accumulator
                    if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < ROWS; ++j) {
#pragma unroll
                    for (int jt = 0; jt < Cfg::BTAPS; ++jt) {
                        const int r = (tap + jt) / 3, sft = (tap + jt) % 3;
                        // tap (r, sft) of output row j = window of input row r + j shifted by sft pixels
                        const uint64_t da =
                            make_smem_desc(sa + (r + j) * Cfg::ROW_BYTES + sft * 128, 0, 1024);
                        const uint64_t db =
                            make_smem_desc(b_base + sb_i * Cfg::B_STAGE + jt * Cfg::B_TILE, 0, 1024);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16_cg<CG>(tmem_d + (uint32_t)(j * BN), da + (uint64_t)(2 * k),
                                             db + (uint64_t)(2 * k), idesc, first | (uint32_t)((jt | k) != 0));
                    }
                    }
                    umma_commit_cg<CG>(emptyB(sb_i));
                    if (tap + Cfg::BTAPS >= 9) umma_commit_cg<CG>(emptyA(sa_i));
                    if (tap + Cfg::BTAPS >= 9 && cc == cchunks - 1) umma_commit_cg<CG>(tfull_bar(as));
                    }
                    __syncwarp();
                    if (++sb_i == Cfg::SB) { sb_i = 0; pb ^= 1u; }
                }
                if (++sa_i == Cfg::SA) { sa_i = 0; pa ^= 1u; }
            }
            if (++as == 2) { as = 0; aphase ^= 1u; }
        }
    } else if (is_epilogue_warp<BN>(warp)) {
        // ---------------- epilogue (lane quadrant x column half) --------------------------------------
        const int quad = warp & 3;
        const int chalf = warp > 6 ? 1 : 0;
        const int row_in_tile = quad * 32 + lane;
        int as = 0;
        uint32_t aphase = 0;
        constexpr int NCH = EpiCfg<BN>::NCH;
        float ssum[NCH], ssq[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) { ssum[c] = 0.f; ssq[c] = 0.f; }
        StatRegs<BN, EPI> sr;
        sr.clear();
        for (int mu = m_first; mu < m_units; mu += m_step) {
            const int mt = mu * CG + (int)rank;
            const int qt = mt % p.qtiles;
            const int t = mt / p.qtiles;      // = n * HoT + (row pair)
            const int pr = (t % p.HoT) * ROWS;
            const int n = t / p.HoT;
            const int q = qt * 128 + row_in_tile;
            uint4 ypre[EPI == EPI_STORE_BNRED ? ROWS : 1][EPI == EPI_STORE_BNRED ? NCH * 4 : 1];
            if constexpr (EPI == EPI_STORE_BNRED) {
#pragma unroll
                for (int j = 0; j < ROWS; ++j)
                    bnred_prefetch<BN, EPI>(p.epi, ((long long)n * p.Ho + pr + j) * p.Wo + q,
                                            q < p.Wo && mt < p.m_tiles && pr + j < p.Ho, n0, chalf, ypre[j]);
            }
            mbar_wait(tfull_bar(as), aphase);
            tc_fence_after();
            if (ROWS == 2 && EPI == EPI_AFFINE_RELU && p.epi.pooled != nullptr) {
                const bool v0 = q < p.Wo && mt < p.m_tiles && pr < p.Ho;
                const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * ROWS * BN);
                epilogue_tile_pool2<BN>(p.epi, t0, t0 + (uint32_t)BN, ((long long)n * p.Ho + pr) * p.Wo + q,
                                        n, pr, q, v0, v0 && pr + 1 < p.Ho, n0, lane, chalf, cs);
            } else
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                const bool valid = q < p.Wo && mt < p.m_tiles && pr + j < p.Ho;
                const long long m = ((long long)n * p.Ho + pr + j) * p.Wo + q;
                const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) +
                                      (uint32_t)((as * ROWS + j) * BN);
                epilogue_tile<BN, EPI>(p.epi, trow, m, valid, n0, lane, chalf, ssum, ssq, sr, cs, hs,
                                       ypre[EPI == EPI_STORE_BNRED ? j : 0]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 2) mbar_arrive_cluster(mapa_rank(tempty_bar(as), 0));
                else mbar_arrive(tempty_bar(as));
            }
            if (++as == 2) { as = 0; aphase ^= 1u; }
        }
        finish_stat_regs<BN, EPI>(sr, lane, ssum, ssq);
        if (EpiTraits<EPI>::SUMS) {
            float* red = reinterpret_cast<float*>(gbase + Cfg::SA * Cfg::A_STAGE + Cfg::B_AREA +
                                                  Cfg::BAR_BYTES);
            write_cta_stats<BN>(red, p.epi.stats + (long long)((int)rank * nunits + unit) * (2 * BN),
                                warp, lane, quad, chalf, ssum, ssq);
        }
    }

    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_cg<Cfg::TMEM_COLS, CG>(tmem_base);
    }
}

}  // namespace ub
