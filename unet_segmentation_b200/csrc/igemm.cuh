// Implicit-GEMM convolution kernels on tcgen05 tensor cores (sm_100a).
//
//  * igemm_kmajor_kernel  — forward 3x3 valid conv, its data gradient, the 2x2/stride-2
//    transposed conv and its data gradient. A operand = im2col TMA tiles of an NHWC bf16 view
//    (up to two sources: zero-copy channel concat), B operand = packed bf16 weights, both K-major,
//    fp32 accumulators in TMEM, persistent CTAs, warp-specialised (TMA / MMA / 4 epilogue warps).
//  * igemm_wgrad_kernel   — weight gradients: reduction over pixels, both operands MN-major,
//    split-K over CTAs with fp32 partial tiles reduced in a fixed order afterwards.
//
// Reference ops replaced: nn.Conv2d(k=3,p=0) fwd/bwd (models/unet_model.py:11,15),
// nn.ConvTranspose2d(k=2,s=2) fwd/bwd (:45), torch.cat of the cropped skip (:131-143).
#pragma once
#include "common.cuh"

namespace ub {

// EPI_AFFINE_RELU_HEAD (eval, last conv unit, Cout = 64): folded BN + ReLU followed by the 1x1 output
// convolution and the z1 > z0 mask in the SAME epilogue — every epilogue thread holds one pixel's 64
// activations, so the logits are 64 FMAs per class and the activation tensor is never written
// (reference models/unet_model.py:56-63,145; scripts/predict.py:85-92).
// EPI_STORE_BNRED (data gradients): EPI_STORE plus the REDUCE pass of the BatchNorm + ReLU backward of the
// layer that receives this gradient (reference nn.BatchNorm2d / nn.ReLU backward, models/unet_model.py:
// 12-13,16-17): the tile just computed IS the upstream gradient g of that layer, so the epilogue fetches
// the layer's stored pre-BN output y at the same pixels, forms dyh = g * [y*scale + shift > 0] and
// accumulates the per-channel sums  S1 = sum dyh,  S2 = sum dyh * (y - mean)  with the machinery of the
// forward statistics (one deterministic partial row per CTA). The stand-alone reduce kernel — a full HBM
// pass over (y, g), 4 B/element — is then skipped. g enters the sums as rounded to bf16, i.e. exactly
// the value the apply pass reads back.
enum : int { EPI_CONV_STATS = 0, EPI_STORE = 1, EPI_AFFINE_RELU = 2, EPI_CONVT = 3,
             EPI_AFFINE_RELU_HEAD = 4, EPI_STORE_BNRED = 5 };
template <int EPI> struct EpiTraits {
    static constexpr bool SUMS = (EPI == EPI_CONV_STATS || EPI == EPI_STORE_BNRED);   // per-column sums
};
constexpr int HEAD_EPI_MAX_CLASSES = 8;

struct IgemmParams {
    int M;               // GEMM rows = base pixels = N*Ho*Wo
    int Wo, Ho;          // traversal extents of the im2col bounding box
    int lower;           // lower corner of the bounding box (same for w and h)
    int tstride;         // traversal stride
    int taps, tapw;      // filter taps and taps per filter row
    int cchunks0, cchunks1;  // 64-channel chunks per tap taken from source 0 / source 1
    int m_tiles, n_tiles;
    __nv_bfloat16* out;  // [M][ldo] (EPI_CONV_STATS / STORE / AFFINE_RELU)
    long long ldo;
    const float* bias;   // per GEMM column, may be null
    const float* scale;  // EPI_AFFINE_RELU: y = relu(acc*scale + shift)
    const float* shift;
    float* stats;        // EPI_CONV_STATS: [gridDim.x][2][BN] per-CTA partial (sum, sumsq); BNRED: (S1, S2)
    const __nv_bfloat16* red_y;   // EPI_STORE_BNRED: pre-BN output of the receiving layer, [M][ldo]
    const float* red_mean;        // EPI_STORE_BNRED: its batch mean per column (scale / shift above)
    // EPI_AFFINE_RELU_HEAD: head weights [nc][64] + bias [nc] (fp32), logits NCHW fp32, u8 mask
    const float* head_w;
    const float* head_b;
    float* head_logits;
    unsigned char* head_mask;   // may be null
    int head_nc, head_hw;       // classes, pixels per image (Ho*Wo)
    // EPI_AFFINE_RELU in the two-row row-run kernel: optional fused 2x2 floor max-pool (eval path),
    // pooled [N][Ho/2][Wo/2][ldo] written next to the activation
    __nv_bfloat16* pooled;
    int pool_Ho, pool_Wo;       // extents of the conv output (= un-pooled activation)
    // EPI_CONVT: GEMM column = q*ct_cout + co, q = dy*2+dx; row = (n,h,w) of the input
    int ct_cout, ct_H, ct_W;
    long long ct_sN, ct_sH, ct_sW;  // element strides of the destination [N,2H,2W,*] view
};

// CG = CTAs per MMA (1, or 2 = cta_group::2 pair computing a 256 x BN tile; each CTA stages its own
// 128 rows of A and BN/2 rows of B).
template <int BN, int CG = 1>
struct IgemmCfg {
    static constexpr int A_BYTES = 128 * 128;
    static constexpr int B_ROWS = BN / CG;
    static constexpr int B_BYTES = B_ROWS * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (CG == 2) ? (BN == 256 ? 6 : 8)
                                            : ((BN == 256) ? 4 : (BN == 128 ? 6 : 8));
    static constexpr int BAR_BYTES = 256;
    // epilogue scratch: BN-statistics rows of the 4 lane quadrants, or the fused head's weights
    static constexpr int STAT_BYTES = (BN == 64) ? 2304 : 4 * 2 * BN * 4;
    static constexpr int CONST_BYTES = 3 * BN * 4;   // per-column epilogue constants of the n tile
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + STAT_BYTES + CONST_BYTES + 1024;
    static constexpr uint32_t TMEM_COLS = 2 * BN;
};

// Warp roles (352 threads): 0 = A producer, 1 = MMA issuer + TMEM allocator, 6 = B producer,
// 2-5 and 7-10 = epilogue (lane quadrant = warp & 3; warps 7-10 take the upper half of the columns).
// BN = 64 keeps four epilogue warps (one 32-column chunk pair each): its tiles are MMA-issue bound
// and a second epilogue warp on the MMA warp's scheduler costs more than it saves (measured).
constexpr int IGEMM_THREADS = 352;
constexpr int igemm_threads(int bn) { return bn == 64 ? 224 : IGEMM_THREADS; }   // BN = 64: warps 0-6
template <int BN> struct EpiCfg {
    static constexpr int HALVES = (BN == 64) ? 1 : 2;      // column halves = epilogue warps / 4
    static constexpr int NCH = BN / 32 / HALVES;           // 32-column chunks per warp
};
template <int BN>
__device__ __forceinline__ bool is_epilogue_warp(int warp) {
    return warp >= 2 && warp != 6 && (EpiCfg<BN>::HALVES == 2 || warp < 6);
}

// BN = 64 statistics epilogue: per-thread register accumulators (row of the thread x 64 columns, sum
// and sum of squares) across ALL tiles of the CTA, reduced across lanes once at the end — 64
// instructions per tile and chunk instead of the ~250 of a per-tile warp transpose-reduction, which
// paced the resident-weight kernel (its MMA warp waited 23 % of the time for a free accumulator).
// The 224-thread launch of BN = 64 leaves the registers for it.
template <int BN, int EPI> struct StatRegs {
    static constexpr bool ON = (BN == 64 && EPI == 0 /* EPI_CONV_STATS */);
    float s[ON ? 2 : 1][ON ? 32 : 1];
    float q[ON ? 2 : 1][ON ? 32 : 1];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int a = 0; a < (ON ? 2 : 1); ++a)
#pragma unroll
            for (int b = 0; b < (ON ? 32 : 1); ++b) { s[a][b] = 0.f; q[a][b] = 0.f; }
    }
};

// End of a CONV_STATS kernel: the epilogue warps of a CTA combine their per-quadrant column sums in
// shared memory (fixed order) and write ONE partial row [2][BN] per CTA, so the finalisation kernel
// walks 4x fewer rows. `red` = [4 quadrants][2][BN] floats of shared memory.
template <int BN>
__device__ __forceinline__ void write_cta_stats(float* red, float* dst_row, int warp, int lane,
                                                int quad, int chalf,
                                                const float (&ssum)[EpiCfg<BN>::NCH],
                                                const float (&ssq)[EpiCfg<BN>::NCH]) {
    constexpr int NCH = EpiCfg<BN>::NCH;
    constexpr int NT = 128 * EpiCfg<BN>::HALVES;   // epilogue threads
    float* mine = red + quad * (2 * BN) + chalf * (NCH * 32);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        mine[c * 32 + lane] = ssum[c];
        mine[BN + c * 32 + lane] = ssq[c];
    }
    asm volatile("bar.sync 1, %0;" ::"r"(NT) : "memory");   // epilogue warps only
    const int e = (warp < 6 ? warp - 2 : warp - 3) * 32 + lane;   // 0 .. NT-1
    for (int i = e; i < 2 * BN; i += NT)
        dst_row[i] = ((red[i] + red[2 * BN + i]) + red[4 * BN + i]) + red[6 * BN + i];
}

template <int BN, int EPI>
__device__ __forceinline__ void finish_stat_regs(StatRegs<BN, EPI>& sr, int lane,
                                                 float (&ssum)[EpiCfg<BN>::NCH],
                                                 float (&ssq)[EpiCfg<BN>::NCH]) {
    if constexpr (StatRegs<BN, EPI>::ON) {
#pragma unroll
        for (int cl = 0; cl < 2; ++cl) {
            ssum[cl] = warp_column_sum(sr.s[cl], lane);
            ssq[cl] = warp_column_sum(sr.q[cl], lane);
        }
    }
}

// Epilogue of one 128 x BN accumulator tile: TMEM -> registers (32 columns at a time), bias /
// folded-BN affine, bf16 store (row-major, or 2x2 pixel-shuffle scatter for the transposed conv),
// and the per-channel sum / sum-of-squares of the BatchNorm statistics.
//   trow: TMEM address of this warp's lane quadrant and accumulator stage; m: GEMM row of the thread.
//   Eight epilogue warps share a tile: the warp of column half `chalf` handles chunks
//   [chalf*NCH, chalf*NCH + NCH) of its lane quadrant, so the BatchNorm-statistics
//   epilogue of short-K tiles keeps up with the MMAs.
template <int BN, int EPI>
__device__ __forceinline__ void epilogue_tile(const IgemmParams& p, uint32_t trow, long long m,
                                              bool valid, int n0, int lane, int chalf,
                                              float (&ssum)[EpiCfg<BN>::NCH],
                                              float (&ssq)[EpiCfg<BN>::NCH],
                                              StatRegs<BN, EPI>& sr, const float* cs,
                                              const float* hs = nullptr,
                                              const uint4* ypre = nullptr) {
                long long ct_row = 0;
                if (EPI == EPI_CONVT) {
                    const int w = (int)(m % p.ct_W);
                    const long long t = m / p.ct_W;
                    const int h = (int)(t % p.ct_H);
                    const long long n = t / p.ct_H;
                    ct_row = n * p.ct_sN + (long long)(2 * h) * p.ct_sH + (long long)(2 * w) * p.ct_sW;
                }
                float hacc[HEAD_EPI_MAX_CLASSES];
                if (EPI == EPI_AFFINE_RELU_HEAD) {
    #pragma unroll
                    for (int hc = 0; hc < HEAD_EPI_MAX_CLASSES; ++hc) hacc[hc] = 0.f;
                }
    #pragma unroll
                for (int cl = 0; cl < EpiCfg<BN>::NCH; ++cl) {
                    const int c = chalf * EpiCfg<BN>::NCH + cl;
                    uint32_t r[32];
                    tmem_ld_32x32(trow + (uint32_t)(c * 32), r);
                    tmem_ld_wait();
                    const int col0 = n0 + c * 32;
                    float v[32];
    #pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                    if (EPI == EPI_AFFINE_RELU || EPI == EPI_AFFINE_RELU_HEAD) {
    #pragma unroll
                        for (int i4 = 0; i4 < 8; ++i4) {
                            const float4 a4 = reinterpret_cast<const float4*>(cs + c * 32)[i4];
                            const float4 b4 = reinterpret_cast<const float4*>(cs + BN + c * 32)[i4];
                            v[4 * i4 + 0] = fmaxf(fmaf(v[4 * i4 + 0], a4.x, b4.x), 0.f);
                            v[4 * i4 + 1] = fmaxf(fmaf(v[4 * i4 + 1], a4.y, b4.y), 0.f);
                            v[4 * i4 + 2] = fmaxf(fmaf(v[4 * i4 + 2], a4.z, b4.z), 0.f);
                            v[4 * i4 + 3] = fmaxf(fmaf(v[4 * i4 + 3], a4.w, b4.w), 0.f);
                        }
                        if (EPI == EPI_AFFINE_RELU_HEAD) {
                            // BN = 64: this thread sees all 64 channels of its pixel over the two chunks
    #pragma unroll
                            for (int i = 0; i < 32; ++i)
                                v[i] = __bfloat162float(__float2bfloat16_rn(v[i]));   // as it would be stored
    #pragma unroll
                            for (int hc = 0; hc < HEAD_EPI_MAX_CLASSES; ++hc) {
                                if (hc < p.head_nc) {
                                    const float4* w4 = reinterpret_cast<const float4*>(hs + hc * 64 + c * 32);
    #pragma unroll
                                    for (int i4 = 0; i4 < 8; ++i4) {
                                        const float4 wv = w4[i4];
                                        hacc[hc] = fmaf(v[4 * i4 + 0], wv.x, hacc[hc]);
                                        hacc[hc] = fmaf(v[4 * i4 + 1], wv.y, hacc[hc]);
                                        hacc[hc] = fmaf(v[4 * i4 + 2], wv.z, hacc[hc]);
                                        hacc[hc] = fmaf(v[4 * i4 + 3], wv.w, hacc[hc]);
                                    }
                                }
                            }
                        }
                    } else if (p.bias != nullptr) {
    #pragma unroll
                        for (int i4 = 0; i4 < 8; ++i4) {
                            const float4 a4 = reinterpret_cast<const float4*>(cs + c * 32)[i4];
                            v[4 * i4 + 0] += a4.x; v[4 * i4 + 1] += a4.y;
                            v[4 * i4 + 2] += a4.z; v[4 * i4 + 3] += a4.w;
                        }
                    }
                    if (EPI == EPI_STORE_BNRED) {
                        // g as it is stored (bf16), then dyh and dyh * (y - mean) from the prefetched y
    #pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __bfloat162float(__float2bfloat16_rn(v[i]));
                    }
                    if (valid && EPI != EPI_AFFINE_RELU_HEAD) {
                        __nv_bfloat16* dst;
                        if (EPI == EPI_CONVT) {
                            const int qd = col0 / p.ct_cout;
                            const int co = col0 - qd * p.ct_cout;
                            dst = p.out + ct_row + (long long)(qd >> 1) * p.ct_sH +
                                  (long long)(qd & 1) * p.ct_sW + co;
                        } else {
                            dst = p.out + m * p.ldo + col0;
                        }
                        uint4* d4 = reinterpret_cast<uint4*>(dst);
    #pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 o;
                            o.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
                            o.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                            o.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
                            o.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                            d4[j] = o;
                        }
                    }
                    if (EPI == EPI_STORE_BNRED) {
                        float s2[32];
    #pragma unroll
                        for (int i4 = 0; i4 < 8; ++i4) {
                            const float4 a4 = reinterpret_cast<const float4*>(cs + c * 32)[i4];
                            const float4 b4 = reinterpret_cast<const float4*>(cs + BN + c * 32)[i4];
                            const float4 m4 = reinterpret_cast<const float4*>(cs + 2 * BN + c * 32)[i4];
                            const uint4 yr = ypre[cl * 4 + (i4 >> 1)];
                            const uint32_t y01 = (i4 & 1) ? yr.z : yr.x, y23 = (i4 & 1) ? yr.w : yr.y;
                            const float yv[4] = {bf16_lo(y01), bf16_hi(y01), bf16_lo(y23), bf16_hi(y23)};
                            const float sc[4] = {a4.x, a4.y, a4.z, a4.w}, sh[4] = {b4.x, b4.y, b4.z, b4.w};
                            const float mu[4] = {m4.x, m4.y, m4.z, m4.w};
    #pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int i = 4 * i4 + e;
                                const float dyh = (valid && fmaf(yv[e], sc[e], sh[e]) > 0.f) ? v[i] : 0.f;
                                v[i] = dyh;
                                s2[i] = dyh * (yv[e] - mu[e]);
                            }
                        }
                        ssum[cl] += warp_column_sum(v, lane);
                        ssq[cl] += warp_column_sum(s2, lane);
                    } else if (EPI == EPI_CONV_STATS && StatRegs<BN, EPI>::ON) {
    #pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float x = valid ? v[i] : 0.f;
                            sr.s[StatRegs<BN, EPI>::ON ? cl : 0][StatRegs<BN, EPI>::ON ? i : 0] += x;
                            sr.q[StatRegs<BN, EPI>::ON ? cl : 0][StatRegs<BN, EPI>::ON ? i : 0] =
                                fmaf(x, x, sr.q[StatRegs<BN, EPI>::ON ? cl : 0][StatRegs<BN, EPI>::ON ? i : 0]);
                        }
                    } else if (EPI == EPI_CONV_STATS) {
                        float s2[32];
    #pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            v[i] = valid ? v[i] : 0.f;
                            s2[i] = v[i] * v[i];
                        }
                        ssum[cl] += warp_column_sum(v, lane);
                        ssq[cl] += warp_column_sum(s2, lane);
                    }
                }
                if (EPI == EPI_AFFINE_RELU_HEAD && valid) {
                    // consecutive lanes = consecutive pixels: coalesced NCHW logits / mask stores
                    const long long n = m / p.head_hw, hw = m % p.head_hw;
    #pragma unroll
                    for (int hc = 0; hc < HEAD_EPI_MAX_CLASSES; ++hc) {
                        if (hc < p.head_nc) {
                            hacc[hc] += hs[HEAD_EPI_MAX_CLASSES * 64 + hc];
                            p.head_logits[(n * p.head_nc + hc) * p.head_hw + hw] = hacc[hc];
                        }
                    }
                    if (p.head_mask)   // one class: sigmoid(z) > 0.5 == z > 0 (scripts/inference.py:39,85)
                        p.head_mask[m] = (p.head_nc >= 2 ? hacc[1] > hacc[0] : hacc[0] > 0.f) ? 255 : 0;
                }
}

// Eval epilogue of a TWO-ROW tile with the 2x2 floor max-pool of the reference's Down block fused in
// (nn.MaxPool2d(2), models/unet_model.py:28): folded BN + ReLU on both output rows of the tile, both
// stored as the activation (it is the skip connection), and their 2x2 maxima — the vertical pair
// lives in the same thread's two accumulators, the horizontal pair in the neighbouring lane — stored
// to the pooled tensor. max commutes with the (monotone) bf16 rounding, so the result equals pooling
// the stored activation.  trow0 / trow1: TMEM addresses of the two rows; m0: GEMM row of the thread in
// output row `pr` (even), q its column.
template <int BN>
__device__ __forceinline__ void epilogue_tile_pool2(const IgemmParams& p, uint32_t trow0, uint32_t trow1,
                                                    long long m0, int n, int pr, int q, bool valid0,
                                                    bool valid1, int n0, int lane, int chalf,
                                                    const float* cs) {
    const bool pool_ok = valid1 && (q + 1 < p.pool_Wo) && ((q & 1) == 0);   // valid1 implies valid0
#pragma unroll
    for (int cl = 0; cl < EpiCfg<BN>::NCH; ++cl) {
        const int c = chalf * EpiCfg<BN>::NCH + cl;
        const int col0 = n0 + c * 32;
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(trow0 + (uint32_t)(c * 32), r0);
        tmem_ld_32x32(trow1 + (uint32_t)(c * 32), r1);
        tmem_ld_wait();
        float mx[32];
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
            const float4 a4 = reinterpret_cast<const float4*>(cs + c * 32)[i4];
            const float4 b4 = reinterpret_cast<const float4*>(cs + BN + c * 32)[i4];
            const float sc[4] = {a4.x, a4.y, a4.z, a4.w}, sh[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = 4 * i4 + e;
                const float v0 = fmaxf(fmaf(__uint_as_float(r0[i]), sc[e], sh[e]), 0.f);
                const float v1 = fmaxf(fmaf(__uint_as_float(r1[i]), sc[e], sh[e]), 0.f);
                r0[i] = __float_as_uint(v0);
                r1[i] = __float_as_uint(v1);
                mx[i] = fmaxf(v0, v1);
            }
        }
        if (valid0) {
            uint4* d4 = reinterpret_cast<uint4*>(p.out + m0 * p.ldo + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                d4[j] = make_uint4(pack_bf16x2(__uint_as_float(r0[8 * j + 0]), __uint_as_float(r0[8 * j + 1])),
                                   pack_bf16x2(__uint_as_float(r0[8 * j + 2]), __uint_as_float(r0[8 * j + 3])),
                                   pack_bf16x2(__uint_as_float(r0[8 * j + 4]), __uint_as_float(r0[8 * j + 5])),
                                   pack_bf16x2(__uint_as_float(r0[8 * j + 6]), __uint_as_float(r0[8 * j + 7])));
        }
        if (valid1) {
            uint4* d4 = reinterpret_cast<uint4*>(p.out + (m0 + p.pool_Wo) * p.ldo + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                d4[j] = make_uint4(pack_bf16x2(__uint_as_float(r1[8 * j + 0]), __uint_as_float(r1[8 * j + 1])),
                                   pack_bf16x2(__uint_as_float(r1[8 * j + 2]), __uint_as_float(r1[8 * j + 3])),
                                   pack_bf16x2(__uint_as_float(r1[8 * j + 4]), __uint_as_float(r1[8 * j + 5])),
                                   pack_bf16x2(__uint_as_float(r1[8 * j + 6]), __uint_as_float(r1[8 * j + 7])));
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
        if (pool_ok) {
            const long long pm = ((long long)n * (p.pool_Ho >> 1) + (pr >> 1)) * (p.pool_Wo >> 1) + (q >> 1);
            uint4* d4 = reinterpret_cast<uint4*>(p.pooled + pm * p.ldo + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                d4[j] = make_uint4(pack_bf16x2(mx[8 * j + 0], mx[8 * j + 1]), pack_bf16x2(mx[8 * j + 2], mx[8 * j + 3]),
                                   pack_bf16x2(mx[8 * j + 4], mx[8 * j + 5]), pack_bf16x2(mx[8 * j + 6], mx[8 * j + 7]));
        }
    }
}

// EPI_STORE_BNRED: fetch the y values of this thread's row (NCH chunks x 32 columns = NCH x 4 x 16 bytes)
// BEFORE waiting for the accumulator, so that their HBM latency hides behind the tile's MMAs.
template <int BN, int EPI>
__device__ __forceinline__ void bnred_prefetch(const IgemmParams& p, long long m, bool valid, int n0,
                                               int chalf, uint4 (&ypre)[EpiCfg<BN>::NCH * 4]) {
    if constexpr (EPI == EPI_STORE_BNRED) {
#pragma unroll
        for (int cl = 0; cl < EpiCfg<BN>::NCH; ++cl) {
            const int c = chalf * EpiCfg<BN>::NCH + cl;
            const uint4* src = reinterpret_cast<const uint4*>(p.red_y + m * p.ldo + n0 + c * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                ypre[cl * 4 + j] = valid ? __ldg(src + j) : make_uint4(0u, 0u, 0u, 0u);
        }
    }
}

// Per-column epilogue constants of n tile [n0, n0+BN) -> shared memory (broadcast LDS.128 reads
// instead of 32-64 cached global loads per 32-column chunk and tile):
//   cs[0..BN) = scale (AFFINE_RELU kinds) or bias (other kinds, 0 if absent), cs[BN..2BN) = shift.
template <int BN, int EPI>
__device__ __forceinline__ void stage_epilogue_consts(const IgemmParams& p, int n0, float* cs) {
    for (int i = threadIdx.x; i < BN; i += blockDim.x) {
        if (EPI == EPI_AFFINE_RELU || EPI == EPI_AFFINE_RELU_HEAD || EPI == EPI_STORE_BNRED) {
            cs[i] = p.scale[n0 + i];
            cs[BN + i] = p.shift[n0 + i];
            if (EPI == EPI_STORE_BNRED) cs[2 * BN + i] = p.red_mean[n0 + i];
        } else {
            cs[i] = p.bias ? p.bias[n0 + i] : 0.f;
            cs[BN + i] = 0.f;
        }
    }
    __syncthreads();
}

// Head weights of EPI_AFFINE_RELU_HEAD -> shared memory: hs[c*64 + k] (c < 8), bias at hs[512 + c].
template <int EPI>
__device__ __forceinline__ void stage_head_weights(const IgemmParams& p, float* hs) {
    if constexpr (EPI == EPI_AFFINE_RELU_HEAD) {
        for (int i = threadIdx.x; i < HEAD_EPI_MAX_CLASSES * 64; i += blockDim.x)
            hs[i] = i < p.head_nc * 64 ? p.head_w[i] : 0.f;
        for (int i = threadIdx.x; i < HEAD_EPI_MAX_CLASSES; i += blockDim.x)
            hs[HEAD_EPI_MAX_CLASSES * 64 + i] = (p.head_b && i < p.head_nc) ? p.head_b[i] : 0.f;
        __syncthreads();
    }
}

template <int BN, int EPI, int CG>
__global__ void __launch_bounds__(igemm_threads(BN), 1)
igemm_kmajor_kernel(const __grid_constant__ CUtensorMap mapA0,
                    const __grid_constant__ CUtensorMap mapA1,
                    const __grid_constant__ CUtensorMap mapB, const IgemmParams p) {
    pdl_trigger();   // let the next kernel's CTAs be scheduled while this grid drains
    using Cfg = IgemmCfg<BN, CG>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - raw);
    const uint32_t bar_base = base + Cfg::STAGES * Cfg::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 4);
    volatile uint32_t* tmem_slot_g =
        reinterpret_cast<volatile uint32_t*>(gbase + Cfg::STAGES * Cfg::STAGE_BYTES +
                                             8 * (2 * Cfg::STAGES + 4));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // pair geometry: `unit` = CTA (CG 1) or CTA pair (CG 2); rank 0 issues the MMAs
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const int unit = blockIdx.x / CG;
    const int nunits = gridDim.x / CG;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA0);
        tma_prefetch_desc(&mapA1);
        tma_prefetch_desc(&mapB);
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(full_bar(s), 2);   // the leader's two producer threads (A operand, B operand)
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 4 * EpiCfg<BN>::HALVES * CG);   // epilogue warps of the unit
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_cg<Cfg::TMEM_COLS, CG>(tmem_slot);
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_g;
    pdl_wait();   // prologue done; from here on global memory of the preceding kernels is read
    float* hs = reinterpret_cast<float*>(gbase + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES);
    float* cs = reinterpret_cast<float*>(gbase + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES +
                                         Cfg::STAT_BYTES);
    stage_head_weights<EPI>(p, hs);
    stage_epilogue_consts<BN, EPI>(p, (unit % p.n_tiles) * BN, cs);

    const int cchunks = p.cchunks0 + p.cchunks1;
    const int kblocks = p.taps * cchunks;
    const int cta_n = unit % p.n_tiles;
    const int m_first = unit / p.n_tiles;
    const int m_step = nunits / p.n_tiles;
    const int m_units = (p.m_tiles + CG - 1) / CG;
    const int n0 = cta_n * BN;

    // Role code is executed by WHOLE warps with warp-uniform control flow; only the instruction that
    // must be issued once (TMA, MMA, commit, expect_tx) sits under elect_one(). In a single-lane
    // divergent branch the compiler has to wrap every TMA/MMA issue in an elect + R2UR.BROADCAST
    // waterfall loop to obtain uniform-register operands, which costs hundreds of cycles per issue.
    if (warp == 0) {
        // ------------------------------ TMA producer, A operand (im2col) ------------------------------
        // The two operands are fed by two different warps; each arms the leader's stage barrier with
        // the bytes of the whole unit (in a pair the peer's loads complete on the leader's barrier).
        int stage = 0;
        uint32_t phase = 0;
        for (int mu = m_first; mu < m_units; mu += m_step) {
            int mt = mu * CG + (int)rank;
            if (mt >= p.m_tiles) mt = p.m_tiles - 1;   // odd tail: reload a valid tile, rows masked
            const int m0 = mt * 128;
            const int q = m0 % p.Wo;
            const int t = m0 / p.Wo;
            const int pr = t % p.Ho;
            const int n = t / p.Ho;
            const int cw = p.lower + q * p.tstride;
            const int ch = p.lower + pr * p.tstride;
            int cc = 0;
            uint16_t offw = 0, offh = 0;
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                const uint32_t sa = base + stage * Cfg::STAGE_BYTES;
                const uint32_t fb = (CG == 2) ? mapa_rank(full_bar(stage), 0) : full_bar(stage);
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(full_bar(stage), CG * Cfg::A_BYTES);
                    if (cc < p.cchunks0)
                        tma_load_im2col_cg<CG>(sa, &mapA0, fb, cc * 64, cw, ch, n, offw, offh);
                    else
                        tma_load_im2col_cg<CG>(sa, &mapA1, fb, (cc - p.cchunks0) * 64, cw, ch, n,
                                               offw, offh);
                }
                __syncwarp();
                // K order = (channel chunk, tap): the same summation order as the row-run kernel, so
                // a layer gives bit-identical results whichever kernel its width selects.
                if (++offw == (uint16_t)p.tapw) {
                    offw = 0;
                    if (++offh == (uint16_t)(p.taps / p.tapw)) { offh = 0; ++cc; }
                }
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 6) {
        // ------------------------------ TMA producer, B operand (weights) ------------------------------
        int stage = 0;
        uint32_t phase = 0;
        const int nrow0 = n0 + (int)rank * Cfg::B_ROWS;
        for (int mu = m_first; mu < m_units; mu += m_step) {
            int cc = 0, tap = 0;
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                const uint32_t fb = (CG == 2) ? mapa_rank(full_bar(stage), 0) : full_bar(stage);
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(full_bar(stage), CG * Cfg::B_BYTES);
                    tma_load_3d_cg<CG>(base + stage * Cfg::STAGE_BYTES + Cfg::A_BYTES, &mapB, fb,
                                       cc * 64, nrow0, tap);
                }
                __syncwarp();
                if (++tap == p.taps) { tap = 0; ++cc; }
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ------------------------------ MMA issuer (leader CTA) ------------------------------
        constexpr uint32_t idesc = make_idesc_bf16(128 * CG, BN, 0, 0);
        int stage = 0;
        uint32_t phase = 0;
        int as = 0;
        uint32_t aphase = 0;
        for (int mu = m_first; mu < m_units; mu += m_step) {
            mbar_wait(tempty_bar(as), aphase ^ 1u);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t sa = base + stage * Cfg::STAGE_BYTES;
                const uint32_t sb = sa + Cfg::A_BYTES;
                const uint64_t da = make_smem_desc(sa, 0, 1024);
                const uint64_t db = make_smem_desc(sb, 0, 1024);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)  // 16 bf16 = 32 B along K => +2 in 16-byte units
                        umma_bf16_cg<CG>(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                         (uint32_t)((kb | k) != 0));
                    umma_commit_cg<CG>(empty_bar(stage));
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
            }
            if (elect_one()) umma_commit_cg<CG>(tfull_bar(as));
            __syncwarp();
            if (++as == 2) { as = 0; aphase ^= 1u; }
        }
    } else if (is_epilogue_warp<BN>(warp)) {
        // ------------------------------ epilogue ------------------------------
        const int quad = warp & 3;  // TMEM lane quadrant this warp may access
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
            const long long m = (long long)(mu * CG + (int)rank) * 128 + row_in_tile;
            const bool valid = m < p.M;
            uint4 ypre[EPI == EPI_STORE_BNRED ? NCH * 4 : 1];
            if constexpr (EPI == EPI_STORE_BNRED) bnred_prefetch<BN, EPI>(p, m, valid, n0, chalf, ypre);
            mbar_wait(tfull_bar(as), aphase);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN);

            epilogue_tile<BN, EPI>(p, trow, m, valid, n0, lane, chalf, ssum, ssq, sr, cs, hs, ypre);
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
            // partial row of "virtual CTA" rank*nunits + unit: nunits % n_tiles == 0, so the channel
            // tile of a row is still (row index % n_tiles) for bn_finalize_kernel
            float* red = reinterpret_cast<float*>(gbase + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES);
            write_cta_stats<BN>(red, p.stats + (long long)((int)rank * nunits + unit) * (2 * BN), warp,
                                lane, quad, chalf, ssum, ssq);
        }
    }

    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_cg<Cfg::TMEM_COLS, CG>(tmem_base);
    }
}

// =================================================================================================
// Weight gradient: D[(tap, ca), cb] = sum_pix  Aact[pix + tap, ca] * Bmat[pix, cb]
//   A side ("im2col'd activations", up to two sources): 64-row chunks indexed (tap, channel chunk)
//   B side: plain [pixels][channels] matrix. Both MN-major; K = pixels, 64 per pipeline stage.
//   conv3x3:  A = layer input x,  B = dY       -> D[(tap,ci), co]
//   convT2x2: A = d(up) (stride 2, 4 taps),  B = layer input x -> D[(q,co), ci]
// =================================================================================================
struct WgradParams {
    int Mpix;
    int Wo, Ho, lower, tstride, taps, tapw;
    int cchunks0, cchunks1;
    int a_chunks_total;  // taps * (cchunks0 + cchunks1)
    int n_tiles, splits;
    int kblocks_total;   // ceil(Mpix / KPIX)
    float* ws;           // [splits][a_chunks_total*64][ldw]
    long long ldw;
    long long split_stride;
};

template <int BN, int CG = 1>
struct WgradCfg {
    // Pixels (GEMM K) per pipeline stage. One barrier wait + tcgen05 fence costs the MMA thread
    // ~230 cycles, so narrow tiles (BN <= 128, 48-64 cycles per MMA) take 128 pixels = 8 MMAs per
    // stage; BN = 256 is already execution-bound with 64.
    // BN = 192 is the "shifted dY" form for 64-output-channel layers (see igemm_wgrad_kernel): three
    // 64-column copies of dY displaced by 0 / 1 / 2 pixels along w share every read of the A operand.
    // BN = 384 is the same idea for 128 output channels as a CTA pair: an N = 256 MMA over dY displaced by
    // 0 and 1 pixels (CTA r stages the copy displaced by r) plus an N = 128 MMA over the copy displaced by 2
    // (CTA r stages channel chunk r of it), both fed by one staging of A.
    static constexpr bool SHIFT = (BN == 192 || BN == 384);
    static constexpr int KPIX = (BN >= 192) ? 64 : 128;
    static constexpr int CHUNK_BYTES = KPIX * 128;   // one [KPIX pixels][64 channels] MN-major chunk
    static constexpr int A_BYTES = 2 * CHUNK_BYTES;  // per CTA: two 64-row chunks = 128 GEMM rows
    static constexpr int B_CHUNKS = BN / 64 / CG;    // per CTA: BN / CG columns of dY
    static constexpr int B_BYTES = B_CHUNKS * CHUNK_BYTES;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = SHIFT ? 5
                                  : (CG == 2) ? (BN == 256 ? 6 : 4)
                                              : ((BN == 256) ? 4 : (BN == 128 ? 3 : 4));
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;
    static constexpr uint32_t TMEM_COLS = BN == 192 ? 256 : (BN == 384 ? 512 : BN);   // powers of two
    static_assert(CG == 1 || BN >= 128, "a CTA pair splits dY by 64-channel chunks");
    static_assert(!SHIFT || (BN == 192 && CG == 1) || (BN == 384 && CG == 2),
                  "shifted forms: 192 columns on one CTA, 384 on a pair");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

// BN = 192 ("shifted dY", WgradCfg::SHIFT): for a 3x3 layer with 64 output channels the plain form runs
// M128 x N64 MMAs, each re-reading 4 KB of MN-major A from shared memory (64 B/clk) for 32 cycles of
// math. Here the GEMM columns are (filter column dx, output channel): three copies of the dY tile
// displaced by 0 / 1 / 2 pixels along w (im2col loads over the input-wide pixel grid, zero filled
// outside dY), so one read of A serves three taps; the GEMM rows are (filter row dy, input channel).
template <int BN, int CG>
__global__ void __launch_bounds__(256, 1)
igemm_wgrad_kernel(const __grid_constant__ CUtensorMap mapA0,
                   const __grid_constant__ CUtensorMap mapA1,
                   const __grid_constant__ CUtensorMap mapB, const WgradParams p) {
    pdl_trigger();   // let the next kernel's CTAs be scheduled while this grid drains
    using Cfg = WgradCfg<BN, CG>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - raw);
    const uint32_t bar_base = base + Cfg::STAGES * Cfg::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
    const uint32_t tfull_bar = bar_base + 8u * (2 * Cfg::STAGES);
    const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 1);
    volatile uint32_t* tmem_slot_g = reinterpret_cast<volatile uint32_t*>(
        gbase + Cfg::STAGES * Cfg::STAGE_BYTES + 8 * (2 * Cfg::STAGES + 1));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const int unit = blockIdx.x / CG;              // (m tile of 128*CG rows, n tile)
    const int m_tile = unit / p.n_tiles;
    const int n_tile = unit % p.n_tiles;
    // 64-row A chunks of this CTA: c = (m_tile*CG + rank)*2 + {0, 1}; the leader's come first
    const int chunk0 = (m_tile * CG + (int)rank) * 2;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA0);
        tma_prefetch_desc(&mapA1);
        tma_prefetch_desc(&mapB);
        // threads arming a stage (leader CTA): the B thread plus one per existing 64-row A chunk
        const uint32_t nprod = (chunk0 + 1 < p.a_chunks_total) ? 3u : 2u;
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(full_bar(s), nprod);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tfull_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_cg<Cfg::TMEM_COLS, CG>(tmem_slot);
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_g;
    pdl_wait();   // prologue done; from here on global memory of the preceding kernels is read

    const int split = blockIdx.y;
    const int kb_begin = (int)((long long)p.kblocks_total * split / p.splits);
    const int kb_end = (int)((long long)p.kblocks_total * (split + 1) / p.splits);
    const int cchunks = p.cchunks0 + p.cchunks1;
    const int n0 = n_tile * BN;

    // Three producer threads (different warps) feed one stage: the two 64-row im2col chunks of the
    // A side and the B-side box. In a pair every load completes on the LEADER's stage barrier, which
    // the leader's producers arm with the bytes of both CTAs.
    if (warp == 0 || warp == 6 || warp == 7) {
        const int role = warp == 0 ? 0 : (warp == 6 ? 1 : 2);   // 0/1: A chunk j, 2: B
        int stage = 0;
        uint32_t phase = 0;
        int a_cc = 0;
        uint16_t a_offw = 0, a_offh = 0;
        bool active = true;
        uint32_t tx = Cfg::B_BYTES * CG;
        if (role < 2) {
            const int c = chunk0 + role;
            active = c < p.a_chunks_total;
            // the peer's chunk of the same role (two chunks further) may lie past the end
            tx = (CG == 2 && c + 2 < p.a_chunks_total) ? 2u * Cfg::CHUNK_BYTES
                                                        : (uint32_t)Cfg::CHUNK_BYTES;
            if (active) {
                const int tap = c / cchunks;
                a_cc = c % cchunks;
                a_offw = (uint16_t)(tap % p.tapw);
                a_offh = (uint16_t)(tap / p.tapw);
            }
        }
        // base pixel (q, pr, n) of the first k-block of this split, then advanced by KPIX pixels
        int m0 = kb_begin * Cfg::KPIX;
        int q = m0 % p.Wo;
        int pr, n;
        { const int t = m0 / p.Wo; pr = t % p.Ho; n = t / p.Ho; }
        if (active) {
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                const uint32_t sa = base + stage * Cfg::STAGE_BYTES;
                const uint32_t fb = (CG == 2) ? mapa_rank(full_bar(stage), 0) : full_bar(stage);
                if (role < 2) {
                    const int cw = p.lower + q * p.tstride;
                    const int ch = p.lower + pr * p.tstride;
                    if (elect_one()) {
                        if (rank == 0) mbar_expect_tx(full_bar(stage), tx);
                        if (a_cc < p.cchunks0)
                            tma_load_im2col_cg<CG>(sa + role * Cfg::CHUNK_BYTES, &mapA0, fb, a_cc * 64,
                                                   cw, ch, n, a_offw, a_offh);
                        else
                            tma_load_im2col_cg<CG>(sa + role * Cfg::CHUNK_BYTES, &mapA1, fb,
                                                   (a_cc - p.cchunks0) * 64, cw, ch, n, a_offw, a_offh);
                    }
                    __syncwarp();
                    q += Cfg::KPIX;
                    while (q >= p.Wo) {
                        q -= p.Wo;
                        if (++pr == p.Ho) { pr = 0; ++n; }
                    }
                } else if (Cfg::SHIFT) {
                    // B side, shifted form: mapB is an im2col map of dY whose base pixels run over the
                    // INPUT-wide grid (w from -2, zero filled outside dY); chunk j is dY displaced by
                    // j pixels to the right, i.e. tap column dx = j of the filter.
                    if (elect_one()) {
                        if (rank == 0) mbar_expect_tx(full_bar(stage), tx);
                        const uint32_t sb0 = sa + Cfg::A_BYTES;
                        if (BN == 192) {
#pragma unroll
                            for (int j = 0; j < 3; ++j)
                                tma_load_im2col_cg<CG>(sb0 + j * Cfg::CHUNK_BYTES, &mapB, fb, 0, q - 2, pr,
                                                       n, (uint16_t)(2 - j), (uint16_t)0);
                        } else {
                            // pair: columns [128 r, 128 r + 128) of the N = 256 MMA = dY displaced by r
                            // (both channel chunks); columns [64 r, 64 r + 64) of the N = 128 MMA =
                            // channel chunk r of dY displaced by 2
                            const uint16_t ow = (uint16_t)(2 - (int)rank);
                            tma_load_im2col_cg<CG>(sb0, &mapB, fb, 0, q - 2, pr, n, ow, (uint16_t)0);
                            tma_load_im2col_cg<CG>(sb0 + Cfg::CHUNK_BYTES, &mapB, fb, 64, q - 2, pr, n,
                                                   ow, (uint16_t)0);
                            tma_load_im2col_cg<CG>(sb0 + 2 * Cfg::CHUNK_BYTES, &mapB, fb, 64 * (int)rank,
                                                   q - 2, pr, n, (uint16_t)0, (uint16_t)0);
                        }
                    }
                    __syncwarp();
                    q += Cfg::KPIX;
                    while (q >= p.Wo) {
                        q -= p.Wo;
                        if (++pr == p.Ho) { pr = 0; ++n; }
                    }
                } else {
                    // B side: one 3-D box (64 ch, KPIX pixels, B_CHUNKS chunks) -> [chunk][pixel][64 ch]
                    if (elect_one()) {
                        if (rank == 0) mbar_expect_tx(full_bar(stage), tx);
                        tma_load_3d_cg<CG>(sa + Cfg::A_BYTES, &mapB, fb, 0, m0,
                                           n_tile * (BN / 64) + (int)rank * Cfg::B_CHUNKS);
                    }
                    __syncwarp();
                    m0 += Cfg::KPIX;
                }
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1 && rank == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128 * CG, BN == 384 ? 256 : BN, 1, 1);
        constexpr uint32_t idesc2 = make_idesc_bf16(128 * CG, 128, 1, 1);   // BN = 384: third filter column
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t sa = base + stage * Cfg::STAGE_BYTES;
            const uint32_t sb = sa + Cfg::A_BYTES;
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < Cfg::KPIX / 16; ++k) {  // 16 pixels = 16 rows of 128 B
                    const uint64_t da = make_smem_desc(sa + k * 2048, Cfg::CHUNK_BYTES, 1024);
                    const uint64_t db = make_smem_desc(sb + k * 2048, Cfg::CHUNK_BYTES, 1024);
                    const uint32_t acc = (uint32_t)((kb != kb_begin) || (k != 0));
                    umma_bf16_cg<CG>(tmem_base, da, db, idesc, acc);
                    if (BN == 384) {
                        const uint64_t db2 = make_smem_desc(sb + 2 * Cfg::CHUNK_BYTES + k * 2048,
                                                            Cfg::CHUNK_BYTES, 1024);
                        umma_bf16_cg<CG>(tmem_base + 256u, da, db2, idesc2, acc);
                    }
                }
                umma_commit_cg<CG>(empty_bar(stage));
            }
            __syncwarp();
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) umma_commit_cg<CG>(tfull_bar);
        __syncwarp();
    } else if (warp >= 2 && warp < 6) {
        const int quad = warp & 3;
        const int row = (m_tile * CG + (int)rank) * 128 + quad * 32 + lane;
        const bool valid = row < p.a_chunks_total * 64;
        float* dst = p.ws + (long long)split * p.split_stride + (long long)row * p.ldw + n0;
        if (kb_end > kb_begin) {
            mbar_wait(tfull_bar, 0);
            tc_fence_after();
        }
        const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t r[32];
            if (kb_end > kb_begin) {
                tmem_ld_32x32(trow + (uint32_t)(c * 32), r);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = 0u;
            }
            if (valid) {
                uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    d4[j] = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
            }
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
