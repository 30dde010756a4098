// First DoubleConv convolution (inc.double_conv.0, reference models/unet_model.py:73 -> :11-13):
// C_in = n_channels (1 for DIC-C2DH-HeLa), K = 9*C_in, arithmetic intensity ~9 FLOP/B. It stays in
// fp32 on CUDA cores (rounding the low-contrast input image to bf16 is 75 % of the
// end-to-end bf16 drift) and its pre-BN output is never stored: statistics, BN-apply, and the
// backward pass recompute the 9-tap convolution from the fp32 input (HBM traffic = input read +
// one bf16 activation write).
//
// Thread mapping: one thread = one output pixel x 8 output channels (16 B bf16 store); the channel
// group of a thread is fixed (threadIdx.x % (Co/8)) so per-channel sums live in registers.
#pragma once
#include "common.cuh"
#include "elementwise.cuh"

namespace ub {

struct FirstConvArgs {
    const float* x;   // [N][Ci][H][W] fp32 (NCHW, as handed over by the caller)
    int N, Ci, H, W;  // input dims; output is (H-2) x (W-2)
    int Co;           // output channels (64); Co % 8 == 0, 256 % (Co/8) == 0
    const float* w;   // [Co][Ci][3][3] fp32 master weights
    const float* bias;   // [Co] or null
    const float* scale;  // apply / backward
    const float* shift;
    const float* mean;
    const float* rstd;
    __nv_bfloat16* a;    // apply: [N][H-2][W-2][Co] bf16
    float* partial;      // stats / bwd reduce: [gridDim.x][2][Co]
    View g;              // backward: gradient of a (direct view)
    const float* dgamma;
    const float* dbeta;
    float inv_count;
    float* wpartial;     // wgrad: [gridDim.x][Co][9] for input channel `ci_sel`
    int ci_sel;
};

enum : int { FC_STATS = 0, FC_APPLY = 1, FC_BWD_REDUCE = 2, FC_BWD_WGRAD = 3 };

// CI1 = true: single input channel (the reference's DIC-C2DH-HeLa case): the 9x8 weights of the
// thread's channel group live in registers, no shared-memory traffic in the inner loop.
template <int MODE, bool CI1>
static __global__ void __launch_bounds__(256)
first_conv_kernel(const FirstConvArgs A) {
    extern __shared__ float wsm[];  // [Ci*9][Co] weights, transposed for 8-wide reads
    const int Co = A.Co, CG = Co >> 3, Ci = A.Ci;
    const int cg = threadIdx.x % CG;
    float wr[9][8];
    if (CI1) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
#pragma unroll
            for (int k = 0; k < 8; ++k) wr[tap][k] = A.w[(cg * 8 + k) * 9 + tap];
    } else {
        for (int i = threadIdx.x; i < Co * Ci * 9; i += blockDim.x) {
            const int co = i / (Ci * 9), r = i % (Ci * 9);
            wsm[r * Co + co] = A.w[i];
        }
        __syncthreads();
    }
    const int Ho = A.H - 2, Wo = A.W - 2;

    float bi[8], sc[8], sh[8], mu[8], rs[8], kb[8], kg[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = cg * 8 + k;
        bi[k] = A.bias ? A.bias[c] : 0.f;
        if (MODE != FC_STATS) { sc[k] = A.scale[c]; sh[k] = A.shift[c]; }
        if (MODE >= FC_BWD_REDUCE) { mu[k] = A.mean[c]; rs[k] = A.rstd[c]; }
        if (MODE == FC_BWD_WGRAD) {
            kb[k] = A.dbeta[c] * A.inv_count;
            kg[k] = A.dgamma[c] * A.inv_count;
        }
    }
    float acc0[8], acc1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc0[k] = 0.f; acc1[k] = 0.f; }
    float wacc[9][8];
    if (MODE == FC_BWD_WGRAD) {
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int k = 0; k < 8; ++k) wacc[t][k] = 0.f;
    }

    const unsigned npix = (unsigned)A.N * Ho * Wo;
    const unsigned gstride = gridDim.x * 256u / CG;
    for (unsigned pq = (blockIdx.x * 256u + threadIdx.x) / CG; pq < npix; pq += gstride) {
        const size_t i = (size_t)pq * CG + cg;
        const unsigned wq = pq % (unsigned)Wo, tq = pq / (unsigned)Wo;
        const unsigned hq = tq % (unsigned)Ho, n = tq / (unsigned)Ho;
        float y[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = bi[k];
        float xsel[9];
        if (CI1) {
            const float* xp = A.x + ((size_t)n * A.H + hq) * A.W + wq;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const float xv = __ldg(xp + (tap / 3) * A.W + (tap % 3));
                xsel[tap] = xv;
#pragma unroll
                for (int k = 0; k < 8; ++k) y[k] = fmaf(xv, wr[tap][k], y[k]);
            }
        } else
        for (int ci = 0; ci < Ci; ++ci) {
            const float* xp = A.x + (((size_t)n * Ci + ci) * A.H + hq) * A.W + wq;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const float xv = __ldg(xp + (tap / 3) * A.W + (tap % 3));
                if (MODE == FC_BWD_WGRAD && ci == A.ci_sel) xsel[tap] = xv;
                const float4 w0 = *reinterpret_cast<const float4*>(&wsm[(ci * 9 + tap) * Co + cg * 8]);
                const float4 w1 =
                    *reinterpret_cast<const float4*>(&wsm[(ci * 9 + tap) * Co + cg * 8 + 4]);
                y[0] = fmaf(xv, w0.x, y[0]); y[1] = fmaf(xv, w0.y, y[1]);
                y[2] = fmaf(xv, w0.z, y[2]); y[3] = fmaf(xv, w0.w, y[3]);
                y[4] = fmaf(xv, w1.x, y[4]); y[5] = fmaf(xv, w1.y, y[5]);
                y[6] = fmaf(xv, w1.z, y[6]); y[7] = fmaf(xv, w1.w, y[7]);
            }
        }
        if (MODE == FC_STATS) {
#pragma unroll
            for (int k = 0; k < 8; ++k) { acc0[k] += y[k]; acc1[k] = fmaf(y[k], y[k], acc1[k]); }
        } else if (MODE == FC_APPLY) {
            Vec8 o;
#pragma unroll
            for (int k = 0; k < 8; ++k) o.v[k] = fmaxf(fmaf(y[k], sc[k], sh[k]), 0.f);
            *reinterpret_cast<uint4*>(A.a + i * 8) = pack8(o);
        } else {
            const __nv_bfloat16* gptr = reinterpret_cast<const __nv_bfloat16*>(A.g.ptr) +
                                        (size_t)(n * A.g.sN + hq * A.g.sH + wq * A.g.sW) + cg * 8;
            const Vec8 gv = unpack8(ldg16(gptr));
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float act = fmaf(y[k], sc[k], sh[k]);
                const float dyh = act > 0.f ? gv.v[k] : 0.f;
                const float xh = (y[k] - mu[k]) * rs[k];
                if (MODE == FC_BWD_REDUCE) {
                    acc0[k] += dyh;
                    acc1[k] = fmaf(dyh, xh, acc1[k]);
                } else {
                    const float dy = sc[k] * (dyh - kb[k] - xh * kg[k]);
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) wacc[tap][k] = fmaf(dy, xsel[tap], wacc[tap][k]);
                }
            }
        }
    }

    if (MODE == FC_STATS || MODE == FC_BWD_REDUCE) {
        __shared__ float red[256 * 16];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            red[threadIdx.x * 16 + k] = acc0[k];
            red[threadIdx.x * 16 + 8 + k] = acc1[k];
        }
        __syncthreads();
        for (int j = threadIdx.x; j < CG * 16; j += blockDim.x) {
            const int g2 = j / 16, e = j % 16;
            float s = 0.f;
            for (int tt = g2; tt < 256; tt += CG) s += red[tt * 16 + e];
            A.partial[(long long)blockIdx.x * 2 * Co + (e < 8 ? 0 : Co) + g2 * 8 + (e & 7)] = s;
        }
    } else if (MODE == FC_BWD_WGRAD) {
        __shared__ float red[256 * 9];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            __syncthreads();
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) red[threadIdx.x * 9 + tap] = wacc[tap][k];
            __syncthreads();
            for (int j = threadIdx.x; j < CG * 9; j += blockDim.x) {
                const int g2 = j / 9, tap = j % 9;
                float s = 0.f;
                for (int tt = g2; tt < 256; tt += CG) s += red[tt * 9 + tap];
                A.wpartial[((long long)blockIdx.x * Co + g2 * 8 + k) * 9 + tap] = s;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// C_in == 1 fast path (the reference's single-channel microscopy images): one thread = PX
// horizontally adjacent output pixels x 8 channels. The 3 x (PX+2) input patch and the 9x8 weights
// live in registers, so an item costs 3*(PX+2) cached loads for 72*PX FMAs and PX 16-byte stores /
// gradient loads are in flight together. The conv bias is folded into the BN constants.
// ---------------------------------------------------------------------------------------------
template <int MODE, int PX>
static __global__ void __launch_bounds__(256, (MODE >= FC_BWD_REDUCE) ? 1 : 2)
first_conv1_kernel(const FirstConvArgs A) {
    const int Co = A.Co, CG = Co >> 3;
    const unsigned cg = threadIdx.x % CG;
    const unsigned Ho = A.H - 2, Wo = A.W - 2, W = A.W;
    const unsigned GW = (Wo + PX - 1) / PX;  // pixel groups per output row
    float wr[9][8];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int k = 0; k < 8; ++k) wr[tap][k] = A.w[(cg * 8 + k) * 9 + tap];
    // y = conv + bias;  act = y*sc + sh = conv*sc + shb;  xhat = (y - mu)*rs = (conv - mub)*rs
    float bi[8], sc[8], shb[8], mub[8], rs[8], kb[8], kg[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = cg * 8 + k;
        bi[k] = A.bias ? A.bias[c] : 0.f;
        if (MODE != FC_STATS) { sc[k] = A.scale[c]; shb[k] = fmaf(bi[k], sc[k], A.shift[c]); }
        if (MODE >= FC_BWD_REDUCE) { mub[k] = A.mean[c] - bi[k]; rs[k] = A.rstd[c]; }
        if (MODE == FC_BWD_WGRAD) {
            kb[k] = A.dbeta[c] * A.inv_count;
            kg[k] = A.dgamma[c] * A.inv_count;
        }
    }
    float acc0[8], acc1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc0[k] = 0.f; acc1[k] = 0.f; }
    float wacc[9][8];
    if (MODE == FC_BWD_WGRAD) {
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int k = 0; k < 8; ++k) wacc[t][k] = 0.f;
    }
    const unsigned ngroups = (unsigned)A.N * Ho * GW;
    const unsigned gstride = gridDim.x * 256u / CG;
    const __nv_bfloat16* gb = reinterpret_cast<const __nv_bfloat16*>(A.g.ptr);
    for (unsigned gi = (blockIdx.x * 256u + threadIdx.x) / CG; gi < ngroups; gi += gstride) {
        const unsigned gw = gi % GW, t = gi / GW, hq = t % Ho, n = t / Ho;
        const unsigned wq0 = gw * PX;
        float xp[3][PX + 2];
        const float* xrow = A.x + ((size_t)n * A.H + hq) * W + wq0;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < PX + 2; ++c) xp[r][c] = (wq0 + c < W) ? __ldg(xrow + r * W + c) : 0.f;
        uint4 graw[PX];
        if (MODE >= FC_BWD_REDUCE) {
#pragma unroll
            for (int j = 0; j < PX; ++j)
                if (wq0 + j < Wo)
                    graw[j] = ldg16(gb + (size_t)(n * A.g.sN + hq * A.g.sH + (wq0 + j) * A.g.sW) +
                                    cg * 8);
        }
        const size_t pq0 = ((size_t)n * Ho + hq) * Wo + wq0;
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            if (wq0 + j < Wo) {
                float y[8];  // raw convolution (bias folded into the constants)
#pragma unroll
                for (int k = 0; k < 8; ++k) y[k] = 0.f;
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const float xv = xp[tap / 3][j + tap % 3];
#pragma unroll
                    for (int k = 0; k < 8; ++k) y[k] = fmaf(xv, wr[tap][k], y[k]);
                }
                if (MODE == FC_STATS) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float yb = y[k] + bi[k];
                        acc0[k] += yb;
                        acc1[k] = fmaf(yb, yb, acc1[k]);
                    }
                } else if (MODE == FC_APPLY) {
                    Vec8 o;
#pragma unroll
                    for (int k = 0; k < 8; ++k) o.v[k] = fmaxf(fmaf(y[k], sc[k], shb[k]), 0.f);
                    *reinterpret_cast<uint4*>(A.a + (pq0 + j) * Co + cg * 8) = pack8(o);
                } else {
                    const Vec8 gv = unpack8(graw[j]);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float act = fmaf(y[k], sc[k], shb[k]);
                        const float dyh = act > 0.f ? gv.v[k] : 0.f;
                        const float xh = (y[k] - mub[k]) * rs[k];
                        if (MODE == FC_BWD_REDUCE) {
                            acc0[k] += dyh;
                            acc1[k] = fmaf(dyh, xh, acc1[k]);
                        } else {
                            const float dy = sc[k] * (dyh - kb[k] - xh * kg[k]);
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap)
                                wacc[tap][k] = fmaf(dy, xp[tap / 3][j + tap % 3], wacc[tap][k]);
                        }
                    }
                }
            }
        }
    }
    if (MODE == FC_STATS || MODE == FC_BWD_REDUCE) {
        __shared__ float red[256 * 16];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            red[threadIdx.x * 16 + k] = acc0[k];
            red[threadIdx.x * 16 + 8 + k] = acc1[k];
        }
        __syncthreads();
        for (int j = threadIdx.x; j < CG * 16; j += blockDim.x) {
            const int g2 = j / 16, e = j % 16;
            float s = 0.f;
            for (int tt = g2; tt < 256; tt += CG) s += red[tt * 16 + e];
            A.partial[(long long)blockIdx.x * 2 * Co + (e < 8 ? 0 : Co) + g2 * 8 + (e & 7)] = s;
        }
    } else if (MODE == FC_BWD_WGRAD) {
        __shared__ float red[256 * 9];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            __syncthreads();
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) red[threadIdx.x * 9 + tap] = wacc[tap][k];
            __syncthreads();
            for (int j = threadIdx.x; j < CG * 9; j += blockDim.x) {
                const int g2 = j / 9, tap = j % 9;
                float s = 0.f;
                for (int tt = g2; tt < 256; tt += CG) s += red[tt * 9 + tap];
                A.wpartial[((long long)blockIdx.x * Co + g2 * 8 + k) * 9 + tap] = s;
            }
        }
    }
}

// out[co][ci_sel][tap] = sum over blocks of wpartial[b][co][tap]
static __global__ void first_wgrad_finalize_kernel(const float* __restrict__ wpartial, int blocks, int Co,
                                            int Ci, int ci_sel, float* __restrict__ dw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Co * 9) return;
    double s = 0.0;
    for (int b = 0; b < blocks; ++b) s += (double)wpartial[(long long)b * Co * 9 + i];
    const int co = i / 9, tap = i % 9;
    dw[((long long)co * Ci + ci_sel) * 9 + tap] = (float)s;
}

}  // namespace ub
