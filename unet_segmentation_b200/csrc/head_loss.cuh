// 1x1 output convolution (reference OutConv, models/unet_model.py:56-63) and the pixel-weighted
// cross-entropy (reference utils/losses.py:27,49,54,57): HBM-bound CUDA-core kernels.
#pragma once
#include "common.cuh"
#include "elementwise.cuh"

namespace ub {

constexpr int HEAD_MAX_CLASSES = 8;

// Fast path of the 1x1 head for K in {64, 128, 256}: K/8 lanes per pixel (one coalesced 16-byte
// load each, channel group fixed per thread so the weights live in registers), 4 pixels in flight per
// thread, partial dot products combined with xor-shuffles — the same arithmetic order as the head
// fused into the last BN-apply kernel (elementwise.cuh), so eval and training logits of identical
// activations agree bit for bit.
template <int NCT>
static __global__ void __launch_bounds__(256)
head_fwd_vec_kernel(const __nv_bfloat16* __restrict__ a, unsigned npix, unsigned HW, int K, int NC,
                    const float* __restrict__ w, const float* __restrict__ b,
                    float* __restrict__ logits, unsigned char* __restrict__ mask) {
    pdl_entry();
    const unsigned CG = (unsigned)K >> 3;
    const unsigned cg = threadIdx.x % CG;
    float wr[NCT][8];
#pragma unroll
    for (int c = 0; c < NCT; ++c)
#pragma unroll
        for (int k = 0; k < 8; ++k) wr[c][k] = c < NC ? w[c * K + cg * 8 + k] : 0.f;
    const unsigned gstride = gridDim.x * 256u / CG;
    const unsigned first = (blockIdx.x * 256u + threadIdx.x) / CG;
    const unsigned iters = (npix + 4u * gstride - 1u) / (4u * gstride);   // uniform: shuffles inside
    for (unsigned it = 0; it < iters; ++it) {
        const unsigned p0 = first + it * 4u * gstride;
        uint4 raw[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned p = p0 + j * gstride;
            if (p < npix) raw[j] = ldg16(a + (size_t)p * K + cg * 8);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned p = p0 + j * gstride;
            const bool ok = p < npix;
            float acc[NCT];
#pragma unroll
            for (int c = 0; c < NCT; ++c) acc[c] = 0.f;
            if (ok) {
                const Vec8 x = unpack8(raw[j]);
#pragma unroll
                for (int c = 0; c < NCT; ++c)
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[c] = fmaf(x.v[k], wr[c][k], acc[c]);
            }
#pragma unroll
            for (int c = 0; c < NCT; ++c) {
                if (c < NC) {
                    for (unsigned off = 1; off < CG; off <<= 1)
                        acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], off);
                }
            }
            if (ok && cg == 0) {
                const unsigned n = p / HW, hw = p % HW;
#pragma unroll
                for (int c = 0; c < NCT; ++c) {
                    acc[c] += (b && c < NC) ? b[c] : 0.f;
                    if (c < NC) logits[((size_t)n * NC + c) * HW + hw] = acc[c];
                }
                // 2+ classes: softmax[1] > 0.5 == z1 > z0 (scripts/predict.py:85-92); one class: the
                // stale callers' sigmoid(z) > 0.5 == z > 0 (scripts/inference.py:39,85)
                if (mask) mask[p] = (NC >= 2 ? (NCT >= 2 && acc[1] > acc[0]) : acc[0] > 0.f) ? 255 : 0;
            }
        }
    }
}

// Generic K (any multiple of 8): one thread per pixel, weights in shared memory.
// logits[n][c][h][w] (fp32 NCHW) = sum_k a[n][h][w][k] * w[c][k] + b[c];  optional u8 mask for the
// 2-class eval path: 255 where z1 > z0 (== softmax(z)[1] > 0.5, reference scripts/predict.py:85-92).
static __global__ void __launch_bounds__(256)
head_fwd_kernel(const __nv_bfloat16* __restrict__ a, long long P, long long HW, int K, int NC,
                const float* __restrict__ w, const float* __restrict__ b,
                float* __restrict__ logits, unsigned char* __restrict__ mask) {
    pdl_entry();
    extern __shared__ float wsm[];  // [NC][K] + [NC]
    for (int i = threadIdx.x; i < NC * K; i += blockDim.x) wsm[i] = w[i];
    for (int i = threadIdx.x; i < NC; i += blockDim.x) wsm[NC * K + i] = b ? b[i] : 0.f;
    __syncthreads();
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P;
         p += (long long)gridDim.x * blockDim.x) {
        float acc[HEAD_MAX_CLASSES];
#pragma unroll
        for (int c = 0; c < HEAD_MAX_CLASSES; ++c) acc[c] = c < NC ? wsm[NC * K + c] : 0.f;
        for (int k0 = 0; k0 < K; k0 += 8) {
            const Vec8 x = unpack8(ldg16(a + p * K + k0));
#pragma unroll
            for (int c = 0; c < HEAD_MAX_CLASSES; ++c) {
                if (c < NC) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[c] = fmaf(x.v[k], wsm[c * K + k0 + k], acc[c]);
                }
            }
        }
        const unsigned n = (unsigned)p / (unsigned)HW, hw = (unsigned)p % (unsigned)HW;
#pragma unroll
        for (int c = 0; c < HEAD_MAX_CLASSES; ++c)
            if (c < NC) logits[((size_t)n * NC + c) * HW + hw] = acc[c];
        if (mask) mask[p] = (NC >= 2 ? acc[1] > acc[0] : acc[0] > 0.f) ? 255 : 0;
    }
}

// Backward of the 1x1 head: da[p][k] = sum_c dl[c][p] * w[c][k] (bf16 out);
// partial[b][c][k] = sum_p dl[c][p]*a[p][k];  partial_b[b][c] = sum_p dl[c][p].
// One thread = one pixel x 8 channels, channel group fixed per thread.
template <int NCT>  // compile-time bound on the class count (register arrays sized by it)
static __global__ void __launch_bounds__(256)
head_bwd_kernel(const float* __restrict__ dlogits, const __nv_bfloat16* __restrict__ a, long long P,
                long long HW, int K, int NC, const float* __restrict__ w,
                __nv_bfloat16* __restrict__ da, float* __restrict__ partial) {
    pdl_entry();
    constexpr int HEAD_MAX_CLASSES = NCT;
    const int CG = K >> 3;
    const int cg = threadIdx.x % CG;
    float wr[NCT][8];
    float accw[NCT][8];
    float accb[NCT];
#pragma unroll
    for (int c = 0; c < NCT; ++c) {
        accb[c] = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            wr[c][k] = c < NC ? w[c * K + cg * 8 + k] : 0.f;
            accw[c][k] = 0.f;
        }
    }
    const unsigned gstride = gridDim.x * 256u / CG;
    for (unsigned p0 = (blockIdx.x * 256u + threadIdx.x) / CG; p0 < (unsigned)P; p0 += 4 * gstride) {
        uint4 raw[4];
        float dl[4][NCT];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned pu = p0 + j * gstride;
            if (pu < (unsigned)P) {
                raw[j] = ldg16(a + (size_t)pu * K + cg * 8);
                const size_t n = pu / (unsigned)HW, hw = pu % (unsigned)HW;
#pragma unroll
                for (int c = 0; c < NCT; ++c)
                    dl[j][c] = c < NC ? __ldg(dlogits + (n * NC + c) * HW + hw) : 0.f;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned pu = p0 + j * gstride;
            if (pu < (unsigned)P) {
                const Vec8 x = unpack8(raw[j]);
                Vec8 o;
#pragma unroll
                for (int k = 0; k < 8; ++k) o.v[k] = 0.f;
#pragma unroll
                for (int c = 0; c < NCT; ++c) {
                    accb[c] += dl[j][c];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        o.v[k] = fmaf(dl[j][c], wr[c][k], o.v[k]);
                        accw[c][k] = fmaf(dl[j][c], x.v[k], accw[c][k]);
                    }
                }
                *reinterpret_cast<uint4*>(da + (size_t)pu * K + cg * 8) = pack8(o);
            }
        }
    }
    // block reduce: partial[b][c][K] and partial bias at [b][NC*K + c]
    __shared__ float red[256 * 9];
    float* pb = partial + (long long)blockIdx.x * (NC * K + NC);
#pragma unroll
    for (int c = 0; c < HEAD_MAX_CLASSES; ++c) {
        if (c < NC) {
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 8; ++k) red[threadIdx.x * 9 + k] = accw[c][k];
            red[threadIdx.x * 9 + 8] = accb[c];
            __syncthreads();
            for (int j = threadIdx.x; j < CG * 8; j += blockDim.x) {
                const int g2 = j / 8, k = j % 8;
                float s = 0.f;
                for (int tt = g2; tt < 256; tt += CG) s += red[tt * 9 + k];
                pb[c * K + g2 * 8 + k] = s;
            }
            if (threadIdx.x == 0) {
                // every channel group saw every pixel: take group 0's copies only
                float s = 0.f;
                for (int tt = 0; tt < 256; tt += CG) s += red[tt * 9 + 8];
                pb[NC * K + c] = s;
            }
        }
    }
}

// out[j] = sum_b partial[b][j]  (fixed order). blockDim = (32 elements, 32 slices), see
// finalize_combine.
static __global__ void __launch_bounds__(1024)
reduce_partials_kernel(const float* __restrict__ partial, int blocks, int len,
                       float* __restrict__ out0, int len0, float* __restrict__ out1) {
    pdl_entry();
    const int j = blockIdx.x * 32 + threadIdx.x;
    double s = 0.0, unused = 0.0;
    if (j < len)
        for (int b = threadIdx.y; b < blocks; b += FIN_SLICES) s += (double)partial[(long long)b * len + j];
    finalize_combine(s, unused);
    if (threadIdx.y != 0 || j >= len) return;
    if (j < len0) out0[j] = (float)s;
    else if (out1) out1[j - len0] = (float)s;
}

// ---------------------------------------------------------------------------------------------
// Weighted cross-entropy, forward + gradient in one pass.
//   loss = mean_{n,h,w} w * (logsumexp_c z - z_t);   dz_c = w * (softmax_c - [c==t]) / count
// All three inputs may be arbitrarily strided (the reference hands over centre-cropped,
// squeezed views: scripts/train.py:118-126). ignore_index (-100) pixels contribute 0 loss / grad
// but still count in the mean (reduction='none' followed by .mean()).
// ---------------------------------------------------------------------------------------------
struct WceArgs {
    const float* z; long long zN, zC, zH, zW;
    const long long* t; long long tN, tH, tW;
    const float* wm; long long wN, wH, wW;
    int N, C, H, W;
    float* dz;        // contiguous [N][C][H][W] or null
    float* partial;   // [gridDim.x]
    int* err;         // set to 1 when a target is out of range
};

static __global__ void __launch_bounds__(256)
wce_fwd_bwd_kernel(const WceArgs A) {
    pdl_entry();
    const long long HW = (long long)A.H * A.W;
    const long long P = (long long)A.N * HW;
    const float inv_count = 1.f / (float)P;
    float local = 0.f;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P;
         p += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(p / HW);
        const int hw = (int)(p % HW);
        const int h = hw / A.W, w = hw % A.W;
        const float* zp = A.z + n * A.zN + h * A.zH + w * A.zW;
        const long long tgt = A.t[n * A.tN + h * A.tH + w * A.tW];
        const float wt = A.wm[n * A.wN + h * A.wH + w * A.wW];
        const bool ignored = tgt == -100;
        if (!ignored && (tgt < 0 || tgt >= A.C)) {
            // The reference's nn.CrossEntropyLoss faults here (device-side assert). This path raises
            // the error flag (read by the Python module), poisons the loss with NaN so that the step
            // cannot pass for a valid one, and writes a defined (zero) gradient for the pixel.
            *A.err = 1;
            local = __int_as_float(0x7fc00000);
            if (A.dz)
                for (int c = 0; c < A.C; ++c) A.dz[((long long)n * A.C + c) * HW + hw] = 0.f;
            continue;
        }
        float loss, zt = 0.f;
        if (A.C == 2) {
            const float z0 = zp[0], z1 = zp[A.zC];
            const float m = fmaxf(z0, z1);
            const float e0 = expf(z0 - m), e1 = expf(z1 - m);
            const float se = e0 + e1;
            const float lse = m + logf(se);
            zt = tgt == 1 ? z1 : z0;
            loss = ignored ? 0.f : wt * (lse - zt);
            if (A.dz) {
                const float g = ignored ? 0.f : wt * inv_count;
                const float inv = 1.f / se;
                float* d = A.dz + (long long)n * 2 * HW + hw;
                d[0] = g * (e0 * inv - (tgt == 0 ? 1.f : 0.f));
                d[HW] = g * (e1 * inv - (tgt == 1 ? 1.f : 0.f));
            }
        } else {
            float m = -INFINITY;
            for (int c = 0; c < A.C; ++c) m = fmaxf(m, zp[c * A.zC]);
            float se = 0.f;
            for (int c = 0; c < A.C; ++c) {
                const float zc = zp[c * A.zC];
                se += expf(zc - m);
                if (c == tgt) zt = zc;
            }
            const float lse = m + logf(se);
            loss = ignored ? 0.f : wt * (lse - zt);
            if (A.dz) {
                const float g = ignored ? 0.f : wt * inv_count;
                const float inv = 1.f / se;
                for (int c = 0; c < A.C; ++c)
                    A.dz[((long long)n * A.C + c) * HW + hw] =
                        g * (expf(zp[c * A.zC] - m) * inv - (c == tgt ? 1.f : 0.f));
            }
        }
        local += loss;
    }
    // warp shuffle -> block -> one partial per block
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
    __shared__ float wsum[8];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += wsum[i];
        A.partial[blockIdx.x] = s;
    }
}

static __global__ void wce_finalize_kernel(const float* __restrict__ partial, int blocks, double count,
                                    float* __restrict__ loss) {
    pdl_entry();
    __shared__ double sm[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < blocks; i += blockDim.x) s += (double)partial[i];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off >= 1; off >>= 1) {
        if ((int)threadIdx.x < off) sm[threadIdx.x] += sm[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss = (float)(sm[0] / count);
}

// out = in * (*scalar)   (loss backward: grad_output is a device scalar)
static __global__ void scale_by_scalar_kernel(const float* __restrict__ in, const float* __restrict__ s,
                                       float* __restrict__ out, long long n) {
    pdl_entry();
    const float k = *s;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        out[i] = in[i] * k;
}

}  // namespace ub
