// HBM-bound kernels between the tensor-core convolutions: batch-norm statistics finalisation,
// BN-apply + ReLU (+ 2x2 max-pool), their backward (with the max-pool scatter and the
// skip-connection gradient gathered on the fly), weight packing, and the split-K reduction of
// weight gradients. All activations are NHWC bf16 with C % 8 == 0; one thread moves 16 bytes.
//
// Reference ops replaced: nn.BatchNorm2d train/eval (models/unet_model.py:12,16),
// nn.ReLU (:13,17), nn.MaxPool2d(2) (:28), _center_crop + torch.cat backward (:88-102,131-143).
#pragma once
#include "common.cuh"

namespace ub {

struct Vec8 {
    float v[8];
};
__device__ __forceinline__ Vec8 unpack8(const uint4& u) {
    Vec8 r;
    r.v[0] = bf16_lo(u.x); r.v[1] = bf16_hi(u.x);
    r.v[2] = bf16_lo(u.y); r.v[3] = bf16_hi(u.y);
    r.v[4] = bf16_lo(u.z); r.v[5] = bf16_hi(u.z);
    r.v[6] = bf16_lo(u.w); r.v[7] = bf16_hi(u.w);
    return r;
}
__device__ __forceinline__ uint4 pack8(const Vec8& r) {
    uint4 u;
    u.x = pack_bf16x2(r.v[0], r.v[1]);
    u.y = pack_bf16x2(r.v[2], r.v[3]);
    u.z = pack_bf16x2(r.v[4], r.v[5]);
    u.w = pack_bf16x2(r.v[6], r.v[7]);
    return u;
}
__device__ __forceinline__ uint4 ldg16(const void* p) {
    return __ldg(reinterpret_cast<const uint4*>(p));
}
// bf16 rounding of relu(y*scale+shift): the exact arithmetic of the forward pass, so that the
// backward pass can recompute activations (ReLU mask, pool arg-max) bit-identically.
__device__ __forceinline__ float bn_relu_bf16(float y, float sc, float sh) {
    return __bfloat162float(__float2bfloat16_rn(fmaxf(fmaf(y, sc, sh), 0.f)));
}

// ---------------------------------------------------------------------------------------------
// BN statistics finalisation (train mode). Partials come from the conv epilogue:
// stats[cta][2][BN] (one row per CTA, its epilogue warps pre-combined); CTA b covers channel tile
// (b % n_tiles).
// Biased variance normalises, unbiased variance updates running_var (torch semantics).
// ---------------------------------------------------------------------------------------------
// All finalisation kernels use blockDim = (32 channels, 32 slices): the per-CTA partials are summed
// by 32 threads per channel in double precision and combined through shared memory in a fixed
// order (deterministic), instead of one thread walking hundreds of partial rows serially (each
// dependent load -> add step costs ~0.35 us, so the rows per thread are what these kernels cost).
constexpr int FIN_SLICES = 32;
__device__ __forceinline__ void finalize_combine(double& s, double& q) {
    __shared__ double sm[2][FIN_SLICES][32];
    sm[0][threadIdx.y][threadIdx.x] = s;
    sm[1][threadIdx.y][threadIdx.x] = q;
    __syncthreads();
    if (threadIdx.y == 0) {
        s = 0.0; q = 0.0;
        for (int k = 0; k < FIN_SLICES; ++k) { s += sm[0][k][threadIdx.x]; q += sm[1][k][threadIdx.x]; }
    }
}
__device__ __forceinline__ void bn_finalize_write(int c, double s, double q, double count,
                                                  const float* gamma, const float* beta,
                                                  float* running_mean, float* running_var,
                                                  float momentum, float eps, float* scale,
                                                  float* shift, float* save_mean, float* save_rstd) {
    const double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * rstd;
    scale[c] = sc;
    shift[c] = beta[c] - (float)mean * sc;
    save_mean[c] = (float)mean;
    save_rstd[c] = rstd;
    if (running_mean) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}
static __global__ void __launch_bounds__(1024)
bn_finalize_kernel(const float* __restrict__ stats, int grid_ctas, int n_tiles, int BN, int C,
                   double count, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* running_mean, float* running_var, long long* num_batches_tracked,
                   float momentum, float eps, float* __restrict__ scale, float* __restrict__ shift,
                   float* __restrict__ save_mean, float* __restrict__ save_rstd) {
    pdl_entry();
    const int c = blockIdx.x * 32 + threadIdx.x;
    if (c == 0 && threadIdx.y == 0 && num_batches_tracked) *num_batches_tracked += 1;
    double s = 0.0, q = 0.0;
    if (c < C) {
        const int tile = c / BN, col = c % BN;
        const int rows = (grid_ctas - tile + n_tiles - 1) / n_tiles;  // CTAs of this channel tile
        for (int r = threadIdx.y; r < rows; r += FIN_SLICES) {
            const float* p = stats + (long long)(tile + r * n_tiles) * (2 * BN);
            s += (double)p[col];
            q += (double)p[BN + col];
        }
    }
    finalize_combine(s, q);
    if (threadIdx.y == 0 && c < C)
        bn_finalize_write(c, s, q, count, gamma, beta, running_mean, running_var, momentum, eps,
                          scale, shift, save_mean, save_rstd);
}

// Generic variant: partials[blocks][2][C] (first-layer statistics).
static __global__ void __launch_bounds__(1024)
bn_finalize_flat_kernel(const float* __restrict__ part, int blocks, int C, double count,
                        const float* __restrict__ gamma, const float* __restrict__ beta,
                        float* running_mean, float* running_var, long long* num_batches_tracked,
                        float momentum, float eps, float* __restrict__ scale,
                        float* __restrict__ shift, float* __restrict__ save_mean,
                        float* __restrict__ save_rstd) {
    pdl_entry();
    const int c = blockIdx.x * 32 + threadIdx.x;
    if (c == 0 && threadIdx.y == 0 && num_batches_tracked) *num_batches_tracked += 1;
    double s = 0.0, q = 0.0;
    if (c < C) {
        for (int b = threadIdx.y; b < blocks; b += FIN_SLICES) {
            s += (double)part[(long long)b * 2 * C + c];
            q += (double)part[(long long)b * 2 * C + C + c];
        }
    }
    finalize_combine(s, q);
    if (threadIdx.y == 0 && c < C)
        bn_finalize_write(c, s, q, count, gamma, beta, running_mean, running_var, momentum, eps,
                          scale, shift, save_mean, save_rstd);
}

// Eval mode: fold conv bias + running statistics into a per-channel affine.
static __global__ void bn_fold_eval_kernel(int C, const float* __restrict__ conv_bias,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ rm, const float* __restrict__ rv,
                                    float eps, float* __restrict__ scale,
                                    float* __restrict__ shift) {
    pdl_entry();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float sc = gamma[c] / sqrtf(rv[c] + eps);
    scale[c] = sc;
    shift[c] = beta[c] + ((conv_bias ? conv_bias[c] : 0.f) - rm[c]) * sc;
}

// ---------------------------------------------------------------------------------------------
// BN-apply + ReLU (+ fused 2x2/2 floor-mode max-pool).   y -> a (and pooled p)
// ---------------------------------------------------------------------------------------------
// Index arithmetic is 32-bit (hosts reject tensors with >= 2^31 16-byte items) and the channel
// group of a thread is fixed (256 % (C/8) == 0), so the per-channel affine lives in registers and the
// grid-stride loop needs no division; 4 independent 16-byte loads are in flight per thread.
template <bool POOL>
static __global__ void __launch_bounds__(256)
bn_apply_relu_kernel(const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ a,
                     __nv_bfloat16* __restrict__ pooled, unsigned char* __restrict__ amax, int N,
                     int H, int W, int C, const float* __restrict__ scale,
                     const float* __restrict__ shift) {
    pdl_entry();
    const unsigned CG = (unsigned)C >> 3;
    const unsigned cg = threadIdx.x % CG;
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { sc[k] = scale[cg * 8 + k]; sh[k] = shift[cg * 8 + k]; }
    const unsigned gstride = gridDim.x * 256u / CG;
    const unsigned first = (blockIdx.x * 256u + threadIdx.x) / CG;
    if (!POOL) {
        const unsigned npix = (unsigned)N * H * W;
        for (unsigned p0 = first; p0 < npix; p0 += 4 * gstride) {
            uint4 raw[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned p = p0 + j * gstride;
                if (p < npix) raw[j] = ldg16(y + (size_t)p * C + cg * 8);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned p = p0 + j * gstride;
                if (p < npix) {
                    const Vec8 x = unpack8(raw[j]);
                    Vec8 o;
#pragma unroll
                    for (int k = 0; k < 8; ++k) o.v[k] = fmaxf(fmaf(x.v[k], sc[k], sh[k]), 0.f);
                    *reinterpret_cast<uint4*>(a + (size_t)p * C + cg * 8) = pack8(o);
                }
            }
        }
    } else {
        const unsigned HW2 = (H + 1) >> 1, WW2 = (W + 1) >> 1, Hp = H >> 1, Wp = W >> 1;
        const unsigned nwin = (unsigned)N * HW2 * WW2;
        for (unsigned wi = first; wi < nwin; wi += gstride) {
            const unsigned wp = wi % WW2, t = wi / WW2, hp = t % HW2, n = t / HW2;
            uint4 raw[4];
            bool inb[4];
            size_t off[4];
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const unsigned h = hp * 2 + (d >> 1), w = wp * 2 + (d & 1);
                inb[d] = h < (unsigned)H && w < (unsigned)W;
                off[d] = ((size_t)(n * H + h) * W + w) * C + cg * 8;
                if (inb[d]) raw[d] = ldg16(y + off[d]);
            }
            Vec8 mx;
            unsigned am[8];   // first arg-max of the window in row-major order (torch tie rule)
#pragma unroll
            for (int k = 0; k < 8; ++k) { mx.v[k] = -1.f; am[k] = 0u; }  // post-ReLU values are >= 0
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                if (inb[d]) {
                    const Vec8 x = unpack8(raw[d]);
                    Vec8 o;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        o.v[k] = bn_relu_bf16(x.v[k], sc[k], sh[k]);
                        if (o.v[k] > mx.v[k]) { mx.v[k] = o.v[k]; am[k] = (unsigned)d; }
                    }
                    *reinterpret_cast<uint4*>(a + off[d]) = pack8(o);
                }
            }
            if (hp < Hp && wp < Wp) {
                const size_t po = ((size_t)(n * Hp + hp) * Wp + wp) * C + cg * 8;
                *reinterpret_cast<uint4*>(pooled + po) = pack8(mx);
                if (amax) {
                    uint2 pk;
                    pk.x = am[0] | (am[1] << 8) | (am[2] << 16) | (am[3] << 24);
                    pk.y = am[4] | (am[5] << 8) | (am[6] << 16) | (am[7] << 24);
                    *reinterpret_cast<uint2*>(amax + po) = pk;
                }
            }
        }
    }
}

// BN-apply + ReLU of the LAST conv unit fused with the 1x1 output convolution (reference OutConv,
// models/unet_model.py:56-63, :145): the logits are dotted from the activation values while they
// are still in registers (as rounded to bf16, i.e. exactly what is stored), so the separate head
// kernel does not re-read the 128 B/pixel activation. The C/8 lanes that share a pixel combine
// their partial dot products with xor-shuffles; lane 0 of the group writes the NCHW fp32 logits.
template <int NCT>
static __global__ void __launch_bounds__(256)
bn_apply_relu_head_kernel(const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ a,
                          unsigned npix, unsigned HW, int C, const float* __restrict__ scale,
                          const float* __restrict__ shift, int NC, const float* __restrict__ hw_,
                          const float* __restrict__ hb, float* __restrict__ logits) {
    pdl_entry();
    const unsigned CG = (unsigned)C >> 3;   // 8 or 16 or 32: a pixel's lanes are one aligned group
    const unsigned cg = threadIdx.x % CG;
    float sc[8], sh[8], wr[NCT][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { sc[k] = scale[cg * 8 + k]; sh[k] = shift[cg * 8 + k]; }
#pragma unroll
    for (int c = 0; c < NCT; ++c)
#pragma unroll
        for (int k = 0; k < 8; ++k) wr[c][k] = c < NC ? hw_[c * C + cg * 8 + k] : 0.f;
    const unsigned gstride = gridDim.x * 256u / CG;
    const unsigned first = (blockIdx.x * 256u + threadIdx.x) / CG;
    const unsigned iters = (npix + 4u * gstride - 1u) / (4u * gstride);   // uniform: shuffles inside
    for (unsigned it = 0; it < iters; ++it) {
        const unsigned p0 = first + it * 4u * gstride;
        uint4 raw[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned p = p0 + j * gstride;
            if (p < npix) raw[j] = ldg16(y + (size_t)p * C + cg * 8);
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
                Vec8 o;
#pragma unroll
                for (int k = 0; k < 8; ++k) o.v[k] = bn_relu_bf16(x.v[k], sc[k], sh[k]);
                *reinterpret_cast<uint4*>(a + (size_t)p * C + cg * 8) = pack8(o);
#pragma unroll
                for (int c = 0; c < NCT; ++c)
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[c] = fmaf(o.v[k], wr[c][k], acc[c]);
            }
#pragma unroll
            for (int c = 0; c < NCT; ++c) {
                if (c < NC) {   // NC is uniform
                    for (unsigned off = 1; off < CG; off <<= 1)
                        acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], off);
                }
            }
            if (ok && cg == 0) {
                const unsigned n = p / HW, hw = p % HW;
#pragma unroll
                for (int c = 0; c < NCT; ++c)
                    if (c < NC) logits[((size_t)n * NC + c) * HW + hw] = acc[c] + (hb ? hb[c] : 0.f);
            }
        }
    }
}

// Stand-alone 2x2 max-pool (eval path, where BN+ReLU is folded into the conv epilogue).
static __global__ void __launch_bounds__(256)
maxpool2_kernel(const __nv_bfloat16* __restrict__ a, __nv_bfloat16* __restrict__ pooled, int N,
                int H, int W, int C) {
    pdl_entry();
    const unsigned CG = (unsigned)C >> 3, Hp = H >> 1, Wp = W >> 1;
    const unsigned total = (unsigned)N * Hp * Wp * CG;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned cg = i % CG;
        unsigned t = i / CG;
        const unsigned wp = t % Wp; t /= Wp;
        const unsigned hp = t % Hp, n = t / Hp;
        uint4 raw[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const unsigned h = hp * 2 + (d >> 1), w = wp * 2 + (d & 1);
            raw[d] = ldg16(a + ((size_t)(n * H + h) * W + w) * C + cg * 8);
        }
        Vec8 mx = unpack8(raw[0]);
#pragma unroll
        for (int d = 1; d < 4; ++d) {
            const Vec8 x = unpack8(raw[d]);
#pragma unroll
            for (int k = 0; k < 8; ++k) mx.v[k] = fmaxf(mx.v[k], x.v[k]);
        }
        *reinterpret_cast<uint4*>(pooled + (size_t)i * 8) = pack8(mx);
    }
}

// ---------------------------------------------------------------------------------------------
// BN + ReLU backward. Upstream gradient of the post-ReLU activation a[N,H,W,C] is either
//   DIRECT: a tensor view g (possibly a channel slice of a wider buffer), or
//   POOL_SKIP: gathered on the fly from the gradient of the pooled tensor (routed to the first
//     arg-max of each 2x2 window, recomputed from y) plus the gradient of the centre-cropped skip
//     connection (a channel slice of d(concat)), so d(a) is never materialised.
// Pass 1 (reduce): per-block partial sums of dyh = g*[a>0] and dyh*xhat.
// Pass 2 (apply):  dy = scale*(dyh - mean(dyh) - xhat*mean(dyh*xhat)), written as bf16.
// ---------------------------------------------------------------------------------------------
struct BnBwdArgs {
    const __nv_bfloat16* y;  // pre-BN conv output [N,H,W,C]
    int N, H, W, C;
    const float* scale;      // gamma*rstd
    const float* shift;      // beta - mean*scale
    const float* mean;
    const float* rstd;
    View g;                  // DIRECT
    View gp;                 // POOL_SKIP: grad of pooled [N,H/2,W/2,C]
    View gs;                 // POOL_SKIP: grad of cropped skip [N,th,tw,C] (may be a slice)
    int crop_h, crop_w;
    int has_skip;
    const unsigned char* amax;  // POOL_SKIP: arg-max (0..3) per pooled element, saved by the forward
    float* partial;          // reduce: [gridDim.x][2][C]
    const float* dgamma;     // apply
    const float* dbeta;
    float inv_count;
    __nv_bfloat16* dy;       // apply: [N,H,W,C]
};

// PIX (POOL_SKIP only): the forward saved the pool arg-max byte, so the window's winner is decoded
// instead of being recomputed from y (one thread = one 2x2 window x 8 channels in both variants).
// Launch bounds: the direct variant is held to 85 registers (3 CTAs per SM) so that two of its CTAs
// still fit beside a weight-gradient CTA (256 threads x 74 registers) when the two kernels overlap.
template <bool POOL_SKIP, bool APPLY, bool PIX = false>
static __global__ void __launch_bounds__(256, POOL_SKIP ? 2 : 3)
bn_bwd_kernel(const BnBwdArgs A) {
    pdl_entry();
    const unsigned C = A.C, CG = C >> 3, H = A.H, W = A.W;
    const unsigned cg = threadIdx.x % CG;  // host guarantees 256 % CG == 0
    // Per-channel constants folded so that the inner loop is 5-6 instructions per element:
    //   reduce: dyh = g*[y*sc+sh > 0];  sum dyh;  sum dyh*(y - mean)   (rstd applied by the finalisation)
    //   apply:  dy = sc*(dyh - dbeta/n - xhat*dgamma/n) = sc*dyh - (c1 + c2*y),
    //           c2 = sc*rstd*dgamma/n,  c1 = sc*dbeta/n - c2*mean
    float sc[8], sh[8], mu[8], c1[8], c2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        sc[k] = A.scale[cg * 8 + k];
        sh[k] = A.shift[cg * 8 + k];
        mu[k] = A.mean[cg * 8 + k];
        if (APPLY) {
            c2[k] = sc[k] * A.rstd[cg * 8 + k] * (A.dgamma[cg * 8 + k] * A.inv_count);
            c1[k] = sc[k] * (A.dbeta[cg * 8 + k] * A.inv_count) - c2[k] * mu[k];
        }
    }
    float acc_b[8], acc_g[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc_b[k] = 0.f; acc_g[k] = 0.f; }

    const unsigned gstride = gridDim.x * 256u / CG;
    const unsigned first = (blockIdx.x * 256u + threadIdx.x) / CG;

    auto process = [&](size_t pix, const Vec8& yv, const Vec8& gv) {
        Vec8 o;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float act = fmaf(yv.v[k], sc[k], sh[k]);
            const float dyh = act > 0.f ? gv.v[k] : 0.f;
            if (APPLY) {
                o.v[k] = fmaf(sc[k], dyh, -fmaf(c2[k], yv.v[k], c1[k]));
            } else {
                acc_b[k] += dyh;
                acc_g[k] = fmaf(dyh, yv.v[k] - mu[k], acc_g[k]);
            }
        }
        if (APPLY) *reinterpret_cast<uint4*>(A.dy + pix * C + cg * 8) = pack8(o);
    };

    if (POOL_SKIP && PIX) {
        // One thread = one 2x2 pooling window x 8 channels (arg-max saved by the forward pass): the
        // pooled gradient and the arg-max byte are loaded and decoded ONCE per window instead of once
        // per pixel, and the window coordinates advance incrementally (no division in the loop).
        const unsigned Hp = H >> 1, Wp = W >> 1;
        const unsigned HW2 = (H + 1) >> 1, WW2 = (W + 1) >> 1;   // windows incl. the odd last row / col
        const unsigned nwin = (unsigned)A.N * HW2 * WW2;
        const __nv_bfloat16* gpb = reinterpret_cast<const __nv_bfloat16*>(A.gp.ptr);
        const __nv_bfloat16* gsb = reinterpret_cast<const __nv_bfloat16*>(A.gs.ptr);
        const unsigned dWq = gstride % WW2, dHq = (gstride / WW2) % HW2, dN = (gstride / WW2) / HW2;
        unsigned wq = first % WW2, hq = (first / WW2) % HW2, n = (first / WW2) / HW2;
        for (unsigned wi = first; wi < nwin; wi += gstride) {
            uint4 yr[4], gsr[4], gpr;
            uint2 amr;
            bool inb[4], ins[4];
            size_t pix[4];
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const unsigned h = hq * 2 + (d >> 1), w = wq * 2 + (d & 1);
                inb[d] = h < H && w < W;
                pix[d] = (size_t)(n * H + h) * W + w;
                if (inb[d]) yr[d] = ldg16(A.y + pix[d] * C + cg * 8);
                const int hs = (int)h - A.crop_h, ws = (int)w - A.crop_w;
                ins[d] = inb[d] && A.has_skip && hs >= 0 && hs < A.gs.H && ws >= 0 && ws < A.gs.W;
                if (ins[d])
                    gsr[d] = ldg16(gsb + (size_t)(n * A.gs.sN + hs * A.gs.sH + ws * A.gs.sW) + cg * 8);
            }
            const bool full = hq < Hp && wq < Wp;
            if (full) {
                gpr = ldg16(gpb + (size_t)(n * A.gp.sN + hq * A.gp.sH + wq * A.gp.sW) + cg * 8);
                amr = __ldg(reinterpret_cast<const uint2*>(
                    A.amax + ((size_t)(n * Hp + hq) * Wp + wq) * C + cg * 8));
            }
            Vec8 gpv;
            unsigned a8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { gpv.v[k] = 0.f; a8[k] = 0xFFu; }
            if (full) {
                gpv = unpack8(gpr);
#pragma unroll
                for (int k = 0; k < 8; ++k) a8[k] = ((k < 4 ? amr.x : amr.y) >> (8 * (k & 3))) & 0xFFu;
            }
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                if (!inb[d]) continue;
                Vec8 gv;
                if (ins[d]) {
                    gv = unpack8(gsr[d]);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) gv.v[k] = 0.f;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) gv.v[k] += (a8[k] == (unsigned)d) ? gpv.v[k] : 0.f;
                process(pix[d], unpack8(yr[d]), gv);
            }
            // advance to window wi + gstride
            wq += dWq;
            const unsigned cw = wq >= WW2 ? 1u : 0u;
            wq -= cw ? WW2 : 0u;
            hq += dHq + cw;
            const unsigned chh = hq >= HW2 ? 1u : 0u;
            hq -= chh ? HW2 : 0u;
            n += dN + chh;
        }
    } else if (!POOL_SKIP) {
        const unsigned npix = (unsigned)A.N * H * W;
        const __nv_bfloat16* gb = reinterpret_cast<const __nv_bfloat16*>(A.g.ptr);
        const bool glin = A.g.sW == (long long)C && A.g.sH == (long long)W * C &&
                          A.g.sN == (long long)H * W * C;
        for (unsigned p0 = first; p0 < npix; p0 += 4 * gstride) {
            uint4 yr[4], gr[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned p = p0 + j * gstride;
                if (p < npix) {
                    yr[j] = ldg16(A.y + (size_t)p * C + cg * 8);
                    size_t goff;
                    if (glin) {
                        goff = (size_t)p * C;
                    } else {
                        const unsigned w = p % W, t = p / W, h = t % H, n = t / H;
                        goff = (size_t)(n * A.g.sN + h * A.g.sH + w * A.g.sW);
                    }
                    gr[j] = ldg16(gb + goff + cg * 8);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned p = p0 + j * gstride;
                if (p < npix) process((size_t)p, unpack8(yr[j]), unpack8(gr[j]));
            }
        }
    } else {
        const unsigned HW2 = (H + 1) >> 1, WW2 = (W + 1) >> 1;
        const unsigned nwin = (unsigned)A.N * HW2 * WW2;
        const __nv_bfloat16* gpb = reinterpret_cast<const __nv_bfloat16*>(A.gp.ptr);
        const __nv_bfloat16* gsb = reinterpret_cast<const __nv_bfloat16*>(A.gs.ptr);
        for (unsigned wi = first; wi < nwin; wi += gstride) {
            const unsigned wq = wi % WW2, t = wi / WW2, hq = t % HW2, n = t / HW2;
            uint4 yr[4], gsr[4], gpr;
            bool inb[4], ins[4];
            size_t pix[4];
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const unsigned h = hq * 2 + (d >> 1), w = wq * 2 + (d & 1);
                inb[d] = h < H && w < W;
                pix[d] = (size_t)(n * H + h) * W + w;
                if (inb[d]) yr[d] = ldg16(A.y + pix[d] * C + cg * 8);
                const int hs = (int)h - A.crop_h, ws = (int)w - A.crop_w;
                ins[d] = inb[d] && A.has_skip && hs >= 0 && hs < A.gs.H && ws >= 0 && ws < A.gs.W;
                if (ins[d])
                    gsr[d] = ldg16(gsb + (size_t)(n * A.gs.sN + hs * A.gs.sH + ws * A.gs.sW) + cg * 8);
            }
            const bool full = (hq < (H >> 1)) && (wq < (W >> 1));
            if (full)
                gpr = ldg16(gpb + (size_t)(n * A.gp.sN + hq * A.gp.sH + wq * A.gp.sW) + cg * 8);
            Vec8 mx;
#pragma unroll
            for (int k = 0; k < 8; ++k) mx.v[k] = 0.f;
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                if (inb[d]) {
                    const Vec8 x = unpack8(yr[d]);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        mx.v[k] = fmaxf(mx.v[k], bn_relu_bf16(x.v[k], sc[k], sh[k]));
                }
            }
            Vec8 gpv;
            if (full) gpv = unpack8(gpr);
            unsigned taken = 0u;
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                if (!inb[d]) continue;
                const Vec8 x = unpack8(yr[d]);
                Vec8 gv;
                if (ins[d]) {
                    gv = unpack8(gsr[d]);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) gv.v[k] = 0.f;
                }
                if (full) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float av = bn_relu_bf16(x.v[k], sc[k], sh[k]);
                        if (!((taken >> k) & 1u) && av == mx.v[k]) {
                            gv.v[k] += gpv.v[k];
                            taken |= 1u << k;
                        }
                    }
                }
                process(pix[d], x, gv);
            }
        }
    }

    if (!APPLY) {
        // block reduction over the 256/CG threads that share a channel group
        __shared__ float red[256 * 16];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            red[threadIdx.x * 16 + k] = acc_b[k];
            red[threadIdx.x * 16 + 8 + k] = acc_g[k];
        }
        __syncthreads();
        for (unsigned j = threadIdx.x; j < CG * 16; j += blockDim.x) {
            const unsigned g2 = j / 16, e = j % 16;
            float s = 0.f;
            for (unsigned tt = g2; tt < 256; tt += CG) s += red[tt * 16 + e];
            const unsigned c = g2 * 8 + (e & 7);
            A.partial[(size_t)blockIdx.x * 2 * C + (e < 8 ? 0 : C) + c] = s;
        }
    }
}

// dbeta = sum over blocks of the first partial, dgamma = rstd * (sum of the second) — the reduce pass
// accumulates dyh*(y - mean); `rstd` null = the partials already hold dyh*xhat (first-conv path).
// Fixed order => deterministic.
static __global__ void __launch_bounds__(1024)
bn_bwd_finalize_kernel(const float* __restrict__ part, int blocks, int C,
                       const float* __restrict__ rstd, float* __restrict__ dgamma,
                       float* __restrict__ dbeta) {
    pdl_entry();
    const int c = blockIdx.x * 32 + threadIdx.x;
    double b = 0.0, g = 0.0;
    if (c < C) {
        for (int k = threadIdx.y; k < blocks; k += FIN_SLICES) {
            b += (double)part[(long long)k * 2 * C + c];
            g += (double)part[(long long)k * 2 * C + C + c];
        }
    }
    finalize_combine(b, g);
    if (threadIdx.y == 0 && c < C) {
        dbeta[c] = (float)b;
        dgamma[c] = (float)(rstd ? g * (double)rstd[c] : g);
    }
}

// Same, for partial rows written by a tensor-core epilogue (EPI_STORE_BNRED): stats[cta][2][BN], CTA b
// covers channel tile (b % n_tiles) — the layout bn_finalize_kernel reads for the forward statistics.
static __global__ void __launch_bounds__(1024)
bn_bwd_finalize_tiled_kernel(const float* __restrict__ stats, int grid_ctas, int n_tiles, int BN, int C,
                             const float* __restrict__ rstd, float* __restrict__ dgamma,
                             float* __restrict__ dbeta) {
    pdl_entry();
    const int c = blockIdx.x * 32 + threadIdx.x;
    double b = 0.0, g = 0.0;
    if (c < C) {
        const int tile = c / BN, col = c % BN;
        const int rows = (grid_ctas - tile + n_tiles - 1) / n_tiles;
        for (int r = threadIdx.y; r < rows; r += FIN_SLICES) {
            const float* p = stats + (long long)(tile + r * n_tiles) * (2 * BN);
            b += (double)p[col];
            g += (double)p[BN + col];
        }
    }
    finalize_combine(b, g);
    if (threadIdx.y == 0 && c < C) {
        dbeta[c] = (float)b;
        dgamma[c] = (float)(g * (double)rstd[c]);
    }
}

// ---------------------------------------------------------------------------------------------
// Weight packing (fp32 torch layouts -> bf16 GEMM operands)
// ---------------------------------------------------------------------------------------------
// All GEMM B operands are stored tap-major, [tap][N][C]: K index = (tap, channel).
// conv [Co][Ci][3][3] -> fprop B [tap][Co][Ci]   and   dgrad B [tap'][Ci][Co], tap' = 8 - tap
static __global__ void pack_conv3x3_kernel(const float* __restrict__ w, int Co, int Ci,
                                    __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd) {
    pdl_entry();
    const long long total = (long long)Co * Ci * 9;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        // i indexes the fprop layout (coalesced writes)
        const int ci = (int)(i % Ci);
        const int co = (int)((i / Ci) % Co);
        const int tap = (int)(i / ((long long)Ci * Co));
        const __nv_bfloat16 v = __float2bfloat16_rn(w[((long long)co * Ci + ci) * 9 + tap]);
        wf[i] = v;
        if (wd) wd[((long long)(8 - tap) * Ci + ci) * Co + co] = v;
    }
}
// convT [Ci][Co][2][2] -> fwd B [(q*Co+co)][Ci] (one tap, N = 4*Co)  and  bwd-data B [q][Ci][Co]
static __global__ void pack_convT2x2_kernel(const float* __restrict__ w, int Ci, int Co,
                                     __nv_bfloat16* __restrict__ wf,
                                     __nv_bfloat16* __restrict__ wb) {
    pdl_entry();
    const long long total = (long long)Ci * Co * 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Ci);
        const int co = (int)((i / Ci) % Co);
        const int q = (int)(i / ((long long)Ci * Co));
        const __nv_bfloat16 v = __float2bfloat16_rn(w[((long long)ci * Co + co) * 4 + q]);
        wf[i] = v;
        if (wb) wb[((long long)q * Ci + ci) * Co + co] = v;
    }
}
// bias of the transposed conv replicated over the 4 sub-pixel positions (GEMM column order)
static __global__ void tile_bias4_kernel(const float* __restrict__ b, int Co, float* __restrict__ out) {
    pdl_entry();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 4 * Co) out[i] = b[i % Co];
}

// Split-K reduction of weight-gradient partial tiles + permutation into the torch layout
// (cols % 32 == 0, RC % 8 == 0):
//   ws[split][row = tap*RC + rc][col]  ->  out[(col*RC + rc)*T + tap]
//   conv3x3: RC = Ci, col = co, T = 9;   convT2x2: RC = Co, col = ci, T = 4
// One block = 32 columns x 8 `rc` values x all T taps; blockDim = (32, 8, WGR_SLICES). Thread
// (tx, ty, tz) sums, over every WGR_SLICES-th split, the T values of (col0+tx, rc0+ty) — reads are
// coalesced along the columns, the T chains are independent and the split loop (a chain of dependent
// load -> add steps) is cut WGR_SLICES ways — then the block combines the slices and transposes
// through shared memory so that the torch-layout rows out[col][rc0..rc0+8)[0..T) (8*T contiguous
// floats per column) are written with full sectors. Fixed summation order: deterministic.
constexpr int WGR_SLICES = 4;
template <int T>
static __global__ void __launch_bounds__(256 * WGR_SLICES)
wgrad_reduce_kernel(const float* __restrict__ ws, int splits, long long split_stride, int cols,
                    int RC, float* __restrict__ out, float* zero0, int nzero0, float* zero1,
                    int nzero1, int tn) {
    pdl_entry();
    __shared__ float sm[WGR_SLICES][32][8 * T + 1];
    const int tx = threadIdx.x, ty = threadIdx.y, tz = threadIdx.z;
    const int tid = (tz * 8 + ty) * 32 + tx;
    if (blockIdx.x == 0 && blockIdx.y == 0) {
        // bias gradients that are analytically zero in training mode (a bias ahead of a BatchNorm is
        // removed by the mean subtraction) are cleared here instead of by extra launches
        for (int i = tid; i < nzero0; i += 256 * WGR_SLICES) zero0[i] = 0.f;
        for (int i = tid; i < nzero1; i += 256 * WGR_SLICES) zero1[i] = 0.f;
    }
    const int col0 = blockIdx.x * 32, rc0 = blockIdx.y * 8;
    float acc[T];
#pragma unroll
    for (int t = 0; t < T; ++t) acc[t] = 0.f;
    // workspace rows are (t / tn, rc), its columns (t % tn, col): tn = 1 is the plain tap-major layout,
    // tn = 3 the shifted form (rows = filter row dy, column blocks = filter column dx)
    const long long ldw = (long long)cols * tn;
    const float* base = ws + (long long)(rc0 + ty) * ldw + col0 + tx;
    for (int k = tz; k < splits; k += WGR_SLICES) {
        const float* p = base + (long long)k * split_stride;
#pragma unroll
        for (int t = 0; t < T; ++t)
            acc[t] += __ldg(p + (long long)(t / tn) * RC * ldw + (long long)(t % tn) * cols);
    }
#pragma unroll
    for (int t = 0; t < T; ++t) sm[tz][tx][ty * T + t] = acc[t];
    __syncthreads();
    for (int i = tid; i < 32 * 8 * T; i += 256 * WGR_SLICES) {
        const int c = i / (8 * T), r = i % (8 * T);
        float v = sm[0][c][r];
#pragma unroll
        for (int z = 1; z < WGR_SLICES; ++z) v += sm[z][c][r];
        out[((long long)(col0 + c) * RC + rc0) * T + r] = v;
    }
}

static __global__ void fill_zero_kernel(float* p, long long n) {
    pdl_entry();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        p[i] = 0.f;
}

}  // namespace ub
