// nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True): the `bilinear=True` branch of the
// reference's Up block (models/unet_model.py:40-43), forward and backward, NHWC bf16 with fp32
// arithmetic. HBM-bound: the output (forward) / the upstream gradient (backward) is 4x the input.
//
// Source position of output index o along an axis of length n -> 2n (align_corners):
//   s = o * (n - 1) / (2n - 1) (float, as ATen's area_pixel_compute_source_index), i0 = (int)s,
//   i1 = i0 + (i0 < n - 1), weights (1 - (s - i0), s - i0).
// Backward is the exact adjoint in gather form: an input index i collects every output index whose
// i0 or i1 equals i (<= 3 per tap at this scale), so no atomics and a deterministic sum.
#pragma once
#include "common.cuh"

namespace ub {

struct UpAxis {
    int i0, i1;
    float w0, w1;
};
__device__ __forceinline__ UpAxis up_axis(int o, int n, float ratio) {
    UpAxis a;
    const float s = ratio * (float)o;
    a.i0 = (int)s;
    a.i1 = a.i0 + (a.i0 < n - 1 ? 1 : 0);
    a.w1 = s - (float)a.i0;
    a.w0 = 1.f - a.w1;
    return a;
}
__host__ __device__ inline float up_ratio(int n) { return n > 0 ? (float)(n - 1) / (float)(2 * n - 1) : 0.f; }

__device__ __forceinline__ void up_fma8(float (&acc)[8], const uint4 v, float w) {
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(p[k]);
        acc[2 * k] = fmaf(w, f.x, acc[2 * k]);
        acc[2 * k + 1] = fmaf(w, f.y, acc[2 * k + 1]);
    }
}
__device__ __forceinline__ uint4 up_pack8(const float (&acc)[8]) {
    uint4 o;
    __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) p[k] = __floats2bfloat162_rn(acc[2 * k], acc[2 * k + 1]);
    return o;
}

// x: view [N, H, W, C] (C % 8 == 0, 16-byte aligned rows); out: contiguous [N, 2H, 2W, C].
static __global__ void __launch_bounds__(256)
upsample2x_fwd_kernel(const View x, __nv_bfloat16* __restrict__ out) {
    pdl_entry();
    const int C8 = x.C / 8, Ho = 2 * x.H, Wo = 2 * x.W;
    const float rh = up_ratio(x.H), rw = up_ratio(x.W);
    const long long total = (long long)x.N * Ho * Wo * C8;
    const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(x.ptr);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8);
        long long t = i / C8;
        const int ox = (int)(t % Wo); t /= Wo;
        const int oy = (int)(t % Ho);
        const int n = (int)(t / Ho);
        const UpAxis ay = up_axis(oy, x.H, rh), ax = up_axis(ox, x.W, rw);
        const __nv_bfloat16* p = base + n * x.sN + c8 * 8;
        float top[8] = {0, 0, 0, 0, 0, 0, 0, 0}, bot[8] = {0, 0, 0, 0, 0, 0, 0, 0}, o[8];
        up_fma8(top, *reinterpret_cast<const uint4*>(p + ay.i0 * x.sH + ax.i0 * x.sW), ax.w0);
        up_fma8(top, *reinterpret_cast<const uint4*>(p + ay.i0 * x.sH + ax.i1 * x.sW), ax.w1);
        up_fma8(bot, *reinterpret_cast<const uint4*>(p + ay.i1 * x.sH + ax.i0 * x.sW), ax.w0);
        up_fma8(bot, *reinterpret_cast<const uint4*>(p + ay.i1 * x.sH + ax.i1 * x.sW), ax.w1);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = ay.w0 * top[k] + ay.w1 * bot[k];
        *reinterpret_cast<uint4*>(out + i * 8) = up_pack8(o);
    }
}

// Output indices o in [lo, hi] can touch input index i (i0(o) in {i - 1, i}); exact membership is
// re-checked with up_axis, the bounds only have to be conservative.
__device__ __forceinline__ void up_candidates(int i, int n, float ratio, int& lo, int& hi) {
    const float inv = ratio > 0.f ? 1.f / ratio : 0.f;
    lo = (int)floorf((float)(i - 1) * inv) - 1;
    hi = (int)ceilf((float)(i + 1) * inv) + 1;
    if (lo < 0) lo = 0;
    if (hi > 2 * n - 1) hi = 2 * n - 1;
    if (ratio <= 0.f) { lo = 0; hi = 2 * n - 1; }   // n == 1: both outputs read input 0
}
__device__ __forceinline__ float up_weight_for(const UpAxis& a, int i) {
    return (a.i0 == i ? a.w0 : 0.f) + (a.i1 == i ? a.w1 : 0.f);
}

// g: view [N, 2H, 2W, C] (e.g. the up-sampled channel range of d(concat)); dx: contiguous [N, H, W, C].
static __global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(const View g, __nv_bfloat16* __restrict__ dx, int H, int W) {
    pdl_entry();
    const int C8 = g.C / 8;
    const float rh = up_ratio(H), rw = up_ratio(W);
    const long long total = (long long)g.N * H * W * C8;
    const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(g.ptr);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8);
        long long t = i / C8;
        const int ix = (int)(t % W); t /= W;
        const int iy = (int)(t % H);
        const int n = (int)(t / H);
        int ylo, yhi, xlo, xhi;
        up_candidates(iy, H, rh, ylo, yhi);
        up_candidates(ix, W, rw, xlo, xhi);
        const __nv_bfloat16* p = base + n * g.sN + c8 * 8;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int oy = ylo; oy <= yhi; ++oy) {
            const float wy = up_weight_for(up_axis(oy, H, rh), iy);
            if (wy == 0.f) continue;
            for (int ox = xlo; ox <= xhi; ++ox) {
                const float wx = up_weight_for(up_axis(ox, W, rw), ix);
                if (wx == 0.f) continue;
                up_fma8(acc, *reinterpret_cast<const uint4*>(p + oy * g.sH + ox * g.sW), wy * wx);
            }
        }
        *reinterpret_cast<uint4*>(dx + i * 8) = up_pack8(acc);
    }
}

}  // namespace ub
