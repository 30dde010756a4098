// Device-side weight-map generation: calculate_weight_map of the reference
// (scripts/preprocess_data.py:17-77) for a batch of instance-label masks.
//
// The reference's border term is degenerate: per instance it takes
// minimum(edt(obj), edt(obj == 0)) (:47), which is zero at every pixel (the first transform vanishes
// outside the object, the second inside), so d1 = d2 = 0 (:52-64) and the term is w0 * exp(-0.0)
// = w0 (:72) for every possible mask. What remains is
//     weight = (double)(float)(1 / (class pixels / all pixels)) + w0           (:26-36, :75)
// per pixel, with weight 0 for an absent class — two values per image. Parity target: the stored
// float64 .npy maps, bit-exact. Two HBM-bound passes: count foreground pixels per image (warp
// reductions + one atomic per warp), then emit 4 pixels per thread (1 if H * W % 4 != 0): float64
// as stored by the reference, or float32 = torch.from_numpy(map).float() of utils/dataset.py:110.
#pragma once
#include "common.cuh"

namespace ub {

// VEC = 4: one 4- / 8-byte load per thread (needs H * W % 4 == 0 so that every image of the batch
// stays aligned); VEC = 1: scalar path for any other size.
template <int VEC>
__device__ __forceinline__ void load_fg(const unsigned char* p, bool fg[VEC]) {
    if constexpr (VEC == 4) {
        const uchar4 v = *reinterpret_cast<const uchar4*>(p);
        fg[0] = v.x > 0; fg[1] = v.y > 0; fg[2] = v.z > 0; fg[3] = v.w > 0;
    } else {
        fg[0] = *p > 0;
    }
}
template <int VEC>
__device__ __forceinline__ void load_fg(const unsigned short* p, bool fg[VEC]) {
    if constexpr (VEC == 4) {
        const ushort4 v = *reinterpret_cast<const ushort4*>(p);
        fg[0] = v.x > 0; fg[1] = v.y > 0; fg[2] = v.z > 0; fg[3] = v.w > 0;
    } else {
        fg[0] = *p > 0;
    }
}

// counts[n] += number of labels > 0 in image n. grid = (blocks per image, N); P = H * W.
template <typename LabelT, int VEC>
static __global__ void __launch_bounds__(256)
wmap_count_kernel(const LabelT* __restrict__ labels, unsigned P, unsigned* __restrict__ counts) {
    pdl_entry();
    const LabelT* img = labels + (size_t)blockIdx.y * P;
    const unsigned PV = P / VEC;
    unsigned c = 0;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < PV; i += gridDim.x * blockDim.x) {
        bool fg[VEC];
        load_fg<VEC>(img + VEC * (size_t)i, fg);
#pragma unroll
        for (int e = 0; e < VEC; ++e) c += fg[e] ? 1u : 0u;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(counts + blockIdx.y, c);
}

// out[n][p] = class weight of pixel p (see the file header). OutT = double or float.
template <typename LabelT, typename OutT, int VEC>
static __global__ void __launch_bounds__(256)
wmap_emit_kernel(const LabelT* __restrict__ labels, unsigned P, const unsigned* __restrict__ counts,
                 double border, OutT* __restrict__ out) {
    pdl_entry();
    __shared__ OutT s_w[2];   // [0] background, [1] foreground
    if (threadIdx.x == 0) {
        const unsigned n_fg = counts[blockIdx.y], n_bg = P - n_fg;
        // IEEE double divisions exactly as numpy / Python evaluate 1.0 / (n / total)
        const double wc_bg = n_bg ? __ddiv_rn(1.0, __ddiv_rn((double)n_bg, (double)P)) : 0.0;
        const double wc_fg = n_fg ? __ddiv_rn(1.0, __ddiv_rn((double)n_fg, (double)P)) : 0.0;
        // assignment into the float32 wc_map rounds; the sum with the float64 border term is float64
        s_w[0] = (OutT)__dadd_rn((double)__double2float_rn(wc_bg), border);
        s_w[1] = (OutT)__dadd_rn((double)__double2float_rn(wc_fg), border);
    }
    __syncthreads();
    const OutT w_bg = s_w[0], w_fg = s_w[1];
    const LabelT* img = labels + (size_t)blockIdx.y * P;
    OutT* o = out + (size_t)blockIdx.y * P;
    const unsigned PV = P / VEC;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < PV; i += gridDim.x * blockDim.x) {
        bool fg[VEC];
        load_fg<VEC>(img + VEC * (size_t)i, fg);
        OutT v[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) v[e] = fg[e] ? w_fg : w_bg;
        if constexpr (VEC == 1) {
            o[i] = v[0];
        } else if constexpr (sizeof(OutT) == 8) {
            double2* d = reinterpret_cast<double2*>(o + 4 * (size_t)i);
            d[0] = make_double2((double)v[0], (double)v[1]);
            d[1] = make_double2((double)v[2], (double)v[3]);
        } else {
            *reinterpret_cast<float4*>(o + 4 * (size_t)i) =
                make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
        }
    }
}

}  // namespace ub
