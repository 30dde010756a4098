// Fused multi-tensor SGD-momentum step + bf16 operand-cache refresh.
// Replaces torch.optim.SGD(lr, momentum).step() (reference scripts/train.py:97,131) followed by the
// separate weight re-packing pass: every parameter is read once (p, grad, momentum buffer), updated
// with torch's exact formula, written back, and — for convolution weights — emitted in the same pass
// as the tap-major bf16 GEMM operands the conv kernels consume.
//   g' = g + wd*p;  buf = g' (first step) | momentum*buf + (1-dampening)*g';
//   step = nesterov ? g' + momentum*buf : buf;   p -= lr*step
#pragma once
#include "common.cuh"

namespace ub {

enum : int { SGD_PLAIN = 0, SGD_CONV3 = 1, SGD_CONVT = 2, SGD_CONVT_BIAS = 3 };

struct SgdTensor {
    float* p;
    const float* g;
    float* buf;
    __nv_bfloat16* outA;   // [t][d0][d1]   (conv: fprop operand;  convT: data-gradient operand)
    __nv_bfloat16* outB;   // [tb][d1][d0]  (conv: data-gradient operand, tb = T-1-t;  convT: fprop, tb = t)
    float* bias4;          // SGD_CONVT_BIAS: bias replicated over the 4 sub-pixel positions
    int kind, d0, d1, T;   // parameter shape [d0][d1][T]
    int n;                 // elements
    int block_begin;       // first block of this tensor inside the launch
};
constexpr int SGD_MAX_TENSORS = 36;
struct SgdBatch {
    SgdTensor t[SGD_MAX_TENSORS];
    int count;
    float lr, momentum, dampening, weight_decay;
    int nesterov, first_step;
    int vec;   // 16-byte accesses where the pointers allow it (UB_SGD_VEC=0: scalar path, for A/B runs)
};

__device__ __forceinline__ float sgd_update(float p, float g, float* buf_io, const SgdBatch& B) {
    if (B.weight_decay != 0.f) g = fmaf(B.weight_decay, p, g);
    float step = g;
    if (B.momentum != 0.f) {
        float b = B.first_step ? g : fmaf(B.momentum, *buf_io, (1.f - B.dampening) * g);
        *buf_io = b;
        step = B.nesterov ? fmaf(B.momentum, b, g) : b;
    }
    return p - B.lr * step;
}

// One block = one 32 x 32 x T tile of a conv weight (staged through shared memory so that both
// packed layouts are written in 64-byte runs), or 2048 elements of a plain tensor.
static __global__ void __launch_bounds__(256)
sgd_fused_kernel(const __grid_constant__ SgdBatch B) {
    pdl_entry();
    __shared__ __nv_bfloat16 tile[9 * 32 * 33];
    int ti = 0;
    for (int i = 1; i < B.count; ++i)
        if ((int)blockIdx.x >= B.t[i].block_begin) ti = i;
    const SgdTensor& t = B.t[ti];
    const int blk = blockIdx.x - t.block_begin;
    if (t.kind == SGD_PLAIN || t.kind == SGD_CONVT_BIAS) {
        const int base = blk * 2048;
        for (int i = base + threadIdx.x; i < base + 2048 && i < t.n; i += 256) {
            float bv = t.buf ? t.buf[i] : 0.f;
            const float np = sgd_update(t.p[i], t.g[i], &bv, B);
            t.p[i] = np;
            if (t.buf) t.buf[i] = bv;
            if (t.kind == SGD_CONVT_BIAS && t.bias4) {
#pragma unroll
                for (int q = 0; q < 4; ++q) t.bias4[q * t.d0 + i] = np;
            }
        }
        return;
    }
    const int T = t.T, d0 = t.d0, d1 = t.d1;
    const int tiles1 = d1 / 32;
    const int a0 = (blk / tiles1) * 32, b0 = (blk % tiles1) * 32;   // tile origin in (d0, d1)
    const int row_len = 32 * T;                                        // contiguous floats per d0 row
    // 16-byte accesses: a tile row starts at ((a0+l0)*d1 + b0)*T floats, a multiple of 32 (d1, b0
    // are multiples of 32), so only the base pointers decide the alignment.
    const bool vec = B.vec && (((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) |
                        reinterpret_cast<uintptr_t>(t.buf)) & 15u) == 0u);
    if (vec) {
        const int row4 = row_len / 4;
        for (int idx = threadIdx.x; idx < 32 * row4; idx += 256) {
            const int l0 = idx / row4, r = (idx % row4) * 4;
            const long long gi = ((long long)(a0 + l0) * d1 + b0) * T + r;
            const float4 pv = *reinterpret_cast<const float4*>(t.p + gi);
            const float4 gv = __ldg(reinterpret_cast<const float4*>(t.g + gi));
            float4 bv = t.buf ? *reinterpret_cast<const float4*>(t.buf + gi) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 np;
            np.x = sgd_update(pv.x, gv.x, &bv.x, B);
            np.y = sgd_update(pv.y, gv.y, &bv.y, B);
            np.z = sgd_update(pv.z, gv.z, &bv.z, B);
            np.w = sgd_update(pv.w, gv.w, &bv.w, B);
            *reinterpret_cast<float4*>(t.p + gi) = np;
            if (t.buf) *reinterpret_cast<float4*>(t.buf + gi) = bv;
            const float nv[4] = {np.x, np.y, np.z, np.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int l1 = (r + e) / T, tap = (r + e) % T;
                tile[(tap * 32 + l0) * 33 + l1] = __float2bfloat16_rn(nv[e]);
            }
        }
    } else {
        for (int idx = threadIdx.x; idx < 32 * row_len; idx += 256) {
            const int l0 = idx / row_len, r = idx % row_len;
            const int l1 = r / T, tap = r % T;
            const long long gi = ((long long)(a0 + l0) * d1 + b0) * T + r;
            float bv = t.buf ? t.buf[gi] : 0.f;
            const float np = sgd_update(t.p[gi], t.g[gi], &bv, B);
            t.p[gi] = np;
            if (t.buf) t.buf[gi] = bv;
            tile[(tap * 32 + l0) * 33 + l1] = __float2bfloat16_rn(np);
        }
    }
    __syncthreads();
    // Both packed layouts are written as 16-byte pieces (8 bf16) of their 64-byte runs; the packed
    // operand buffers are library allocations (256-byte aligned), rows are multiples of 32 elements.
    const unsigned short* tl = reinterpret_cast<const unsigned short*>(tile);
    // outA[tap][d0][d1]: rows (tap, l0), 32 consecutive d1
    if (t.outA) {
        for (int idx = threadIdx.x; idx < T * 32 * 4; idx += 256) {
            const int q = idx % 4, l0 = (idx / 4) % 32, tap = idx / 128;
            const unsigned short* src = tl + (tap * 32 + l0) * 33 + q * 8;
            uint4 o;
            o.x = src[0] | ((uint32_t)src[1] << 16);
            o.y = src[2] | ((uint32_t)src[3] << 16);
            o.z = src[4] | ((uint32_t)src[5] << 16);
            o.w = src[6] | ((uint32_t)src[7] << 16);
            *reinterpret_cast<uint4*>(t.outA + ((long long)tap * d0 + a0 + l0) * d1 + b0 + q * 8) = o;
        }
    }
    // outB[tb][d1][d0]: rows (tb, l1), 32 consecutive d0
    if (t.outB) {
        for (int idx = threadIdx.x; idx < T * 32 * 4; idx += 256) {
            const int q = idx % 4, l1 = (idx / 4) % 32, tap = idx / 128;
            const int tb = t.kind == SGD_CONV3 ? T - 1 - tap : tap;
            const unsigned short* src = tl + (tap * 32 + q * 8) * 33 + l1;
            uint4 o;
            o.x = src[0 * 33] | ((uint32_t)src[1 * 33] << 16);
            o.y = src[2 * 33] | ((uint32_t)src[3 * 33] << 16);
            o.z = src[4 * 33] | ((uint32_t)src[5 * 33] << 16);
            o.w = src[6 * 33] | ((uint32_t)src[7 * 33] << 16);
            *reinterpret_cast<uint4*>(t.outB + ((long long)tb * d1 + b0 + l1) * d0 + a0 + q * 8) = o;
        }
    }
}

}  // namespace ub
