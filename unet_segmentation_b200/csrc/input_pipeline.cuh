// Device-side input pipeline: what the reference does on the host for every
// sample — ToTensor (uint8 -> float32 / 255, utils/dataset.py:92-94), binarisation of the instance
// mask to an int64 target ((mask > 0).long(), utils/dataset.py:96-105), the float cast of the float64
// weight map (:110) and the centre crop of target / weight map to the logits size
// (scripts/train.py:39-51,118-126) — as ONE pass over a compactly transferred batch: the host ships
// 1 + 1|2 + 4|8 bytes per pixel instead of 4 + 8 + 4.
#pragma once
#include "common.cuh"

namespace ub {

struct PrepArgs {
    const unsigned char* img;   // [N][H][W] uint8
    const void* labels;         // [N][H][W] uint8 / uint16 instance labels (may be null)
    const void* wmap;           // [N][H][W] float32 / float64 (may be null)
    int label_bytes, wmap_bytes;
    int N, H, W, oh, ow, h0, w0;   // crop window [h0, h0+oh) x [w0, w0+ow)
    float* image;               // [N][1][H][W] float32 in [0, 1]
    long long* target;          // [N][oh][ow] int64 in {0, 1}
    float* weight;              // [N][oh][ow] float32
};

// One thread = 4 horizontally adjacent pixels (W % 4 == 0 required by the host wrapper).
static __global__ void __launch_bounds__(256)
prepare_batch_kernel(const PrepArgs A) {
    pdl_entry();
    const unsigned W4 = (unsigned)A.W >> 2;
    const unsigned total = (unsigned)A.N * A.H * W4;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned w4 = i % W4, t = i / W4, h = t % (unsigned)A.H, n = t / (unsigned)A.H;
        const size_t p0 = ((size_t)n * A.H + h) * A.W + w4 * 4;
        const uchar4 u = *reinterpret_cast<const uchar4*>(A.img + p0);
        float4 f;   // the division ToTensor performs (IEEE, round to nearest)
        f.x = __fdiv_rn((float)u.x, 255.f); f.y = __fdiv_rn((float)u.y, 255.f);
        f.z = __fdiv_rn((float)u.z, 255.f); f.w = __fdiv_rn((float)u.w, 255.f);
        *reinterpret_cast<float4*>(A.image + p0) = f;
        const int hc = (int)h - A.h0;
        if (hc < 0 || hc >= A.oh) continue;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int wc = (int)(w4 * 4 + e) - A.w0;
            if (wc < 0 || wc >= A.ow) continue;
            const size_t src = p0 + e, dst = ((size_t)n * A.oh + hc) * A.ow + wc;
            if (A.labels) {
                const unsigned lbl = A.label_bytes == 2
                                         ? reinterpret_cast<const unsigned short*>(A.labels)[src]
                                         : reinterpret_cast<const unsigned char*>(A.labels)[src];
                A.target[dst] = lbl > 0u ? 1 : 0;
            }
            if (A.wmap)
                A.weight[dst] = A.wmap_bytes == 8 ? (float)reinterpret_cast<const double*>(A.wmap)[src]
                                                  : reinterpret_cast<const float*>(A.wmap)[src];
        }
    }
}

}  // namespace ub
