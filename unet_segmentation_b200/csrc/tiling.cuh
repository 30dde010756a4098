// Overlap-tile inference helpers (semantics in unet_segmentation_b200/tiling.py): the
// mirror-extended tile gather in front of the network and the stitch of finished uint8 tiles into the
// whole-image mask behind it. Both are one launch for a whole batch of tiles, driven by a device
// table of output-space tile origins, so the predict path has no per-tile host work and can be
// replayed from a CUDA graph. HBM-bound byte movers (4 B/px in + 4 B/px out; 1 + 1 B/px).
#pragma once
#include "common.cuh"

namespace ub {

// numpy 'reflect' (no edge repeat) extension of arbitrary length: index into [0, n)
__device__ __forceinline__ int reflect_index(int i, int n) {
    if (n == 1) return 0;
    const int period = 2 * (n - 1);
    i %= period;
    if (i < 0) i += period;
    return i >= n ? period - i : i;
}

// tiles[t][y][x] = image[reflect(oy_t - margin + y)][reflect(ox_t - margin + x)], S % 4 == 0.
// origins: int2 (y, x) per tile in output coordinates; a negative y marks an unused slot (zeros).
static __global__ void __launch_bounds__(256)
extract_tiles_kernel(const float* __restrict__ image, int H, int W, const int2* __restrict__ origins,
                     int T, int S, int margin, float* __restrict__ tiles) {
    const int S4 = S >> 2;
    const long long total = (long long)T * S * S4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int x4 = (int)(i % S4);
        const long long r = i / S4;
        const int y = (int)(r % S), t = (int)(r / S);
        const int2 o = origins[t];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (o.x >= 0) {
            const int iy = reflect_index(o.x - margin + y, H);       // o.x = row origin (y), o.y = column
            const int x0 = o.y - margin + x4 * 4;
            const float* row = image + (long long)iy * W;
            if (x0 >= 0 && x0 + 3 < W && (((long long)iy * W + x0) & 3) == 0) {
                v = __ldg(reinterpret_cast<const float4*>(row + x0));
            } else {
                v.x = __ldg(row + reflect_index(x0, W));
                v.y = __ldg(row + reflect_index(x0 + 1, W));
                v.z = __ldg(row + reflect_index(x0 + 2, W));
                v.w = __ldg(row + reflect_index(x0 + 3, W));
            }
        }
        reinterpret_cast<float4*>(tiles)[i] = v;
    }
}

// full[oy + y][ox + x] = tiles[t][y][x] for the part of the tile inside the H x W image.
// Tiles overlap only where their values are equal (aligned-tile invariant), so write order is free.
// TO % 4 == 0 and origins / W multiples of 4: 4 bytes per thread.
static __global__ void __launch_bounds__(256)
stitch_tiles_kernel(const unsigned char* __restrict__ tiles, const int2* __restrict__ origins, int T,
                    int TO, unsigned char* __restrict__ full, int H, int W) {
    const int T4 = TO >> 2;
    const long long total = (long long)T * TO * T4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int x4 = (int)(i % T4);
        const long long r = i / T4;
        const int y = (int)(r % TO), t = (int)(r / TO);
        const int2 o = origins[t];
        if (o.x < 0) continue;
        const int gy = o.x + y, gx = o.y + x4 * 4;
        if (gy >= H || gx >= W) continue;
        const unsigned v = __ldg(reinterpret_cast<const unsigned*>(tiles) + i);
        unsigned char* dst = full + (long long)gy * W + gx;
        if (gx + 3 < W && (((long long)gy * W + gx) & 3) == 0) {
            *reinterpret_cast<unsigned*>(dst) = v;
        } else {
            for (int k = 0; k < 4 && gx + k < W; ++k) dst[k] = (unsigned char)(v >> (8 * k));
        }
    }
}

}  // namespace ub
