// TMA cost model: cycles per operation vs box size (L2-resident source), one SM and all SMs.
#include <cstdio>
#include <vector>
#include "../unet_segmentation_b200/csrc/tmaps.cuh"
using namespace ub;
constexpr int RING = 200 * 1024;

__global__ void __launch_bounds__(64, 1)
k(const __grid_constant__ CUtensorMap map, int iters, int bytes, int stages, int box_rows, int nd, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t bar = base + RING;
    const uint32_t stage_bytes = (bytes + 1023) & ~1023;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(bar + 8 * s, 1); mbar_init(bar + 8 * (32 + s), 1); }
        fence_mbar_init();
    }
    __syncthreads();
    long long t0 = 0;
    if (threadIdx.x == 0) {
        t0 = clock64();
        int stage = 0; uint32_t phase = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(bar + 8 * (32 + stage), phase ^ 1u);
            mbar_expect_tx(bar + 8 * stage, bytes);
            const uint32_t dst = base + stage * stage_bytes;
            const int c = (it & 7) * 64, r = ((it * 5) & 15) * 4;
            if (nd == 2) tma_load_2d(dst, &map, bar + 8 * stage, c, r);
            else tma_load_3d(dst, &map, bar + 8 * stage, c, r, it & 1);
            if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
    } else if (threadIdx.x == 32) {
        int stage = 0; uint32_t phase = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(bar + 8 * stage, phase);
            mbar_arrive(bar + 8 * (32 + stage));
            if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

int main() {
    long long* d; cudaMalloc(&d, 148 * 8);
    const int smem = RING + 1024 + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    // source: [rows=2048][cols=1024] bf16 = 4 MB (L2 resident)
    const int rows = 2048, cols = 1024; void* buf; cudaMalloc(&buf, (size_t)rows * cols * 2); cudaMemset(buf, 0, (size_t)rows * cols * 2);
    TmapApi& api = tmap_api();
    struct C { int nd, r, z; };
    for (C c : std::vector<C>{{2, 16, 1}, {2, 32, 1}, {2, 64, 1}, {2, 128, 1}, {2, 256, 1}, {3, 130, 3}, {3, 64, 3}, {3, 128, 3}, {3, 64, 9}}) {
        CUtensorMap map;
        cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows / 16, 16};
        cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)cols * 2 * (rows / 16)};
        cuuint32_t box[3] = {64, (cuuint32_t)c.r, (cuuint32_t)c.z}; cuuint32_t es[3] = {1, 1, 1};
        if (c.nd == 2) { dims[1] = rows; }
        CUresult rc = api.tiled(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, c.nd, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) { printf("encode failed nd=%d r=%d z=%d rc=%d\n", c.nd, c.r, c.z, (int)rc); continue; }
        const int bytes = c.r * c.z * 128;
        int stages = RING / ((bytes + 1023) & ~1023); if (stages > 8) stages = 8;
        for (int grid : {1, 148}) {
            for (int rep = 0; rep < 2; ++rep) k<<<grid, 64, smem>>>(map, 1000, bytes, stages, c.r, c.nd, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
            printf("nd=%d box=(64,%3d,%d) %6d B stages %d grid %3d: %7.0f cyc/op %6.1f B/cyc/SM %s\n", c.nd, c.r, c.z, bytes, stages, grid, avg / 1000, bytes / (avg / 1000), cudaGetErrorString(e));
        }
    }
    return 0;
}
