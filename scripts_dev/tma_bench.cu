// Microbenchmark: per-SM throughput of TMA load modes (im2col vs tiled), to pick the A-operand
// feeding strategy of the implicit-GEMM kernels. Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../unet_segmentation_b200/csrc/tmaps.cuh"
using namespace ub;

constexpr int STAGE_BYTES = 17 * 1024;

// mode 0: im2col (128 pix x 64 ch), 1: tiled 4D box, 2: tiled 2D box
template <int STAGES>
__global__ void __launch_bounds__(64, 1)
tma_kernel(const __grid_constant__ CUtensorMap map, int mode, int iters, int bytes, int c_chunks,
           int W, int H, int N, int bw, int bh, long long* cycles_out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t bar = base + STAGES * STAGE_BYTES;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar + 8 * s, 1); mbar_init(bar + 8 * (STAGES + s), 1); }
        fence_mbar_init();
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    long long t0 = 0;
    if (threadIdx.x == 0) {
        t0 = clock64();
        int stage = 0; uint32_t phase = 0;
        unsigned rng = blockIdx.x * 7919u + 13u;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(bar + 8 * (STAGES + stage), phase ^ 1u);
            mbar_expect_tx(bar + 8 * stage, bytes);
            const int n = 0; (void)N; (void)rng; (void)H; (void)W;
            const int h = (it * 3) & 31;
            const int w = (it * 7) & 63;
            const int c = (it & (c_chunks - 1)) * 64;
            const uint32_t dst = base + stage * STAGE_BYTES;
            if (mode == 0) tma_load_im2col(dst, &map, bar + 8 * stage, c, w, h, n, (uint16_t)(it & 1), (uint16_t)((it >> 1) & 1));
            else if (mode == 1) {
                asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                             :: "r"(dst), "l"(reinterpret_cast<uint64_t>(&map)), "r"(bar + 8 * stage), "r"(c), "r"(w), "r"(h), "r"(n) : "memory");
            } else tma_load_2d(dst, &map, bar + 8 * stage, c, (it * 5) & 511);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
    } else if (warp == 1 && (threadIdx.x & 31) == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(bar + 8 * stage, phase);
            mbar_arrive(bar + 8 * (STAGES + stage));   // consume immediately
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles_out[blockIdx.x] = clock64() - t0;
}

template <int STAGES>
void run(const char* name, const CUtensorMap& map, int mode, int bytes, int cch, int W, int H, int N, int bw, int bh, int grid, long long* d_cycles) {
    const int iters = 2000;
    const int smem = STAGES * STAGE_BYTES + 256 + 1024;
    cudaFuncSetAttribute(tma_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 2; ++rep)
        tma_kernel<STAGES><<<grid, 64, smem>>>(map, mode, iters, bytes, cch, W, H, N, bw, bh, d_cycles);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d_cycles, grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
    printf("%-36s grid %3d stages %2d %6d B/op %7.0f cyc/op %6.1f B/cyc/SM (%s)\n", name, grid, STAGES, bytes, avg / iters, bytes / (avg / iters), cudaGetErrorString(e));
}

int main() {
    const int N = 16;
    long long* d_cycles; cudaMalloc(&d_cycles, 148 * 8);
    struct Cfg { const char* name; int mode, H, W, C, bw, bh; };
    std::vector<Cfg> cfgs = {
        {"im2col 128pix x64ch, C=64", 0, 510, 510, 64, 128, 1},
        {"tiled  (64,128,1,1), C=64", 1, 510, 510, 64, 128, 1},
        {"tiled  (64,16,8,1),  C=64", 1, 510, 510, 64, 16, 8},
        {"im2col 128pix x64ch, C=1024", 0, 26, 26, 1024, 8, 8},
        {"tiled  (64,8,8,2),   C=1024", 1, 26, 26, 1024, 8, 8},
        {"tiled2D 64 x 128 rows", 2, 0, 0, 0, 0, 0},
    };
    for (auto& c : cfgs) {
        CUtensorMap map; int bytes = 0, cch = 1; void* buf = nullptr;
        if (c.mode == 2) {
            const int rows = 1024, K = 9216; cudaMalloc(&buf, (size_t)rows * K * 2); cudaMemset(buf, 0, (size_t)rows * K * 2);
            make_tmap_2d(&map, buf, K, rows, (unsigned long long)K * 2, 128); bytes = 128 * 128; cch = K / 64;
        } else {
            size_t elems = (size_t)N * c.H * c.W * c.C; cudaMalloc(&buf, elems * 2); cudaMemset(buf, 0, elems * 2);
            View v = make_view(buf, N, c.H, c.W, c.C); cch = c.C / 64;
            if (c.mode == 0) { make_tmap_im2col(&map, v, 0, -2, 1, 128); bytes = 128 * 128; }
            else {
                TmapApi& api = tmap_api();
                const int bn = (c.bw * c.bh == 64) ? 2 : 1;
                cuuint64_t dims[4] = {(cuuint64_t)c.C, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)N};
                cuuint64_t strides[3] = {(cuuint64_t)c.C * 2, (cuuint64_t)c.W * c.C * 2, (cuuint64_t)c.H * c.W * c.C * 2};
                cuuint32_t box[4] = {64, (cuuint32_t)c.bw, (cuuint32_t)c.bh, (cuuint32_t)bn};
                cuuint32_t es[4] = {1, 1, 1, 1};
                api.tiled(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                bytes = c.bw * c.bh * bn * 128;
            }
        }
        const int W = c.W ? c.W : 600, H = c.H ? c.H : 600;
        for (int grid : {1, 148}) {
            run<2>(c.name, map, c.mode, bytes, cch, W, H, N, c.bw, c.bh, grid, d_cycles);
            run<6>(c.name, map, c.mode, bytes, cch, W, H, N, c.bw, c.bh, grid, d_cycles);
            run<12>(c.name, map, c.mode, bytes, cch, W, H, N, c.bw, c.bh, grid, d_cycles);
        }
        cudaFree(buf);
    }
    return 0;
}
