// Issue-side cost of tcgen05.mma / commit / fence in a single issuing thread (no TMA, garbage smem).
#include <cstdio>
#include "../unet_segmentation_b200/csrc/common.cuh"
using namespace ub;

template <int BN, int MMAS, bool FENCE, bool COMMIT, bool WAITBAR, int SHIFT = 0>
__global__ void __launch_bounds__(128, 1) k(int iters, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    __shared__ uint64_t bars[4];
    __shared__ uint32_t tslot;
    const uint32_t bar0 = smem_u32(&bars[0]), bar1 = smem_u32(&bars[1]), bar2 = smem_u32(&bars[2]);
    if (threadIdx.x == 0) { mbar_init(bar0, 1); mbar_init(bar1, 1); mbar_init(bar2, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<256>(smem_u32(&tslot));
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tslot;
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        uint32_t ph = 0;
        for (int it = 0; it < iters; ++it) {
            if (WAITBAR) { mbar_arrive(bar2); mbar_wait(bar2, ph); ph ^= 1u; }   // already-complete wait
            if (FENCE) tc_fence_after();
            const uint64_t da = make_smem_desc(base + (it & 1) * 16384 + SHIFT, 0, 1024);
            const uint64_t db = make_smem_desc(base + 65536 + (it & 1) * 32768, 0, 1024);
#pragma unroll
            for (int j = 0; j < MMAS; ++j)
                umma_bf16(tmem, da + (uint64_t)(2 * (j & 3)), db + (uint64_t)(2 * (j & 3)), idesc, 1u);
            if (COMMIT) umma_commit(bar0);
        }
        const long long t1 = clock64();
        umma_commit(bar1);
        mbar_wait(bar1, 0);
        const long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<256>(tmem); }
}

template <int BN, int MMAS, bool FENCE, bool COMMIT, bool WAITBAR, int SHIFT = 0>
void run(const char* name, long long* d) {
    const int smem = 65536 + 65536 + 2048, iters = 2000;
    cudaFuncSetAttribute(k<BN, MMAS, FENCE, COMMIT, WAITBAR, SHIFT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int r = 0; r < 2; ++r) k<BN, MMAS, FENCE, COMMIT, WAITBAR, SHIFT><<<1, 128, smem>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-44s BN=%3d: issue %6.1f cyc/iter, total %6.1f cyc/iter (exec floor %4d) %s\n", name, BN, h[0] / (double)iters,
           h[1] / (double)iters, MMAS * BN / 2, cudaGetErrorString(e));
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    run<64, 4, false, false, false, 128>("4 MMA, A start +128 B", d);
    run<64, 4, false, false, false, 256>("4 MMA, A start +256 B", d);
    run<64, 4, false, false, false, 512>("4 MMA, A start +512 B", d);
    run<128, 4, false, false, false, 0>("4 MMA aligned", d);
    run<128, 4, false, false, false, 128>("4 MMA, A start +128 B", d);
    run<256, 4, false, false, false, 128>("4 MMA, A start +128 B", d);
    run<64, 4, false, false, false>("4 MMA", d);
    run<64, 4, false, true, false>("4 MMA + commit", d);
    run<64, 4, true, true, false>("fence + 4 MMA + commit", d);
    run<64, 4, true, true, true>("wait + fence + 4 MMA + commit", d);
    run<64, 1, false, false, false>("1 MMA", d);
    run<64, 8, true, true, true>("wait + fence + 8 MMA + commit", d);
    run<128, 4, true, true, true>("wait + fence + 4 MMA + commit", d);
    run<256, 4, false, false, false>("4 MMA", d);
    run<256, 4, true, true, true>("wait + fence + 4 MMA + commit", d);
    run<256, 8, true, true, true>("wait + fence + 8 MMA + commit", d);
    return 0;
}
