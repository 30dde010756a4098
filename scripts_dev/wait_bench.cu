// Cost of an already-complete mbarrier wait, expect_tx and TMA issue in one thread.
#include <cstdio>
#include "../unet_segmentation_b200/csrc/common.cuh"
using namespace ub;
__device__ __forceinline__ bool test_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
__global__ void k(long long* out) {
    __shared__ uint64_t bars[2];
    const uint32_t b0 = smem_u32(&bars[0]);
    if (threadIdx.x == 0) { mbar_init(b0, 1); fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int iters = 2000;
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) mbar_wait(b0, 1);            // already complete (fresh barrier)
        long long t1 = clock64();
        for (int i = 0; i < iters; ++i) { while (!mbar_try_wait(b0, 1)) {} }
        long long t2 = clock64();
        for (int i = 0; i < iters; ++i) { while (!test_wait(b0, 1)) {} }
        long long t3 = clock64();
        uint32_t ph = 0;
        for (int i = 0; i < iters; ++i) { mbar_arrive(b0); while (!mbar_try_wait(b0, ph)) {} ph ^= 1u; }
        long long t4 = clock64();
        for (int i = 0; i < iters; ++i) { mbar_arrive(b0); while (!test_wait(b0, ph)) {} ph ^= 1u; }
        long long t5 = clock64();
        long long c = 0;
        for (int i = 0; i < iters; ++i) c += clock64();
        long long t6 = clock64();
        out[0] = (t1 - t0); out[1] = (t2 - t1); out[2] = (t3 - t2); out[3] = (t4 - t3); out[4] = (t5 - t4); out[5] = (t6 - t5); out[6] = c;
    }
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    k<<<1, 32>>>(d); k<<<1, 32>>>(d);
    cudaDeviceSynchronize();
    long long h[7]; cudaMemcpy(h, d, 56, cudaMemcpyDeviceToHost);
    const char* n[6] = {"mbar_wait (lib) complete", "try_wait loop complete", "test_wait loop complete", "arrive + try_wait loop", "arrive + test_wait loop", "clock64()"};
    for (int i = 0; i < 6; ++i) printf("%-28s %6.1f cycles\n", n[i], h[i] / 2000.0);
    return 0;
}
