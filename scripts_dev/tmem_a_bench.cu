// A operand of tcgen05.mma from TENSOR MEMORY: correctness of
// tcgen05.cp (shared -> tensor memory, 128x256b) + TS-form MMA against the SS form, and what the pair
// costs — does the copy overlap the MMAs or serialise with them in the tensor pipe?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tmem_a_bench tmem_a_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../unet_segmentation_b200/csrc/common.cuh"
using namespace ub;

__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// physical element offset of (row, k) in a K-major [rows][64 bf16] tile with the 128-byte swizzle
__host__ __device__ inline int swz(int row, int k) {
    const int chunk = (k >> 3) ^ (row & 7);
    return row * 64 + chunk * 8 + (k & 7);
}

constexpr int AROWS = 136;   // 128 + room for a shifted window
// mode 0: SS MMA; mode 1: tcgen05.cp + TS MMA; `shift` = window start in rows (pixels)
__global__ void __launch_bounds__(128, 1) check_kernel(const __nv_bfloat16* A, const __nv_bfloat16* B,
                                                       float* D, int mode, int shift) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* g = smem_raw + (base - raw);
    __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(g);
    __nv_bfloat16* sB = reinterpret_cast<__nv_bfloat16*>(g + 32768);
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < AROWS * 64; i += 128) sA[swz(i / 64, i % 64)] = A[i];
    for (int i = threadIdx.x; i < 64 * 64; i += 128) sB[swz(i / 64, i % 64)] = B[i];
    fence_proxy_async();
    const uint32_t bar0 = smem_u32(&bars[0]);
    if (threadIdx.x == 0) { mbar_init(bar0, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<256>(smem_u32(&tslot));
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tslot;
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    if (threadIdx.x == 0) {
        const uint64_t da = make_smem_desc(base + shift * 128, 0, 1024);
        const uint64_t db = make_smem_desc(base + 32768, 0, 1024);
        const uint32_t ta = tmem + 128;        // A operand columns [128, 160)
        if (mode == 1)
            for (int k = 0; k < 4; ++k) tmem_cp_128x256b(ta + 8 * k, da + (uint64_t)(2 * k));
        for (int k = 0; k < 4; ++k) {
            if (mode == 0) umma_bf16(tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, k != 0);
            else umma_bf16_ts(tmem, ta + 8 * k, db + (uint64_t)(2 * k), idesc, k != 0);
        }
        umma_commit(bar0);
    }
    mbar_wait(bar0, 0);
    tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, r);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) D[(warp * 32 + lane) * 64 + c * 32 + i] = __uint_as_float(r[i]);
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<256>(tmem); }
}

// what: 0 = 36 SS MMAs per iteration, 1 = 36 TS MMAs, 2 = 12 copies only, 3 = 12 copies + 36 TS MMAs
// (one output-row tile of the 64 -> 64 conv: one new input row in three shifts, nine taps x 4 K steps)
template <int what>
__global__ void __launch_bounds__(128, 1) speed_kernel(int iters, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tslot;
    const uint32_t bar0 = smem_u32(&bars[0]);
    if (threadIdx.x == 0) { mbar_init(bar0, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(smem_u32(&tslot));
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tslot;
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    if (threadIdx.x == 0) {
        // descriptors are hoisted and every offset below is a compile-time constant (fully unrolled),
        // as in the production kernels: the loop must be bound by the tensor pipe, not by issue
        const uint64_t db0 = make_smem_desc(base + 110592, 0, 1024);
        const long long t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
            const uint32_t rowbuf = base + (it & 1) * 3 * 17408;      // 3 staged input rows
            const uint64_t da0 = make_smem_desc(rowbuf, 0, 1024);
            const uint32_t td = tmem + (it & 1) * 64;
            uint32_t trow[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) trow[r] = tmem + 128 + (((it + r) & 3) * 96);
            if (what >= 2) {
#pragma unroll
                for (int s = 0; s < 3; ++s)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tmem_cp_128x256b(trow[2] + s * 32 + 8 * k, da0 + (uint64_t)(s * 8 + 2 * k));
            }
            if (what != 2) {
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t db = db0 + (uint64_t)(tap * 512 + 2 * k);
                        if (what == 0)
                            umma_bf16(td, da0 + (uint64_t)((tap / 3) * 1088 + (tap % 3) * 8 + 2 * k), db, idesc,
                                      (uint32_t)((tap | k) != 0));
                        else
                            umma_bf16_ts(td, trow[tap / 3] + (tap % 3) * 32 + 8 * k, db, idesc,
                                         (uint32_t)((tap | k) != 0));
                    }
                }
            }
        }
        const long long t1 = clock64();
        umma_commit(bar0);
        mbar_wait(bar0, 0);
        const long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

int main() {
    const int smem = 110592 + 9 * 8192 + 2048;
    std::vector<__nv_bfloat16> hA(AROWS * 64), hB(64 * 64);
    std::vector<float> fA(AROWS * 64), fB(64 * 64);
    srand(1);
    for (size_t i = 0; i < hA.size(); ++i) { fA[i] = (float)(rand() % 7 - 3); hA[i] = __float2bfloat16(fA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { fB[i] = (float)(rand() % 5 - 2); hB[i] = __float2bfloat16(fB[i]); }
    __nv_bfloat16 *dA, *dB; float* dD; long long* dT;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 64 * 4); cudaMalloc(&dT, 64);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> hD(128 * 64);
    for (int mode = 0; mode < 2; ++mode)
        for (int shift = 0; shift < 3; ++shift) {
            cudaMemset(dD, 0, 128 * 64 * 4);
            check_kernel<<<1, 128, smem>>>(dA, dB, dD, mode, shift);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0; double maxerr = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 64; ++n) {
                    float ref = 0;
                    for (int k = 0; k < 64; ++k) ref += fA[(m + shift) * 64 + k] * fB[n * 64 + k];
                    const double err = fabs(ref - hD[m * 64 + n]);
                    if (err > 1e-3) ++bad;
                    if (err > maxerr) maxerr = err;
                }
            printf("check %s shift %d: %d / 8192 wrong, max |err| %.3g  (%s)\n",
                   mode ? "tcgen05.cp + TS MMA" : "SS MMA            ", shift, bad, maxerr, cudaGetErrorString(e));
            if (e != cudaSuccess) return 1;
        }
    const char* names[4] = {"36 SS MMA (N=64)", "36 TS MMA (A in TMEM)", "12 tcgen05.cp 128x256b (48 KB)",
                            "12 tcgen05.cp + 36 TS MMA"};
    const int iters = 2000;
    auto run = [&](auto kern, int what) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int r = 0; r < 2; ++r) kern<<<1, 128, smem>>>(iters, dT);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, dT, 16, cudaMemcpyDeviceToHost);
        printf("%-34s issue %7.1f cyc/tile, total %7.1f cyc/tile (MMA floor 1152)  %s\n", names[what],
               h[0] / (double)iters, h[1] / (double)iters, cudaGetErrorString(e));
        return e == cudaSuccess;
    };
    if (!run(speed_kernel<0>, 0) || !run(speed_kernel<1>, 1) || !run(speed_kernel<2>, 2) ||
        !run(speed_kernel<3>, 3)) return 1;
    return 0;
}
