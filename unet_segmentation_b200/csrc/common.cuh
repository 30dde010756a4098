// Shared device-side helpers for the sm_100a kernels of libunetb200:
// mbarrier, TMA (tiled + im2col), tcgen05 (alloc / mma / commit / ld), UMMA descriptors.
// Everything here is inline PTX for sm_100a; there is no fallback path.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ub {

// ---------------------------------------------------------------------------------------------
// NHWC view of an activation tensor (element strides). A centre crop or a channel slice of a
// bigger buffer is just another view: the skip-connection crop (reference
// models/unet_model.py:88-102) and the channel concat (:131,135,139,143) never copy.
// ---------------------------------------------------------------------------------------------
struct View {
    const void* ptr;
    int N, H, W, C;
    long long sN, sH, sW;  // element strides; channel stride is 1
};

__host__ __device__ inline View make_view(const void* p, int N, int H, int W, int C) {
    View v;
    v.ptr = p; v.N = N; v.H = H; v.W = W; v.C = C;
    v.sW = C; v.sH = (long long)W * C; v.sN = (long long)H * W * C;
    return v;
}

// ---------------------------------------------------------------------------------------------
// Generic small helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL), opt-in (UB_PDL=1; measured slower on the full step, see
// pdl_enabled()). With it every kernel of the library is launched with the
// programmatic-stream-serialization attribute (ub_launch, ub_internal.h) and
//   * signals `launch_dependents` first thing, so the NEXT kernel's CTAs are scheduled while this
//     grid drains (their launch latency and prologue overlap our tail), and
//   * executes `griddepcontrol.wait` BEFORE its first global-memory access (read or write): the
//     wait returns when the preceding grid has completed and its writes are visible. Completion
//     order is therefore still the stream order, transitively.
// Both instructions are no-ops when a kernel is launched without the attribute.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_trigger() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void pdl_entry() {
    pdl_trigger();
    pdl_wait();
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a pipeline bug must trap (error returned to the host), never hang the GPU.
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 6000000000LL) __trap();  // ~3 s at 2 GHz
    }
}

// ---------------------------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// im2col-mode load of (channels x pixels): `pixels` consecutive base pixels in (w,h,n) raster
// order inside the map's bounding box starting at (w,h,n); every base pixel is displaced by the
// filter-tap offset (offw, offh). Out-of-tensor pixels are zero filled.
__device__ __forceinline__ void tma_load_im2col(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                                int c, int w, int h, int n, uint16_t offw,
                                                uint16_t offh) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n),
        "h"(offw), "h"(offh)
        : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 operands, fp32 accumulate. One thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
// 32 lanes x 32 columns of fp32: thread t of the warp receives row (lane base + t), 32 columns.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): two CTAs of a 2-CTA cluster (one TPC) execute ONE 256-row MMA. Each CTA
// stages its own 128 rows of A and HALF of the B tile, so the B operand crosses L2->SM and the
// shared-memory read port once per pair instead of once per CTA. The MMA is issued by the leader
// (cluster rank 0); TMA loads of both CTAs complete on the leader's stage barrier; tcgen05.commit
// multicasts its arrival to the barrier at the same offset in both CTAs.
// All helpers are templated on CG (1 = single CTA, 2 = pair) so one kernel body serves both.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\t"
                 "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on an mbarrier given by a shared::cluster address (possibly in the peer CTA)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
                 : "memory");
}
template <int CG>
__device__ __forceinline__ void tma_load_3d_cg(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                               int c0, int c1, int c2) {
    if (CG == 1) {
        tma_load_3d(dst, m, bar, c0, c1, c2);
    } else {
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
            ".cta_group::2 [%0], [%1, {%3, %4, %5}], [%2];"
            ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
            : "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tma_load_4d_cg(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                               int c, int w, int h, int n) {
    if (CG == 1) {
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n)
            : "memory");
    } else {
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
            ".cta_group::2 [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n)
            : "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tma_load_im2col_cg(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                                   int c, int w, int h, int n, uint16_t offw,
                                                   uint16_t offh) {
    if (CG == 1) {
        tma_load_im2col(dst, m, bar, c, w, h, n, offw, offh);
    } else {
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
            ".cta_group::2 [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
            ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n),
            "h"(offw), "h"(offh)
            : "memory");
    }
}
template <uint32_t kCols, int CG>
__device__ __forceinline__ void tmem_alloc_cg(uint32_t smem_dst) {
    if (CG == 1) {
        tmem_alloc<kCols>(smem_dst);
    } else {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
                     "n"(kCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
}
template <uint32_t kCols, int CG>
__device__ __forceinline__ void tmem_dealloc_cg(uint32_t taddr) {
    if (CG == 1) {
        tmem_dealloc<kCols>(taddr);
    } else {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
                     : "memory");
    }
}
template <int CG>
__device__ __forceinline__ void umma_bf16_cg(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
    if (CG == 1) {
        umma_bf16(tmem_d, desc_a, desc_b, idesc, accumulate);
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// CG == 2: the arrival is multicast to the barrier at this offset in BOTH CTAs of the pair.
template <int CG>
__device__ __forceinline__ void umma_commit_cg(uint32_t bar) {
    if (CG == 1) {
        umma_commit(bar);
    } else {
        asm volatile(
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64"
            " [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
            : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// UMMA descriptors (bit layout: cute/arch/mma_sm100_desc.hpp of CUTLASS, restated)
// ---------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, 128-byte swizzle.
//   [0,14)  start address >> 4      [16,30) leading byte offset >> 4
//   [32,46) stride byte offset >> 4 [46,48) version = 1 (Blackwell)
//   [61,64) layout type: 2 = SWIZZLE_128B
// K-major tile [rows][64 bf16]: rows are 128 B apart inside an 8-row (1024 B) swizzle atom,
//   SBO = distance between 8-row groups, LBO unused.
// MN-major tile [k][64 bf16]: 64 MN-elements contiguous (128 B), k rows 128 B apart,
//   SBO = distance between 8-k groups (1024 B), LBO = distance between 64-element MN chunks.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32.
//   [4,6) c_format=1 (F32)  [7,10) a_format=1 (BF16)  [10,13) b_format=1 (BF16)
//   [15] a_major  [16] b_major (0 = K-major, 1 = MN-major)  [17,23) N>>3  [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N, uint32_t a_mn,
                                                       uint32_t b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) |
           ((M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// Warp-level column reduction: every lane holds v[0..31] (one row, 32 columns). Afterwards lane l
// holds in v[0] the sum over the 32 lanes (rows) of column l. 31 shuffles instead of 160.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_column_sum(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            float send = up ? v[i] : v[i + off];
            float keep = up ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

// Packed fp32 pairs (sm_100 FFMA2 / FADD2): one issue slot for two IEEE fp32 operations, each lane
// rounded exactly like fmaf / + on its own. A plain FFMA with three register operands issues every
// second cycle per scheduler; the CUDA-core kernels whose inner loop is a 9-tap fp32 convolution
// (first_conv.cuh) are bound by that, not by HBM. ptxas folds a {s, s} pair into the broadcast
// operand form (`R.F32`), so a scalar multiplicand costs no extra move.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(d)
        : "l"(*reinterpret_cast<const uint64_t*>(&a)), "l"(*reinterpret_cast<const uint64_t*>(&b)),
          "l"(*reinterpret_cast<const uint64_t*>(&c)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 ffma2(float s, float2 b, float2 c) {
    return ffma2(make_float2(s, s), b, c);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;"
        : "=l"(d)
        : "l"(*reinterpret_cast<const uint64_t*>(&a)), "l"(*reinterpret_cast<const uint64_t*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}

}  // namespace ub
