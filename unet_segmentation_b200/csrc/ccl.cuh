// Connected-component labelling of a binary mask, bit-exact with the reference's
// get_instance_masks (utils/metrics.py:62-72): skimage.measure.label(connectivity=2) numbers the
// 8-connected components 1..K in raster order of their first pixel, remove_small_objects zeroes
// components with fewer than min_size pixels WITHOUT compacting the surviving ids, and the result
// is cast to uint16 (wrapping).
//
// GPU algorithm: label-equivalence union-find (roots = minimum linear index of a component), then
// raster-order renumbering = exclusive prefix sum over "is root", area histogram by root, filter.
#pragma once
#include "common.cuh"

namespace ub {

// Representative of i with intermediate pointer jumping (path halving, as in ECL-CC): every step
// re-points the node behind to its grand-parent, so repeated finds over the long chains that
// raster-order hooking produces stay cheap. The plain stores race benignly with the atomicMin hooks:
// a parent pointer is only ever replaced by an ancestor with a smaller index.
__device__ __forceinline__ int ccl_find(int* L, int i) {
    int cur = L[i];
    if (cur != i) {
        int prev = i, next;
        while (cur > (next = L[cur])) {
            L[prev] = next;
            prev = cur;
            cur = next;
        }
    }
    return cur;
}
// Read-only find for the flatten pass: there every thread stores the ROOT into its own L[i], so no
// other thread may re-point L[i] to a mere ancestor (which the halving stores above would do).
__device__ __forceinline__ int ccl_find_ro(const int* L, int i) {
    int r = L[i];
    while (r != L[r]) r = L[r];
    return r;
}
__device__ __forceinline__ void ccl_union(int* L, int a, int b) {
    bool done;
    do {
        a = ccl_find(L, a);
        b = ccl_find(L, b);
        if (a < b) {
            const int old = atomicMin(&L[b], a);
            done = (old == b);
            b = old;
        } else if (b < a) {
            const int old = atomicMin(&L[a], b);
            done = (old == a);
            a = old;
        } else {
            done = true;
        }
    } while (!done);
}

// Lane of the first pixel of the horizontal run that contains this lane's pixel, restricted to the
// 32 consecutive pixels of the warp and to the pixel's own image row (x = column of the pixel).
__device__ __forceinline__ int ccl_run_start_lane(unsigned bits, int lane, int x) {
    const unsigned below = (lane == 0) ? 0u : (0xffffffffu >> (32 - lane));   // lanes < lane
    const unsigned zeros = ~bits & below;                                     // background lanes below
    int start = zeros ? 32 - __clz(zeros) : 0;                                // one past the highest
    const int row_first = lane - x;                                           // lane of column 0
    return start > row_first ? start : row_first;
}

// L[i] = linear index of the first pixel of i's in-warp run (horizontal links inside a warp need no
// union afterwards), -1 for background. Whole warps walk 32 consecutive pixels.
static __global__ void ccl_init_kernel(const unsigned char* __restrict__ mask, int* __restrict__ L,
                                int* __restrict__ area, long long n, int W) {
    const int lane = threadIdx.x & 31;
    const long long nwarp32 = (n + 31) / 32 * 32;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nwarp32;
         i += (long long)gridDim.x * blockDim.x) {
        const bool fg = i < n && mask[i] != 0;
        const unsigned bits = __ballot_sync(0xffffffffu, fg);
        if (i < n) {
            int v = -1;
            if (fg) v = (int)i - (lane - ccl_run_start_lane(bits, lane, (int)(i % W)));
            L[i] = v;
            area[i] = 0;
        }
    }
}
// 8-connectivity with the half stencil W, NW, N, NE, pruned to the unions that can change the
// partition: the W link only where an in-warp run starts; the N link is implied when W and NW are
// both foreground (the pixel to the left carries it); NW / NE only when N is background (otherwise
// the upper row's own horizontal links connect them to N).
static __global__ void ccl_merge_kernel(const unsigned char* __restrict__ mask, int* L, int H, int W) {
    const long long n = (long long)H * W;
    const int lane = threadIdx.x & 31;
    const long long nwarp32 = (n + 31) / 32 * 32;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nwarp32;
         i += (long long)gridDim.x * blockDim.x) {
        const bool fg = i < n && mask[i] != 0;
        const unsigned bits = __ballot_sync(0xffffffffu, fg);
        if (!fg) continue;
        const int y = (int)(i / W), x = (int)(i % W);
        const bool west = x > 0 && mask[i - 1];
        if (west && ccl_run_start_lane(bits, lane, x) == lane) ccl_union(L, (int)i, (int)i - 1);
        if (y > 0) {
            const bool nw = x > 0 && mask[i - W - 1];
            if (mask[i - W]) {
                if (!(west && nw)) ccl_union(L, (int)i, (int)i - W);
            } else {
                if (nw) ccl_union(L, (int)i, (int)i - W - 1);
                if (x + 1 < W && mask[i - W + 1]) ccl_union(L, (int)i, (int)i - W + 1);
            }
        }
    }
}
// Flatten to roots, count areas, and count roots per 1024-pixel block (for the scan). Areas are
// aggregated per CTA in a small shared-memory hash (root -> count) before touching global memory: a
// cell that covers a fifth of an 8192 x 8192 mask would otherwise take 10^7 atomics on one address.
constexpr int CCL_HASH = 512;
static __global__ void __launch_bounds__(1024)
ccl_flatten_kernel(int* L, int* area, int* __restrict__ block_roots, long long n) {
    __shared__ int hk[CCL_HASH], hv[CCL_HASH];
    for (int k = threadIdx.x; k < CCL_HASH; k += 1024) { hk[k] = -1; hv[k] = 0; }
    __syncthreads();
    const long long i = (long long)blockIdx.x * 1024 + threadIdx.x;
    int is_root = 0;
    int r = -1;
    if (i < n && L[i] >= 0) {
        r = ccl_find_ro(L, (int)i);
        L[i] = r;
        is_root = (r == (int)i);
    }
    // warp aggregation: lanes with the same root elect one leader that carries their count
    const unsigned peers = __match_any_sync(0xffffffffu, r);
    if (r >= 0 && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) {
        const int cnt = __popc(peers);
        unsigned slot = ((unsigned)r * 2654435761u) >> 23;     // 9 bits
        bool placed = false;
        for (int probe = 0; probe < 16 && !placed; ++probe) {
            const int old = atomicCAS(&hk[slot], -1, r);
            if (old == -1 || old == r) { atomicAdd(&hv[slot], cnt); placed = true; }
            else slot = (slot + 1) & (CCL_HASH - 1);
        }
        if (!placed) atomicAdd(&area[r], cnt);
    }
    const int cnt_roots = __syncthreads_count(is_root);
    for (int k = threadIdx.x; k < CCL_HASH; k += 1024)
        if (hk[k] >= 0) atomicAdd(&area[hk[k]], hv[k]);
    if (threadIdx.x == 0) block_roots[blockIdx.x] = cnt_roots;
}
// Exclusive scan of block_roots, three launches, every level parallel (an 8192 x 8192 mask has 65 536
// blocks of 1024 pixels; the single-block serial version walked them in 64 dependent rounds):
//   1. ccl_scan_chunks_kernel: each CTA scans one chunk of 1024 entries in place (exclusive) and
//      writes the chunk total;   2. ccl_scan_kernel: ONE CTA scans the chunk totals (<= 1024 chunks
//      per round; a 2^31-pixel image has 2048);   3. ccl_scan_add_kernel: adds the chunk offsets.
__device__ __forceinline__ int ccl_block_scan_incl(int v, int* sm) {
    // inclusive scan over the 1024 threads of the CTA: warp shuffles + one shared-memory round
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += y;
    }
    if (lane == 31) sm[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int w = sm[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, w, off);
            if (lane >= off) w += y;
        }
        sm[lane] = w;
    }
    __syncthreads();
    const int res = x + (warp > 0 ? sm[warp - 1] : 0);
    __syncthreads();
    return res;
}
static __global__ void __launch_bounds__(1024)
ccl_scan_chunks_kernel(int* block_roots, int nblocks, int* __restrict__ chunk_tot) {
    __shared__ int sm[32];
    const int idx = blockIdx.x * 1024 + threadIdx.x;
    const int v = idx < nblocks ? block_roots[idx] : 0;
    const int incl = ccl_block_scan_incl(v, sm);
    if (idx < nblocks) block_roots[idx] = incl - v;
    if (threadIdx.x == 1023) chunk_tot[blockIdx.x] = incl;
}
static __global__ void __launch_bounds__(1024)
ccl_scan_kernel(int* vals, int n) {
    __shared__ int sm[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int idx = base + threadIdx.x;
        const int v = idx < n ? vals[idx] : 0;
        const int incl = ccl_block_scan_incl(v, sm);
        const int c = carry;
        if (idx < n) vals[idx] = c + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + incl;
        __syncthreads();
    }
}
static __global__ void __launch_bounds__(1024)
ccl_scan_add_kernel(int* block_roots, int nblocks, const int* __restrict__ chunk_off) {
    const int idx = blockIdx.x * 1024 + threadIdx.x;
    if (idx < nblocks) block_roots[idx] += chunk_off[blockIdx.x];
}
// rank[root] = raster-order ordinal of the root (1-based).
static __global__ void __launch_bounds__(1024)
ccl_rank_kernel(const int* __restrict__ L, const int* __restrict__ block_roots,
                int* __restrict__ rank, long long n) {
    __shared__ int sm[32];
    const long long i = (long long)blockIdx.x * 1024 + threadIdx.x;
    const int is_root = (i < n && L[i] == (int)i) ? 1 : 0;
    const int incl = ccl_block_scan_incl(is_root, sm);
    if (is_root) rank[i] = block_roots[blockIdx.x] + incl;
}
static __global__ void ccl_emit_kernel(const int* __restrict__ L, const int* __restrict__ area,
                                const int* __restrict__ rank, int min_size,
                                unsigned short* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const int r = L[i];
        unsigned short v = 0;
        if (r >= 0 && area[r] >= min_size) v = (unsigned short)(rank[r] & 0xFFFF);
        out[i] = v;
    }
}

}  // namespace ub
