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

__device__ __forceinline__ int ccl_find(const int* L, int i) {
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

static __global__ void ccl_init_kernel(const unsigned char* __restrict__ mask, int* __restrict__ L,
                                int* __restrict__ area, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        L[i] = mask[i] ? (int)i : -1;
        area[i] = 0;
    }
}
// Each foreground pixel merges with its W, NW, N, NE neighbours (8-connectivity, half stencil).
static __global__ void ccl_merge_kernel(const unsigned char* __restrict__ mask, int* L, int H, int W) {
    const long long n = (long long)H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        if (!mask[i]) continue;
        const int y = (int)(i / W), x = (int)(i % W);
        if (x > 0 && mask[i - 1]) ccl_union(L, (int)i, (int)i - 1);
        if (y > 0) {
            if (mask[i - W]) ccl_union(L, (int)i, (int)i - W);
            if (x > 0 && mask[i - W - 1]) ccl_union(L, (int)i, (int)i - W - 1);
            if (x + 1 < W && mask[i - W + 1]) ccl_union(L, (int)i, (int)i - W + 1);
        }
    }
}
// Flatten to roots, count areas, and count roots per 1024-pixel block (for the scan).
static __global__ void __launch_bounds__(1024)
ccl_flatten_kernel(int* L, int* area, int* __restrict__ block_roots, long long n) {
    const long long i = (long long)blockIdx.x * 1024 + threadIdx.x;
    int is_root = 0;
    if (i < n && L[i] >= 0) {
        const int r = ccl_find(L, (int)i);
        L[i] = r;
        atomicAdd(&area[r], 1);
        is_root = (r == (int)i);
    }
    const int cnt = __syncthreads_count(is_root);
    if (threadIdx.x == 0) block_roots[blockIdx.x] = cnt;
}
// Exclusive scan of block_roots (single block; nblocks up to a few hundred thousand).
static __global__ void __launch_bounds__(1024)
ccl_scan_kernel(int* block_roots, int nblocks) {
    __shared__ int sm[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        const int idx = base + threadIdx.x;
        const int v = idx < nblocks ? block_roots[idx] : 0;
        sm[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            int t = 0;
            if ((int)threadIdx.x >= off) t = sm[threadIdx.x - off];
            __syncthreads();
            sm[threadIdx.x] += t;
            __syncthreads();
        }
        const int incl = sm[threadIdx.x];
        const int c = carry;
        if (idx < nblocks) block_roots[idx] = c + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + incl;
        __syncthreads();
    }
}
// rank[root] = raster-order ordinal of the root (1-based).
static __global__ void __launch_bounds__(1024)
ccl_rank_kernel(const int* __restrict__ L, const int* __restrict__ block_roots,
                int* __restrict__ rank, long long n) {
    __shared__ int sm[1024];
    const long long i = (long long)blockIdx.x * 1024 + threadIdx.x;
    const int is_root = (i < n && L[i] == (int)i) ? 1 : 0;
    sm[threadIdx.x] = is_root;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        int t = 0;
        if ((int)threadIdx.x >= off) t = sm[threadIdx.x - off];
        __syncthreads();
        sm[threadIdx.x] += t;
        __syncthreads();
    }
    if (is_root) rank[i] = block_roots[blockIdx.x] + sm[threadIdx.x];
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
