// Device-side elastic deformation: elastic_deform_image_and_mask of the
// reference (utils/augmentations.py:4-39; per sample in utils/dataset.py:83-94 with alpha = 2000,
// sigma = 20, scripts/train.py:34-36) for a whole batch.
//
// The reference's arithmetic lives in scipy.ndimage (gaussian_filter, map_coordinates); the steps
// below restate it in scipy's operation order (oracle/elastic_ref.py spells them out and is pinned
// against the live reference), all in float64 with explicit round-to-nearest adds / multiplies so
// that the compiler cannot contract them into FMAs — results are bit-exact:
//   1. field = 2 u - 1 from the caller's uniform draws u (host numpy RandomState for parity with
//      the reference's seeds, or a device generator);
//   2. separable Gaussian, axis 0 (rows) then axis 1 (columns), zero outside the image, symmetric
//      accumulation acc = f[c] w[c]; acc += (f[c-d] + f[c+d]) w[d] for d = r … 1; the second pass
//      multiplies by alpha;
//   3. per pixel: (y + dy, x + dx) folded back into the image ('reflect'), image sampled
//      bilinearly and rounded half up into uint8, instance mask sampled at the nearest pixel.
// The Gaussian passes stage a tile plus its 2 r halo rows / columns in shared memory (each input
// element is read from HBM ~once instead of 2 r + 1 times) and every thread slides two 4-element
// register windows along the filter axis (2 shared-memory loads per 4 outputs and tap instead of
// 8); taps come from the host (numpy's exp).
#pragma once
#include "common.cuh"

namespace ub {

constexpr int EL_V_ROWS = 64;    // row pass: output rows per CTA (x 32 columns)
constexpr int EL_H_COLS = 128;   // column pass: output columns per CTA (x 32 rows)
constexpr int EL_H_ROWS = 32;
constexpr int EL_H_PITCH = 33;   // transposed tile [column][row], padded against bank conflicts
constexpr int EL_RUN = 4;        // consecutive outputs per thread along the filter axis

__host__ __device__ inline size_t elastic_v_smem(int r) { return ((size_t)(EL_V_ROWS + 2 * r) * 32 + r + 1) * 8; }
__host__ __device__ inline size_t elastic_h_smem(int r) { return ((size_t)(EL_H_COLS + 2 * r) * EL_H_PITCH + r + 1) * 8; }

// EL_RUN consecutive outputs along the filter axis from a zero-padded shared-memory line:
//   acc[k] = x[k] w[r];  acc[k] += (x[k - d] + x[k + d]) w[r - d]  for d = r … 1   (scipy's order)
// c0 points at x[0], consecutive positions are `stride` doubles apart, w[r - d] is the weight of
// distance d. The two operand windows x[k - d] and x[k + d] (k = 0 … 3) move by one position per
// step, so they live in two 4-register rings and each step loads 2 new values instead of 8 (the
// first cut read every operand from shared memory and ran at 88-95 % of the shared-memory
// wavefront peak with the FP64 pipe at 28 %, profiles/r01_summary.md). Ring slot of logical
// element k after s steps: (k + s) & 3 on the left, (k - s) & 3 on the right.
template <int S>
__device__ __forceinline__ void el_blur_step(const double* c0, int stride, double wd, int d,
                                             double (&acc)[EL_RUN], double (&L)[EL_RUN],
                                             double (&R)[EL_RUN]) {
#pragma unroll
    for (int k = 0; k < EL_RUN; ++k)
        acc[k] = __dadd_rn(acc[k], __dmul_rn(__dadd_rn(L[(k + S) & 3], R[(k - S) & 3]), wd));
    // windows of distance d - 1: the left one gains x[3 - (d - 1)], the right one x[d - 1]
    L[S & 3] = c0[(EL_RUN - d) * stride];
    R[(3 - S) & 3] = c0[(d - 1) * stride];
}
__device__ __forceinline__ void el_blur_run(const double* c0, int stride, const double* w, int r,
                                            double (&acc)[EL_RUN]) {
    static_assert(EL_RUN == 4, "ring indexing assumes runs of 4");
    double L[EL_RUN], R[EL_RUN];
#pragma unroll
    for (int k = 0; k < EL_RUN; ++k) {
        acc[k] = __dmul_rn(c0[k * stride], w[r]);
        L[k] = c0[(k - r) * stride];
        R[k] = c0[(k + r) * stride];
    }
    int d = r;
    for (; d >= 4; d -= 4) {
        el_blur_step<0>(c0, stride, w[r - d], d, acc, L, R);
        el_blur_step<1>(c0, stride, w[r - d + 1], d - 1, acc, L, R);
        el_blur_step<2>(c0, stride, w[r - d + 2], d - 2, acc, L, R);
        el_blur_step<3>(c0, stride, w[r - d + 3], d - 3, acc, L, R);
    }
    if (d >= 1) el_blur_step<0>(c0, stride, w[r - d], d, acc, L, R);
    if (d >= 2) el_blur_step<1>(c0, stride, w[r - d + 1], d - 1, acc, L, R);
    if (d >= 3) el_blur_step<2>(c0, stride, w[r - d + 2], d - 2, acc, L, R);
}

// Axis-0 pass. in / out: [images][H][W]; block (32, 8); grid (ceil(W/32), ceil(H/EL_V_ROWS), images).
// taps[0..r] = weights of distance r … 0 (the first half of the symmetric kernel). Thread (tx, ty)
// computes column tx of the tile, rows ty * 8 … ty * 8 + 7 as two runs of 4.
static __global__ void __launch_bounds__(256)
elastic_blur_rows_kernel(const double* __restrict__ in, double* __restrict__ out, int H, int W,
                         const double* __restrict__ taps, int r, int affine) {
    pdl_entry();
    extern __shared__ double el_smem[];
    double* tile = el_smem;                                   // [(EL_V_ROWS + 2r)][32]
    double* w = el_smem + (size_t)(EL_V_ROWS + 2 * r) * 32;   // [r + 1]
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int x = blockIdx.x * 32 + tx, y0 = blockIdx.y * EL_V_ROWS;
    const size_t img = (size_t)blockIdx.z * H * W;
    for (int i = ty * 32 + tx; i <= r; i += 256) w[i] = taps[i];
    for (int j = ty; j < EL_V_ROWS + 2 * r; j += 8) {
        const int y = y0 - r + j;
        double v = 0.0;
        if (y >= 0 && y < H && x < W) {
            v = in[img + (size_t)y * W + x];
            if (affine) v = __dadd_rn(__dmul_rn(v, 2.0), -1.0);
        }
        tile[j * 32 + tx] = v;
    }
    __syncthreads();
    if (x >= W) return;
#pragma unroll 1
    for (int g = 0; g < 8 / EL_RUN; ++g) {
        const int k0 = ty * 8 + g * EL_RUN;
        if (y0 + k0 >= H) break;
        double acc[EL_RUN];
        el_blur_run(tile + (k0 + r) * 32 + tx, 32, w, r, acc);
#pragma unroll
        for (int k = 0; k < EL_RUN; ++k)
            if (y0 + k0 + k < H) out[img + (size_t)(y0 + k0 + k) * W + x] = acc[k];
    }
}

// Axis-1 pass, result times alpha. block (32, 8); grid (ceil(W/EL_H_COLS), ceil(H/EL_H_ROWS), images).
// The tile is stored transposed ([column][row], pitch 33) so that the 32 lanes of a warp = 32 rows
// read consecutive words while each thread slides along its row: thread (tr, tg) computes row tr,
// columns tg * 16 … tg * 16 + 15 as four runs of 4.
static __global__ void __launch_bounds__(256)
elastic_blur_cols_kernel(const double* __restrict__ in, double* __restrict__ out, int H, int W,
                         const double* __restrict__ taps, int r, double alpha) {
    pdl_entry();
    extern __shared__ double el_smem[];
    const int cols = EL_H_COLS + 2 * r;
    double* tile = el_smem;                                   // [cols][EL_H_PITCH]
    double* w = el_smem + (size_t)cols * EL_H_PITCH;          // [r + 1]
    const int tr = threadIdx.x, tg = threadIdx.y;
    const int x0 = blockIdx.x * EL_H_COLS, y0 = blockIdx.y * EL_H_ROWS;
    const size_t img = (size_t)blockIdx.z * H * W;
    for (int i = tg * 32 + tr; i <= r; i += 256) w[i] = taps[i];
    for (int row = tg; row < EL_H_ROWS; row += 8) {           // a warp loads one row, coalesced
        const int y = y0 + row;
        for (int j = tr; j < cols; j += 32) {
            const int x = x0 - r + j;
            tile[j * EL_H_PITCH + row] = (y < H && x >= 0 && x < W) ? in[img + (size_t)y * W + x] : 0.0;
        }
    }
    __syncthreads();
    const int y = y0 + tr;
    if (y >= H) return;
#pragma unroll 1
    for (int g = 0; g < 16 / EL_RUN; ++g) {
        const int k0 = tg * 16 + g * EL_RUN;
        if (x0 + k0 >= W) break;
        double acc[EL_RUN];
        el_blur_run(tile + (k0 + r) * EL_H_PITCH + tr, EL_H_PITCH, w, r, acc);
#pragma unroll
        for (int k = 0; k < EL_RUN; ++k)
            if (x0 + k0 + k < W) out[img + (size_t)y * W + x0 + k0 + k] = __dmul_rn(acc[k], alpha);
    }
}

// scipy's 'reflect' fold of a coordinate (half-sample symmetric), applied for c < 0 and c > n - 1.
__device__ __forceinline__ double el_fold_coordinate(double c, int n) {
    if (c < 0.0) {
        if (n <= 1) return 0.0;
        const double period = 2.0 * (double)n;
        if (c < -period) c = __dadd_rn(__dmul_rn(period, trunc(__ddiv_rn(-c, period))), c);
        return c < -(double)n ? __dadd_rn(c, period) : __dadd_rn(-c, -1.0);
    }
    if (c > (double)(n - 1)) {
        if (n <= 1) return 0.0;
        const double period = 2.0 * (double)n;
        c = __dadd_rn(c, -__dmul_rn(period, trunc(__ddiv_rn(c, period))));
        return c >= (double)n ? __dadd_rn(__dadd_rn(period, -c), -1.0) : c;
    }
    return c;
}
// d c b a | a b c d | d c b a for integer footprint positions
__device__ __forceinline__ int el_fold_index(long long i, int n) {
    if ((unsigned long long)i < (unsigned long long)n) return (int)i;   // in range: the common case
    if (n <= 1) return 0;
    if (i < 0) i = -i - 1;
    i %= 2LL * n;
    return (int)(i >= n ? 2LL * n - 1 - i : i);
}

template <typename LabelIn, typename LabelOut>
static __global__ void __launch_bounds__(256)
elastic_sample_kernel(const unsigned char* __restrict__ image, const LabelIn* __restrict__ labels,
                      const double* __restrict__ dx, const double* __restrict__ dy, int N, int H, int W,
                      unsigned char* __restrict__ image_out, LabelOut* __restrict__ labels_out) {
    pdl_entry();
    const unsigned total = (unsigned)N * H * W, HW = (unsigned)H * W;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned n = i / HW, p = i - n * HW, yi = p / (unsigned)W, xi = p - yi * (unsigned)W;
        const double cy = el_fold_coordinate(__dadd_rn((double)yi, dy[i]), H);
        const double cx = el_fold_coordinate(__dadd_rn((double)xi, dx[i]), W);
        const size_t base = (size_t)n * HW;
        if (image) {
            const double fy = floor(cy), fx = floor(cx);
            const double wy0 = __dadd_rn(1.0, -__dadd_rn(cy, -fy)), wx0 = __dadd_rn(1.0, -__dadd_rn(cx, -fx));
            const double wy1 = __dadd_rn(1.0, -wy0), wx1 = __dadd_rn(1.0, -wx0);
            const int r0 = el_fold_index((long long)fy, H), r1 = el_fold_index((long long)fy + 1, H);
            const int c0 = el_fold_index((long long)fx, W), c1 = el_fold_index((long long)fx + 1, W);
            const unsigned char* im = image + base;
            const double f00 = (double)im[(size_t)r0 * W + c0], f01 = (double)im[(size_t)r0 * W + c1];
            const double f10 = (double)im[(size_t)r1 * W + c0], f11 = (double)im[(size_t)r1 * W + c1];
            double t = __dmul_rn(__dmul_rn(f00, wy0), wx0);
            t = __dadd_rn(t, __dmul_rn(__dmul_rn(f01, wy0), wx1));
            t = __dadd_rn(t, __dmul_rn(__dmul_rn(f10, wy1), wx0));
            t = __dadd_rn(t, __dmul_rn(__dmul_rn(f11, wy1), wx1));
            t = t > 0.0 ? __dadd_rn(t, 0.5) : 0.0;       // round half up, clamp into uint8
            t = t > 255.0 ? 255.0 : t;
            image_out[i] = (unsigned char)(int)t;
        }
        if (labels) {
            const int r = el_fold_index((long long)floor(__dadd_rn(cy, 0.5)), H);
            const int c = el_fold_index((long long)floor(__dadd_rn(cx, 0.5)), W);
            labels_out[i] = (LabelOut)labels[base + (size_t)r * W + c];   // uint16 -> uint8 wraps like astype
        }
    }
}

}  // namespace ub
