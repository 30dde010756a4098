// Internal C++ launch API shared by the op-level C ABI (capi_ops.cu) and the network executor
// (plan.cu). Every function enqueues work on `stream` and returns 0 or a negative ub error code;
// nothing synchronises the device.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <string.h>

#include "common.cuh"

namespace ub {

enum : int {
    UB_OK = 0,
    UB_ERR_CUDA = -1,
    UB_ERR_ARG = -2,
    UB_ERR_UNSUPPORTED = -3,
    UB_ERR_TMAP = -4,
    UB_ERR_NOMEM = -5,
};

void set_last_error(const char* fmt, ...);
const char* last_error();
int num_sms();
void count_launch(int n = 1);   // every kernel launch of the library is counted (bench.py gpu_launches)
long long launch_count();

#define UB_CHECK_CUDA(expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ub::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                               __LINE__);                                                     \
            return ub::UB_ERR_CUDA;                                                           \
        }                                                                                     \
    } while (0)
#define UB_POST_LAUNCH()                 \
    do {                                 \
        ub::count_launch(1);             \
        UB_CHECK_CUDA(cudaGetLastError()); \
    } while (0)
#define UB_TRY(expr)              \
    do {                          \
        int _r = (expr);          \
        if (_r != 0) return _r;   \
    } while (0)

// Kernel launch with programmatic dependent launch (see common.cuh: pdl_*) and an optional
// 2-CTA cluster (cta_group::2 pairs). UB_PDL=0 launches with plain stream serialization.
// pdl_allow(): may this launch start while its predecessor drains? (UB_PDL: 0 never, 1 always, 3 = only edges next to a tiny kernel,
// 2 = every edge except a tensor-core kernel following a large elementwise kernel, whose early
// resident CTAs (352 threads, ~47 K registers, 200 KB smem) would cut the elementwise occupancy.)
bool pdl_allow(bool gemm, unsigned blocks);
template <typename... KArgs, typename... Args>
inline cudaError_t ub_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                             cudaStream_t stream, int cluster, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    unsigned n = 0;
    if (cluster > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = (unsigned)cluster; attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl_allow(smem > 100 * 1024, grid.x * grid.y)) {   // the tensor-core kernels use > 100 KB smem
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr; cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// drop-in for `kernel<<<grid, block, smem, stream>>>(args...)`; UB_POST_LAUNCH() follows as before
#define UB_LAUNCH_NC(kernel, grid, block, smem, stream, ...)                                       \
    UB_CHECK_CUDA(ub::ub_launch(kernel, dim3(grid), dim3(block), (size_t)(smem), stream, 1, __VA_ARGS__))

// ---- igemm.cu -------------------------------------------------------------------------------
struct IgemmEpilogue {
    int kind;                 // EPI_*
    __nv_bfloat16* out;       // plain kinds: [M][ldo]
    long long ldo;
    const float* bias;
    const float* scale;
    const float* shift;
    float* stats;             // EPI_CONV_STATS partials; layout reported through stats_* below
    View ct_dst;              // EPI_CONVT destination view [N,2H,2W,>=Cout] (ptr at channel 0)
    // EPI_AFFINE_RELU_HEAD: 1x1 head weights [nc][64] / bias [nc]; logits NCHW fp32; optional mask
    const float* head_w;
    const float* head_b;
    float* head_logits;
    unsigned char* head_mask;
    int head_nc;
    // EPI_STORE_BNRED (data gradients): pre-BN output `red_y` ([M][ldo], the geometry of `out`) and batch
    // mean of the layer that receives this gradient; its scale / shift go in `scale` / `shift`, the
    // per-CTA partial sums (S1 = sum dyh, S2 = sum dyh (y - mean)) land in `stats` with the layout of the
    // forward statistics (IgemmLaunchInfo describes it; launch_bn_bwd(..., fused) consumes it).
    const __nv_bfloat16* red_y;
    const float* red_mean;
    // EPI_AFFINE_RELU: optional [N][Ho/2][Wo/2][ldo] destination of a fused 2x2 floor max-pool. The
    // launcher fuses it where the kernel selected for the shape can (two-row row-run tiles) and says so
    // in IgemmLaunchInfo::pool_fused; otherwise the caller runs launch_maxpool2.
    __nv_bfloat16* pooled;
};
struct IgemmLaunchInfo {
    int grid, n_tiles, BN, M;
    int pool_fused;
};
// A operand: im2col over src0 (and optionally src1 = second channel range of a zero-copy concat).
// B operand: packed weights [ncols][taps*(C0+C1)] bf16, K-major.
int launch_igemm(const View& src0, const View* src1, int lower, int upper, int tstride, int taps,
                 int tapw, const __nv_bfloat16* wB, int ncols, const IgemmEpilogue& epi,
                 IgemmLaunchInfo* info, cudaStream_t stream);
size_t igemm_stats_floats(int ncols);  // upper bound of the stats partial buffer

// Weight gradient. A side: im2col'd activations (rows = taps * (C0+C1)); B side: [Mpix][cols]
// matrix with row pitch ldb. Result (fp32, torch layout, see wgrad_reduce_kernel) in `out`.
// zero0 / zero1 (optional): float ranges cleared by the reduction kernel (analytically-zero bias
// gradients of the same backward stage).
int launch_wgrad(const View& src0, const View* src1, int lower, int upper, int tstride, int taps,
                 int tapw, const __nv_bfloat16* B, long long ldb, int cols, float* ws,
                 size_t ws_floats, float* out, cudaStream_t stream, float* zero0 = nullptr,
                 int nzero0 = 0, float* zero1 = nullptr, int nzero1 = 0);
size_t wgrad_ws_floats(int rows, int cols, long long mpix);

// ---- kernels.cu ------------------------------------------------------------------------------
int launch_pack_conv3x3(const float* w, int Co, int Ci, __nv_bfloat16* wf, __nv_bfloat16* wd,
                        cudaStream_t s);
int launch_pack_convT(const float* w, int Ci, int Co, __nv_bfloat16* wf, __nv_bfloat16* wb,
                      const float* bias, float* bias4, cudaStream_t s);
int launch_bn_finalize(const float* stats, const IgemmLaunchInfo& info, int C, double count,
                       const float* gamma, const float* beta, float* rm, float* rv,
                       long long* nbt, float momentum, float eps, float* scale, float* shift,
                       float* mean, float* rstd, cudaStream_t s);
int launch_bn_finalize_flat(const float* part, int blocks, int C, double count, const float* gamma,
                            const float* beta, float* rm, float* rv, long long* nbt,
                            float momentum, float eps, float* scale, float* shift, float* mean,
                            float* rstd, cudaStream_t s);
int launch_bn_fold_eval(int C, const float* conv_bias, const float* gamma, const float* beta,
                        const float* rm, const float* rv, float eps, float* scale, float* shift,
                        cudaStream_t s);
int launch_bn_apply_relu(const __nv_bfloat16* y, __nv_bfloat16* a, __nv_bfloat16* pooled,
                         unsigned char* amax, int N, int H, int W, int C, const float* scale,
                         const float* shift, cudaStream_t s);
bool bn_apply_head_supported(int C, int NC);
int launch_bn_apply_relu_head(const __nv_bfloat16* y, __nv_bfloat16* a, int N, int H, int W, int C,
                              const float* scale, const float* shift, int NC, const float* hw,
                              const float* hb, float* logits, cudaStream_t s);
int launch_maxpool2(const __nv_bfloat16* a, __nv_bfloat16* pooled, int N, int H, int W, int C,
                    cudaStream_t s);
struct BnBwdDesc {
    const __nv_bfloat16* y;
    int N, H, W, C;
    const float *scale, *shift, *mean, *rstd;
    bool pool_skip;
    View g;        // direct
    View gp;       // pooled grad
    View gs;       // skip grad
    int crop_h, crop_w;
    bool has_skip;
    const unsigned char* amax;  // pool arg-max saved by launch_bn_apply_relu (may be null)
    float* partial;        // workspace, >= bn_bwd_partial_floats(C)
    float* dgamma;         // out [C]
    float* dbeta;          // out [C]
    __nv_bfloat16* dy;     // out [N,H,W,C]
};
size_t bn_bwd_partial_floats(int C);
// `fused` (optional): the reduce pass already happened in the epilogue of the kernel that produced the
// upstream gradient (EPI_STORE_BNRED); its partial rows are in d.partial with the layout `fused`
// describes. Only the finalisation and the apply pass are launched then.
int launch_bn_bwd(const BnBwdDesc& d, cudaStream_t s, const IgemmLaunchInfo* fused = nullptr);

struct FirstConvDesc {
    const float* x;
    int N, Ci, H, W, Co;
    const float* w;
    const float* bias;
};
size_t first_conv_partial_floats(int Co);
constexpr int FIRST_CONV_COV_DOUBLES = 56;  // patch moments kept between forward and backward (Ci = 1)
// Train-mode BatchNorm statistics of the first conv (+ running-stat update). `cov` (may be null)
// receives the patch moments of a single-channel input for launch_first_conv_bwd.
int launch_first_conv_train_stats(const FirstConvDesc& d, float* ws, const float* gamma,
                                  const float* beta, float* rm, float* rv, long long* nbt,
                                  float momentum, float eps, float* scale, float* shift,
                                  float* mean, float* rstd, double* cov, cudaStream_t s);
int launch_first_conv_apply(const FirstConvDesc& d, const float* scale, const float* shift,
                            __nv_bfloat16* a, cudaStream_t s);
// `a` = the forward activation (needed for Ci = 1: ReLU mask), `cov` = moments saved by the forward
// (null: recomputed into the workspace).
int launch_first_conv_bwd(const FirstConvDesc& d, const float* scale, const float* shift,
                          const float* mean, const float* rstd, const View& g,
                          const __nv_bfloat16* a, const double* cov, float* ws, float* dgamma,
                          float* dbeta, float* dw, cudaStream_t s);

int launch_head_fwd(const __nv_bfloat16* a, int N, int H, int W, int K, int NC, const float* w,
                    const float* b, float* logits, unsigned char* mask, cudaStream_t s);
size_t head_bwd_partial_floats(int K, int NC);
int launch_head_bwd(const float* dlogits, const __nv_bfloat16* a, int N, int H, int W, int K,
                    int NC, const float* w, __nv_bfloat16* da, float* partial, float* dw, float* db,
                    cudaStream_t s);

struct WceDesc {
    const float* z; long long zs[4];
    const long long* t; long long ts[3];
    const float* w; long long ws[3];
    int N, C, H, W;
};
size_t wce_partial_floats();
int launch_wce(const WceDesc& d, float* loss, float* dz, float* partial, int* err, cudaStream_t s);
int launch_scale_by_scalar(const float* in, const float* scalar, float* out, long long n,
                           cudaStream_t s);
int launch_fill_zero(float* p, long long n, cudaStream_t s);

int launch_prepare_batch(const unsigned char* img, const void* labels, int label_bytes,
                         const void* wmap, int wmap_bytes, int N, int H, int W, int oh, int ow,
                         float* image, long long* target, float* weight, cudaStream_t s);
size_t ccl_ws_bytes(int H, int W);
int launch_upsample2x_fwd(const View& x, __nv_bfloat16* out, cudaStream_t s);
int launch_upsample2x_bwd(const View& g, __nv_bfloat16* dx, cudaStream_t s);
int launch_weight_map(const void* labels, int label_bytes, int N, int H, int W, double w0,
                      double sigma, void* out, int out_bytes, unsigned* counts, cudaStream_t s);
size_t elastic_ws_bytes(int N, int H, int W);
int launch_elastic(const unsigned char* img, const void* labels, int label_bytes, int N, int H, int W,
                   const double* noise, const double* taps, int radius, double alpha,
                   unsigned char* img_out, void* labels_out, int label_out_bytes, void* ws,
                   cudaStream_t s);
// overlap-tile helpers (tiling.cuh): origins_yx = T pairs (row, column) of output-space tile origins
// on the device; a negative row marks an unused slot
int launch_extract_tiles(const float* image, int H, int W, const int* origins_yx, int T, int S,
                         int margin, float* tiles, cudaStream_t s);
int launch_stitch_tiles(const unsigned char* tiles, const int* origins_yx, int T, int TO,
                        unsigned char* full, int H, int W, cudaStream_t s);
int launch_ccl(const unsigned char* mask, int H, int W, int min_size, unsigned short* out, void* ws,
               cudaStream_t s);

}  // namespace ub
