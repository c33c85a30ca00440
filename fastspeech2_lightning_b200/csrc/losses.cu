// Loss reductions and small elementwise helpers.
//   masked MSE / MAE means over the whole padded tensor      fs2/loss.py:44-106
//   attention binarisation loss                               fs2/attn/attention_loss.py:65-73
#include "common.cuh"

namespace fs2k {

// Σ f(pred·m − target·m) accumulated in fp64.  pred/target [M,C], m = row_mask[M] (1 = valid).
// TARGET_LOG1P: target is int32 and enters as log(target + 1)  (duration loss, loss.py:79)
template <bool TARGET_LOG1P>
__global__ void __launch_bounds__(256)
masked_loss_fwd_kernel(const float* __restrict__ pred, const void* __restrict__ target,
                       const uint8_t* __restrict__ row_mask, long M, int C, int kind /*0 mse, 1 mae*/,
                       double* __restrict__ sum) {
    pdl_prologue();
    const long N = M * C;
    double acc = 0.0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long)gridDim.x * blockDim.x) {
        const float m = row_mask[i / C] ? 1.f : 0.f;
        float t = TARGET_LOG1P ? logf((float)reinterpret_cast<const int*>(target)[i] + 1.0f)
                               : reinterpret_cast<const float*>(target)[i];
        const float d = pred[i] * m - t * m;
        acc += kind == 0 ? (double)(d * d) : (double)fabsf(d);
    }
    acc = warp_sum_d(acc);
    __shared__ double s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) t += s[w];
        atomicAdd(sum, t);
    }
}

// loss = sum · weight / N
__global__ void scalar_finalize_kernel(const double* __restrict__ sum, double mul, float* __restrict__ out) {
    pdl_prologue();
    out[0] = (float)(sum[0] * mul);
}

// dpred = g · weight/N · f'(pred·m − target·m) · m
template <bool TARGET_LOG1P>
__global__ void __launch_bounds__(256)
masked_loss_bwd_kernel(const float* __restrict__ pred, const void* __restrict__ target,
                       const uint8_t* __restrict__ row_mask, long M, int C, int kind, float coef,
                       const float* __restrict__ gout, float* __restrict__ dpred) {
    pdl_prologue();
    const long N = M * C;
    const float g = gout[0] * coef;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long)gridDim.x * blockDim.x) {
        const float m = row_mask[i / C] ? 1.f : 0.f;
        float t = TARGET_LOG1P ? logf((float)reinterpret_cast<const int*>(target)[i] + 1.0f)
                               : reinterpret_cast<const float*>(target)[i];
        const float d = pred[i] * m - t * m;
        const float fp = kind == 0 ? 2.f * d : (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
        dpred[i] = g * fp * m;
    }
}

// sums[0] += Σ_{hard==1} log(max(soft, eps)) ; sums[1] += Σ hard
__global__ void __launch_bounds__(256)
bin_loss_fwd_kernel(const float* __restrict__ hard, const float* __restrict__ soft, long N, float eps,
                    double* __restrict__ sums) {
    pdl_prologue();
    double a = 0.0, c = 0.0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long)gridDim.x * blockDim.x) {
        const float h = hard[i];
        if (h == 1.0f) a += (double)logf(fmaxf(soft[i], eps));
        c += (double)h;
    }
    a = warp_sum_d(a);
    c = warp_sum_d(c);
    __shared__ double s[2][8];
    if ((threadIdx.x & 31) == 0) { s[0][threadIdx.x >> 5] = a; s[1][threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0, tc = 0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) { ta += s[0][w]; tc += s[1][w]; }
        atomicAdd(&sums[0], ta);
        atomicAdd(&sums[1], tc);
    }
}
__global__ void bin_loss_finalize_kernel(const double* __restrict__ sums, float* __restrict__ out) {
    pdl_prologue();
    out[0] = (float)(-sums[0] / sums[1]);
}
// dsoft = −g / Σhard · [hard==1 ∧ soft ≥ eps] / soft     (clamp passes no gradient below eps)
__global__ void __launch_bounds__(256)
bin_loss_bwd_kernel(const float* __restrict__ hard, const float* __restrict__ soft, long N, float eps,
                    const double* __restrict__ sums, const float* __restrict__ gout, float* __restrict__ dsoft) {
    pdl_prologue();
    const float g = -gout[0] / (float)sums[1];
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long)gridDim.x * blockDim.x) {
        const float s = soft[i];
        dsoft[i] = (hard[i] == 1.0f && s >= eps) ? g / s : 0.f;
    }
}

// out = a·alpha + b·beta
__global__ void __launch_bounds__(256)
axpby_kernel(const float* __restrict__ a, float alpha, const float* __restrict__ b, float beta, long N,
             float* __restrict__ out) {
    pdl_prologue();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long)gridDim.x * blockDim.x)
        out[i] = a[i] * alpha + (b ? b[i] * beta : 0.f);
}

// out[r,:] = table[ids[r],:]
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ table, const long long* __restrict__ ids, long R, int D,
                   float* __restrict__ out) {
    pdl_prologue();
    const long N = R * D;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long)gridDim.x * blockDim.x)
        out[i] = table[(size_t)ids[i / D] * D + (i % D)];
}

__global__ void tanh_kernel(const float* __restrict__ x, long N, float* __restrict__ y) {
    pdl_prologue();
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) y[i] = tanhf(x[i]);
}

}  // namespace fs2k

using namespace fs2k;

static inline int ew_grid(long N) {
    long g = (N + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    return (int)(g < 1 ? 1 : g);
}

extern "C" int fs2k_masked_loss_fwd(const float* pred, const void* target, int target_is_int_log1p,
                                    const uint8_t* row_mask, long M, int C, int kind, float weight, double* scratch,
                                    float* loss, fs2k_stream_t stream) {
    FS2K_REQUIRE(M >= 0 && C > 0 && (kind == 0 || kind == 1), FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(pred && target && row_mask && scratch && loss, FS2K_ERR_NULL);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(double), s);
    if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    const long N = M * C;
    if (N > 0) {
        if (target_is_int_log1p) fs2k_launch(masked_loss_fwd_kernel<true>, dim3(ew_grid(N)), dim3(256), 0, s, pred, target, row_mask, M, C, kind, scratch);
        else fs2k_launch(masked_loss_fwd_kernel<false>, dim3(ew_grid(N)), dim3(256), 0, s, pred, target, row_mask, M, C, kind, scratch);
        FS2K_CHECK_LAUNCH();
    }
    fs2k_launch(scalar_finalize_kernel, dim3(1), dim3(1), 0, s, scratch, N > 0 ? (double)weight / (double)N : 0.0, loss);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_masked_loss_bwd(const float* pred, const void* target, int target_is_int_log1p,
                                    const uint8_t* row_mask, long M, int C, int kind, float weight, const float* gout,
                                    float* dpred, fs2k_stream_t stream) {
    FS2K_REQUIRE(M >= 0 && C > 0 && (kind == 0 || kind == 1), FS2K_ERR_BAD_SHAPE);
    const long N = M * C;
    if (N == 0) return FS2K_OK;
    FS2K_REQUIRE(pred && target && row_mask && gout && dpred, FS2K_ERR_NULL);
    cudaStream_t s = (cudaStream_t)stream;
    const float coef = weight / (float)N;
    if (target_is_int_log1p) fs2k_launch(masked_loss_bwd_kernel<true>, dim3(ew_grid(N)), dim3(256), 0, s, pred, target, row_mask, M, C, kind, coef, gout, dpred);
    else fs2k_launch(masked_loss_bwd_kernel<false>, dim3(ew_grid(N)), dim3(256), 0, s, pred, target, row_mask, M, C, kind, coef, gout, dpred);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_bin_loss_fwd(const float* hard, const float* soft, long N, float eps, double* sums /*[2]*/,
                                 float* loss, fs2k_stream_t stream) {
    FS2K_REQUIRE(N >= 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(hard && soft && sums && loss, FS2K_ERR_NULL);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(sums, 0, 2 * sizeof(double), s);
    if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    if (N > 0) {
        fs2k_launch(bin_loss_fwd_kernel, dim3(ew_grid(N)), dim3(256), 0, s, hard, soft, N, eps, sums);
        FS2K_CHECK_LAUNCH();
    }
    fs2k_launch(bin_loss_finalize_kernel, dim3(1), dim3(1), 0, s, sums, loss);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_bin_loss_bwd(const float* hard, const float* soft, long N, float eps, const double* sums,
                                 const float* gout, float* dsoft, fs2k_stream_t stream) {
    FS2K_REQUIRE(N >= 0, FS2K_ERR_BAD_SHAPE);
    if (N == 0) return FS2K_OK;
    FS2K_REQUIRE(hard && soft && sums && gout && dsoft, FS2K_ERR_NULL);
    fs2k_launch(bin_loss_bwd_kernel, dim3(ew_grid(N)), dim3(256), 0, (cudaStream_t)stream, hard, soft, N, eps, sums, gout, dsoft);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_axpby(const float* a, float alpha, const float* b, float beta, long N, float* out,
                          fs2k_stream_t stream) {
    FS2K_REQUIRE(N >= 0, FS2K_ERR_BAD_SHAPE);
    if (N == 0) return FS2K_OK;
    FS2K_REQUIRE(a && out, FS2K_ERR_NULL);
    fs2k_launch(axpby_kernel, dim3(ew_grid(N)), dim3(256), 0, (cudaStream_t)stream, a, alpha, b, beta, N, out);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_gather_rows(const float* table, const long long* ids, long R, int D, float* out,
                                fs2k_stream_t stream) {
    FS2K_REQUIRE(R >= 0 && D > 0, FS2K_ERR_BAD_SHAPE);
    if (R == 0) return FS2K_OK;
    FS2K_REQUIRE(table && ids && out, FS2K_ERR_NULL);
    fs2k_launch(gather_rows_kernel, dim3(ew_grid(R * D)), dim3(256), 0, (cudaStream_t)stream, table, ids, R, D, out);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_tanh(const float* x, long N, float* y, fs2k_stream_t stream) {
    FS2K_REQUIRE(N >= 0, FS2K_ERR_BAD_SHAPE);
    if (N == 0) return FS2K_OK;
    FS2K_REQUIRE(x && y, FS2K_ERR_NULL);
    fs2k_launch(tanh_kernel, dim3(cdiv(N, 256)), dim3(256), 0, (cudaStream_t)stream, x, N, y);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}
