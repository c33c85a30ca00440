// Normalisation kernels: LayerNorm (row statistics), BatchNorm1d statistics / folding, and the
// per-channel affine + activation pass that follows a BatchNorm in training mode.
//   LayerNorm(256, eps 1e-5)   torchaudio conformer.py:41,103,151,165 ; fs2/layers.py:42
//   BatchNorm1d                torchaudio conformer.py:62-64 ; fs2/layers.py:168-202
// BatchNorm in training mode uses the biased variance of all B·L positions including padding
// (SURVEY §8a note P) and updates running stats with momentum 0.1 and the unbiased variance.
#include <cuda_bf16.h>

#include "common.cuh"

namespace fs2k {

// one warp per row, the row is held in registers (D ≤ 32·MAXV·4)
template <int MAXV>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 float eps, long M, int D, float drop_p, unsigned long long seed, float* __restrict__ y,
                 float* __restrict__ mean_out, float* __restrict__ rstd_out) {
    pdl_prologue();
    seed = seed_with_base(seed);
    const uint32_t thr16 = drop_thr16(drop_p);
    const float inv_keep = drop_inv_keep(thr16);
    const int lane = threadIdx.x & 31;
    const int D4 = D >> 2;
    for (long m = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); m < M;
         m += (long)gridDim.x * (blockDim.x >> 5)) {
        const float4* xr = reinterpret_cast<const float4*>(x + (size_t)m * D);
        float4 v[MAXV];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int q = lane + 32 * i;
            v[i] = q < D4 ? xr[q] : make_float4(0.f, 0.f, 0.f, 0.f);
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
        const float mean = warp_sum(s) / (float)D;
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int q = lane + 32 * i;
            if (q < D4) {
                const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
                ss += (a * a + b * b) + (c * c + d * d);
            }
        }
        const float var = warp_sum(ss) / (float)D;
        const float rstd = 1.0f / sqrtf(var + eps);
        if (lane == 0) {
            if (mean_out) mean_out[m] = mean;
            if (rstd_out) rstd_out[m] = rstd;
        }
        float4* yr = reinterpret_cast<float4*>(y + (size_t)m * D);
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int q = lane + 32 * i;
            if (q < D4) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + q);
                const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + q);
                float4 o;
                o.x = (v[i].x - mean) * rstd * g.x + b.x;
                o.y = (v[i].y - mean) * rstd * g.y + b.y;
                o.z = (v[i].z - mean) * rstd * g.z + b.z;
                o.w = (v[i].w - mean) * rstd * g.w + b.w;
                if (drop_p > 0.f) {
                    const unsigned long long e = (unsigned long long)m * D + q * 4;
                    drop_apply4(o, seed, e, thr16, inv_keep);
                }
                yr[q] = o;
            }
        }
    }
}

// per-channel Σx and Σx² over the M rows of z[M,C], accumulated in fp64.
// CTA = 256 threads = 8 row-groups × 32 channel lanes (float4 → 128 channels per pass).
__global__ void __launch_bounds__(256)
colstats_kernel(const float* __restrict__ z, long M, int C, long rows_per_cta, double* __restrict__ sums) {
    pdl_prologue();
    __shared__ double s_part[8][128][2];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const long m0 = (long)blockIdx.x * rows_per_cta;
    const long m1 = min(M, m0 + rows_per_cta);
    {
        const int c0 = blockIdx.y * 128;  // blockIdx.y = 128-channel block, blockIdx.x = row chunk
        const int c = c0 + lane * 4;
        float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
        double ds[4] = {0, 0, 0, 0}, dq[4] = {0, 0, 0, 0};
        int n = 0;
        if (c < C) {
#pragma unroll 4
            for (long m = m0 + grp; m < m1; m += 8) {
                const float4 v = *reinterpret_cast<const float4*>(z + (size_t)m * C + c);
                s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
                q[0] += v.x * v.x; q[1] += v.y * v.y; q[2] += v.z * v.z; q[3] += v.w * v.w;
                if (++n == 32) {  // flush fp32 partials into fp64 every 32 rows
#pragma unroll
                    for (int k = 0; k < 4; ++k) { ds[k] += s[k]; dq[k] += q[k]; s[k] = 0; q[k] = 0; }
                    n = 0;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            s_part[grp][lane * 4 + k][0] = ds[k] + s[k];
            s_part[grp][lane * 4 + k][1] = dq[k] + q[k];
        }
        __syncthreads();
        if (threadIdx.x < 128 && c0 + threadIdx.x < C) {
            double a = 0, b = 0;
#pragma unroll
            for (int g = 0; g < 8; ++g) { a += s_part[g][threadIdx.x][0]; b += s_part[g][threadIdx.x][1]; }
            atomicAdd(&sums[c0 + threadIdx.x], a);
            atomicAdd(&sums[C + c0 + threadIdx.x], b);
        }
        __syncthreads();
    }
}

// training: batch mean / biased var → (scale, shift) ; running stats ← momentum update (unbiased var)
// eval:     running stats → (scale, shift)
__global__ void bn_finalize_kernel(const double* __restrict__ sums, long M, int C, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, int training,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   long long* __restrict__ num_batches_tracked, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ save_mean,
                                   float* __restrict__ save_rstd) {
    pdl_prologue();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && training && num_batches_tracked) *num_batches_tracked += 1;
    if (c >= C) return;
    float mean, var;
    if (training) {
        const double mu = sums[c] / (double)M;
        double v = sums[C + c] / (double)M - mu * mu;
        if (v < 0) v = 0;
        mean = (float)mu;
        var = (float)v;
        if (running_mean) {
            const double unbiased = M > 1 ? v * (double)M / (double)(M - 1) : v;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
        }
    } else {
        mean = running_mean[c];
        var = running_var[c];
    }
    const float rstd = 1.0f / sqrtf(var + eps);
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    scale[c] = g * rstd;
    shift[c] = b - mean * g * rstd;
    if (save_mean) save_mean[c] = mean;
    if (save_rstd) save_rstd[c] = rstd;
}

// y = act(z·scale[c] + shift[c]) (+ residual)    act: 0 none, 1 relu, 2 silu, 3 tanh
template <bool OUT16>
__global__ void __launch_bounds__(256)
affine_act_kernel(const float* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
                  int act, const float* __restrict__ residual, long M, int C, float drop_p, unsigned long long seed,
                  void* __restrict__ y) {
    pdl_prologue();
    seed = seed_with_base(seed);
    const uint32_t thr16 = drop_thr16(drop_p);
    const float inv_keep = drop_inv_keep(thr16);
    const int C4 = C >> 2;
    const long N = M * C4;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long)gridDim.x * blockDim.x) {
        const int q = (int)(i % C4);
        const float4 v = reinterpret_cast<const float4*>(z)[i];
        const float4 sc = scale ? __ldg(reinterpret_cast<const float4*>(scale) + q) : make_float4(1.f, 1.f, 1.f, 1.f);
        const float4 sh = scale ? __ldg(reinterpret_cast<const float4*>(shift) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        float o[4] = {v.x * sc.x + sh.x, v.y * sc.y + sh.y, v.z * sc.z + sh.z, v.w * sc.w + sh.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (act == 1) o[k] = fmaxf(o[k], 0.f);
            else if (act == 2) o[k] = silu(o[k]);
            else if (act == 3) o[k] = tanhf(o[k]);
            if (drop_p > 0.f) o[k] = drop_keep_g(seed, (unsigned long long)i, k, thr16) ? o[k] * inv_keep : 0.f;
        }
        if (residual) {
            const float4 r = reinterpret_cast<const float4*>(residual)[i];
            o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
        }
        if (OUT16) {  // bf16 mode: the value only feeds a tensor-core contraction (its TMA reads bf16 tiles)
            const __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]), b = __floats2bfloat162_rn(o[2], o[3]);
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t*>(&a);
            pk.y = *reinterpret_cast<const uint32_t*>(&b);
            reinterpret_cast<uint2*>(y)[i] = pk;
        } else {
            reinterpret_cast<float4*>(y)[i] = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

}  // namespace fs2k

using namespace fs2k;

extern "C" int fs2k_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, long M, int D,
                                  float dropout_p, long seed, float* y, float* mean_out, float* rstd_out,
                                  fs2k_stream_t stream) {
    const unsigned long long useed = (unsigned long long)seed;
    FS2K_REQUIRE(M >= 0 && D > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE((D & 3) == 0 && D <= 1024, FS2K_ERR_UNSUPPORTED);
    if (M == 0) return FS2K_OK;
    FS2K_REQUIRE(x && gamma && beta && y, FS2K_ERR_NULL);
    long g = (M + 7) / 8;
    if (g > 148 * 8) g = 148 * 8;
    cudaStream_t s = (cudaStream_t)stream;
    if (D <= 256) fs2k_launch(layernorm_kernel<2>, dim3((int)g), dim3(256), 0, s, x, gamma, beta, eps, M, D, dropout_p, useed, y, mean_out, rstd_out);
    else if (D <= 512) fs2k_launch(layernorm_kernel<4>, dim3((int)g), dim3(256), 0, s, x, gamma, beta, eps, M, D, dropout_p, useed, y, mean_out, rstd_out);
    else fs2k_launch(layernorm_kernel<8>, dim3((int)g), dim3(256), 0, s, x, gamma, beta, eps, M, D, dropout_p, useed, y, mean_out, rstd_out);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_colstats(const float* z, long M, int C, double* sums /* [2*C], zeroed here */,
                             fs2k_stream_t stream) {
    FS2K_REQUIRE(M >= 0 && C > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE((C & 3) == 0, FS2K_ERR_UNSUPPORTED);
    FS2K_REQUIRE(sums && (z || M == 0), FS2K_ERR_NULL);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, s);
    if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    if (M == 0) return FS2K_OK;
    long rows_per_cta;
    const dim3 grid = col_reduce_grid(M, C, &rows_per_cta);
    fs2k_launch(colstats_kernel, dim3(grid), dim3(256), 0, s, z, M, C, rows_per_cta, sums);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_bn_finalize(const double* sums, long M, int C, const float* gamma, const float* beta, float eps,
                                float momentum, int training, float* running_mean, float* running_var,
                                long long* num_batches_tracked, float* scale, float* shift, float* save_mean,
                                float* save_rstd, fs2k_stream_t stream) {
    FS2K_REQUIRE(C > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(scale && shift, FS2K_ERR_NULL);
    FS2K_REQUIRE(training ? (sums != nullptr) : (running_mean && running_var), FS2K_ERR_NULL);
    fs2k_launch(bn_finalize_kernel, dim3(cdiv(C, 128)), dim3(128), 0, (cudaStream_t)stream, sums, M, C, gamma, beta, eps, momentum, training,
                                                                       running_mean, running_var, num_batches_tracked,
                                                                       scale, shift, save_mean, save_rstd);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_affine_act(const float* z, const float* scale, const float* shift, int act, const float* residual,
                               long M, int C, float dropout_p, long seed, float* y, fs2k_stream_t stream) {
    FS2K_REQUIRE(M >= 0 && C > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE((C & 3) == 0 && act >= 0 && act <= 3, FS2K_ERR_UNSUPPORTED);
    if (M == 0) return FS2K_OK;
    FS2K_REQUIRE(z && y && (!scale || shift), FS2K_ERR_NULL);  // scale == NULL: plain activation
    long g = (M * (C >> 2) + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    fs2k_launch(affine_act_kernel<false>, dim3((int)g), dim3(256), 0, (cudaStream_t)stream, z, scale, shift, act, residual, M, C, dropout_p,
                                                                (unsigned long long)seed, (void*)y);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_affine_act_bf16(const float* z, const float* scale, const float* shift, int act, const float* residual,
                                    long M, int C, float dropout_p, long seed, void* y_bf16, fs2k_stream_t stream) {
    FS2K_REQUIRE(M >= 0 && C > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE((C & 3) == 0 && act >= 0 && act <= 3, FS2K_ERR_UNSUPPORTED);
    if (M == 0) return FS2K_OK;
    FS2K_REQUIRE(z && y_bf16 && (!scale || shift), FS2K_ERR_NULL);
    long g = (M * (C >> 2) + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    fs2k_launch(affine_act_kernel<true>, dim3((int)g), dim3(256), 0, (cudaStream_t)stream, z, scale, shift, act, residual, M, C, dropout_p,
                                                                (unsigned long long)seed, y_bf16);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

FS2K_DEFINE_SEED_BASE_SETTER(norms)
