// Backward of LayerNorm and of BatchNorm1d (+ fused activation).
#include <cuda_bf16.h>

#include "common.cuh"

namespace fs2k {

// ---- LayerNorm backward: one warp per row; dγ/dβ accumulated per lane, reduced per CTA, then atomics ----
template <int MAXV>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ gamma, long M, int D, float drop_p,
                     unsigned long long seed, const float* __restrict__ g_add, float* __restrict__ dx,
                     float* __restrict__ dgamma, float* __restrict__ dbeta) {
    pdl_prologue();
    seed = seed_with_base(seed);
    const uint32_t thr16 = drop_thr16(drop_p);
    const float inv_keep = drop_inv_keep(thr16);
    __shared__ float s_red[8][32 * MAXV * 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D4 = D >> 2;
    float4 gam[MAXV], dg[MAXV], db[MAXV];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int q = lane + 32 * i;
        gam[i] = q < D4 ? __ldg(reinterpret_cast<const float4*>(gamma) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        db[i] = dg[i];
    }
    const float inv_d = 1.0f / (float)D;
    // the row loop is a chain of exposed load latencies (load → two warp reductions → store): the next row's operands are
    // requested before the current row is reduced, so a warp always has two rows of loads in flight
    const long stride = (long)gridDim.x * 8;
    float4 gn[MAXV], xn[MAXV];
    float mun = 0.f, rsn = 0.f;
    auto fetch = [&](long m) {
        if (m < M) {
            mun = mean[m];
            rsn = rstd[m];
#pragma unroll
            for (int i = 0; i < MAXV; ++i) {
                const int q = lane + 32 * i;
                if (q < D4) {
                    gn[i] = reinterpret_cast<const float4*>(g + (size_t)m * D)[q];
                    xn[i] = reinterpret_cast<const float4*>(x + (size_t)m * D)[q];
                }
            }
        }
    };
    constexpr bool PIPELINED = MAXV <= 4;  // D = 1024 would spill with two rows of operands in registers
    if (PIPELINED) fetch((long)blockIdx.x * 8 + warp);
    for (long m = (long)blockIdx.x * 8 + warp; m < M; m += stride) {
        if (!PIPELINED) fetch(m);
        const float mu = mun, rs = rsn;
        float4 gc[MAXV], xc[MAXV];
#pragma unroll
        for (int i = 0; i < MAXV; ++i) { gc[i] = gn[i]; xc[i] = xn[i]; }
        if (PIPELINED) fetch(m + stride);
        float4 gg[MAXV], xh[MAXV];
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int q = lane + 32 * i;
            if (q < D4) {
                float4 gv = gc[i];
                if (drop_p > 0.f) {  // the forward output was dropout(LN(x)): same mask on the incoming gradient
                    const unsigned long long e = (unsigned long long)m * D + q * 4;
                    drop_apply4(gv, seed, e, thr16, inv_keep);
                }
                const float4 xv = xc[i];
                xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
                gg[i] = make_float4(gv.x * gam[i].x, gv.y * gam[i].y, gv.z * gam[i].z, gv.w * gam[i].w);
                a += (gg[i].x + gg[i].y) + (gg[i].z + gg[i].w);
                b += (gg[i].x * xh[i].x + gg[i].y * xh[i].y) + (gg[i].z * xh[i].z + gg[i].w * xh[i].w);
                dg[i].x += gv.x * xh[i].x; dg[i].y += gv.y * xh[i].y; dg[i].z += gv.z * xh[i].z; dg[i].w += gv.w * xh[i].w;
                db[i].x += gv.x; db[i].y += gv.y; db[i].z += gv.z; db[i].w += gv.w;
            }
        }
        a = warp_sum(a) * inv_d;
        b = warp_sum(b) * inv_d;
        float4 radd[MAXV];
        if (g_add) {  // the residual branch's gradient joins here (x feeds both LN and the skip connection)
#pragma unroll
            for (int i = 0; i < MAXV; ++i) {
                const int q = lane + 32 * i;
                if (q < D4) radd[i] = reinterpret_cast<const float4*>(g_add + (size_t)m * D)[q];
            }
        }
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int q = lane + 32 * i;
            if (q < D4) {
                float4 o;
                o.x = rs * (gg[i].x - a - xh[i].x * b);
                o.y = rs * (gg[i].y - a - xh[i].y * b);
                o.z = rs * (gg[i].z - a - xh[i].z * b);
                o.w = rs * (gg[i].w - a - xh[i].w * b);
                if (g_add) { o.x += radd[i].x; o.y += radd[i].y; o.z += radd[i].z; o.w += radd[i].w; }
                reinterpret_cast<float4*>(dx + (size_t)m * D)[q] = o;
            }
        }
    }
    // CTA reduction of dγ then dβ
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const float4 v = pass == 0 ? dg[i] : db[i];
            float* p = &s_red[warp][(lane + 32 * i) * 4];
            p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
        }
        __syncthreads();
        for (int c = threadIdx.x; c < D; c += blockDim.x) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += s_red[w][c];
            atomicAdd(pass == 0 ? &dgamma[c] : &dbeta[c], s);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ float act_grad_from_u(float u, int act) {
    if (act == FS2K_ACT_SILU) {
        const float sg = 1.0f / (1.0f + expf(-u));
        return sg * (1.0f + u * (1.0f - sg));
    }
    if (act == FS2K_ACT_TANH) {
        const float t = tanhf(u);
        return 1.0f - t * t;
    }
    if (act == FS2K_ACT_RELU) return u > 0.f ? 1.f : 0.f;
    return 1.f;
}

// sums[c] += Σ gu ; sums[C+c] += Σ gu·ẑ   with gu = g·act'(z·scale+shift), ẑ = (z−mean)·rstd   (fp64)
__global__ void __launch_bounds__(256)
bn_bwd_stats_kernel(const float* __restrict__ g, const float* __restrict__ z, const float* __restrict__ scale,
                    const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ rstd,
                    int act, long M, int C, long rows_per_cta, float drop_p, unsigned long long seed,
                    double* __restrict__ sums) {
    pdl_prologue();
    seed = seed_with_base(seed);
    const uint32_t thr16 = drop_thr16(drop_p);
    const float inv_keep = drop_inv_keep(thr16);
    __shared__ double s_part[8][128][2];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const long m0 = (long)blockIdx.x * rows_per_cta, m1 = min(M, m0 + rows_per_cta);
    {
        const int c0 = blockIdx.y * 128;  // blockIdx.y = 128-channel block, blockIdx.x = row chunk
        const int c = c0 + lane * 4;
        double ds[4] = {0, 0, 0, 0}, dq[4] = {0, 0, 0, 0};
        if (c < C) {
            float sc[4], sh[4], mu[4], rs[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { sc[k] = scale[c + k]; sh[k] = shift[c + k]; mu[k] = mean[c + k]; rs[k] = rstd[c + k]; }
            float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
            int n = 0;
#pragma unroll 2
            for (long m = m0 + grp; m < m1; m += 8) {
                const float4 gv4 = *reinterpret_cast<const float4*>(g + (size_t)m * C + c);
                const float4 zv4 = *reinterpret_cast<const float4*>(z + (size_t)m * C + c);
                const float gv[4] = {gv4.x, gv4.y, gv4.z, gv4.w}, zv[4] = {zv4.x, zv4.y, zv4.z, zv4.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float gk = gv[k];
                    if (drop_p > 0.f) gk = drop_keep(seed, (unsigned long long)m * C + c + k, thr16) ? gk * inv_keep : 0.f;
                    const float gu = gk * act_grad_from_u(zv[k] * sc[k] + sh[k], act);
                    s[k] += gu;
                    q[k] += gu * (zv[k] - mu[k]) * rs[k];
                }
                if (++n == 32) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) { ds[k] += s[k]; dq[k] += q[k]; s[k] = 0; q[k] = 0; }
                    n = 0;
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) { ds[k] += s[k]; dq[k] += q[k]; }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) { s_part[grp][lane * 4 + k][0] = ds[k]; s_part[grp][lane * 4 + k][1] = dq[k]; }
        __syncthreads();
        if (threadIdx.x < 128 && c0 + threadIdx.x < C) {
            double a = 0, b = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) { a += s_part[w][threadIdx.x][0]; b += s_part[w][threadIdx.x][1]; }
            atomicAdd(&sums[c0 + threadIdx.x], a);
            atomicAdd(&sums[C + c0 + threadIdx.x], b);
        }
        __syncthreads();
    }
}

// training: gz = γ·rstd·(gu − Σgu/M − ẑ·Σ(gu·ẑ)/M) ; eval: gz = gu·scale.   dγ = Σ gu·ẑ, dβ = Σ gu (written by block 0)
template <bool OUT16>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ g, const float* __restrict__ z, const float* __restrict__ scale,
                    const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const double* __restrict__ sums, int act, int training, long M, int C, float drop_p,
                    unsigned long long seed, int accumulate, void* __restrict__ gz, float* __restrict__ dgamma,
                    float* __restrict__ dbeta) {
    pdl_prologue();
    seed = seed_with_base(seed);
    const uint32_t thr16 = drop_thr16(drop_p);
    const float inv_keep = drop_inv_keep(thr16);
    const int C4 = C >> 2;
    const long N = M * C4;
    const float inv_m = 1.0f / (float)M;
    if (blockIdx.x == 0)
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)sums[c];
            if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)sums[C + c];
        }
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        const float4 gv4 = reinterpret_cast<const float4*>(g)[i];
        const float4 zv4 = reinterpret_cast<const float4*>(z)[i];
        const float gv[4] = {gv4.x, gv4.y, gv4.z, gv4.w}, zv[4] = {zv4.x, zv4.y, zv4.z, zv4.w};
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float sc = scale[c + k];
            float gk = gv[k];
            if (drop_p > 0.f) gk = drop_keep_g(seed, (unsigned long long)i, k, thr16) ? gk * inv_keep : 0.f;
            const float gu = gk * act_grad_from_u(zv[k] * sc + shift[c + k], act);
            if (training) {
                const float zh = (zv[k] - mean[c + k]) * rstd[c + k];
                o[k] = sc * (gu - (float)sums[c + k] * inv_m - zh * (float)sums[C + c + k] * inv_m);
            } else {
                o[k] = gu * sc;
            }
        }
        if (OUT16) {  // bf16 mode: gz only feeds the data- and weight-gradient contractions
            const __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]), b = __floats2bfloat162_rn(o[2], o[3]);
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t*>(&a);
            pk.y = *reinterpret_cast<const uint32_t*>(&b);
            reinterpret_cast<uint2*>(gz)[i] = pk;
        } else {
            reinterpret_cast<float4*>(gz)[i] = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

}  // namespace fs2k

using namespace fs2k;

static int layernorm_bwd_impl(const float* g, const float* x, const float* mean, const float* rstd,
                              const float* gamma, long M, int D, float dropout_p, long seed, const float* g_add, float* dx,
                              float* dgamma, float* dbeta, int accumulate, fs2k_stream_t stream) {
    const unsigned long long useed = (unsigned long long)seed;
    FS2K_REQUIRE(M >= 0 && D > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE((D & 3) == 0 && D <= 1024, FS2K_ERR_UNSUPPORTED);
    FS2K_REQUIRE(dgamma && dbeta && (M == 0 || (g && x && mean && rstd && gamma && dx)), FS2K_ERR_NULL);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
    if (!accumulate) {  // otherwise the atomics add on top of what dgamma / dbeta already hold
        e = cudaMemsetAsync(dgamma, 0, sizeof(float) * D, s);
        if (e == cudaSuccess) e = cudaMemsetAsync(dbeta, 0, sizeof(float) * D, s);
    }
    if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    if (M == 0) return FS2K_OK;
    // 4 rows per warp: the row loop is a chain of exposed load latencies (8 rows per warp took 19 µs on 16 MB), while
    // the per-CTA dγ/dβ atomics (2·D per CTA) stay a small fraction
    long grid = (M + 31) / 32;
    if (grid > 148 * 8) grid = 148 * 8;
    if (grid < 1) grid = 1;
    if (D <= 256) fs2k_launch(layernorm_bwd_kernel<2>, dim3((int)grid), dim3(256), 0, s, g, x, mean, rstd, gamma, M, D, dropout_p, useed, g_add, dx, dgamma, dbeta);
    else if (D <= 512) fs2k_launch(layernorm_bwd_kernel<4>, dim3((int)grid), dim3(256), 0, s, g, x, mean, rstd, gamma, M, D, dropout_p, useed, g_add, dx, dgamma, dbeta);
    else fs2k_launch(layernorm_bwd_kernel<8>, dim3((int)grid), dim3(256), 0, s, g, x, mean, rstd, gamma, M, D, dropout_p, useed, g_add, dx, dgamma, dbeta);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_layernorm_bwd(const float* g, const float* x, const float* mean, const float* rstd,
                                  const float* gamma, long M, int D, float dropout_p, long seed, float* dx,
                                  float* dgamma, float* dbeta, int accumulate, fs2k_stream_t stream) {
    return layernorm_bwd_impl(g, x, mean, rstd, gamma, M, D, dropout_p, seed, nullptr, dx, dgamma, dbeta, accumulate, stream);
}

// dx = LayerNorm backward + g_add: the gradient of the residual branch (x feeds the LayerNorm AND the skip connection) joins
// in the same launch instead of in a separate elementwise add
extern "C" int fs2k_layernorm_bwd_add(const float* g, const float* x, const float* mean, const float* rstd,
                                      const float* gamma, long M, int D, const float* g_add, float* dx, float* dgamma,
                                      float* dbeta, int accumulate, fs2k_stream_t stream) {
    FS2K_REQUIRE(g_add != nullptr || M == 0, FS2K_ERR_NULL);
    return layernorm_bwd_impl(g, x, mean, rstd, gamma, M, D, 0.f, 0, g_add, dx, dgamma, dbeta, accumulate, stream);
}

static int bn_act_bwd_impl(const float* g, const float* z, const float* scale, const float* shift,
                           const float* mean, const float* rstd, int act, int training, long M, int C,
                           float dropout_p, long seed, double* sums, void* gz, int gz_bf16, float* dgamma,
                           float* dbeta, int accumulate, fs2k_stream_t stream) {
    const unsigned long long useed = (unsigned long long)seed;
    FS2K_REQUIRE(M >= 0 && C > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE((C & 3) == 0 && act >= 0 && act <= 3, FS2K_ERR_UNSUPPORTED);
    FS2K_REQUIRE(sums && (M == 0 || (g && z && scale && shift && mean && rstd && gz)), FS2K_ERR_NULL);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, s);
    if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    if (M == 0) return FS2K_OK;
    long rows;
    const dim3 sgrid = col_reduce_grid(M, C, &rows);
    fs2k_launch(bn_bwd_stats_kernel, dim3(sgrid), dim3(256), 0, s, g, z, scale, shift, mean, rstd, act, M, C, rows, dropout_p, useed, sums);
    FS2K_CHECK_LAUNCH();
    long grid = (M * (C >> 2) + 255) / 256;
    if (grid > 148 * 16) grid = 148 * 16;
    if (gz_bf16)
        fs2k_launch(bn_bwd_apply_kernel<true>, dim3((int)grid), dim3(256), 0, s, g, z, scale, shift, mean, rstd, sums, act, training, M, C, dropout_p, useed, accumulate, gz, dgamma, dbeta);
    else
        fs2k_launch(bn_bwd_apply_kernel<false>, dim3((int)grid), dim3(256), 0, s, g, z, scale, shift, mean, rstd, sums, act, training, M, C, dropout_p, useed, accumulate, gz, dgamma, dbeta);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_bn_act_bwd(const float* g, const float* z, const float* scale, const float* shift,
                               const float* mean, const float* rstd, int act, int training, long M, int C,
                               float dropout_p, long seed, double* sums /* [2C] scratch */, float* gz, float* dgamma,
                               float* dbeta, int accumulate, fs2k_stream_t stream) {
    return bn_act_bwd_impl(g, z, scale, shift, mean, rstd, act, training, M, C, dropout_p, seed, sums, gz, 0, dgamma, dbeta, accumulate, stream);
}

extern "C" int fs2k_bn_act_bwd_bf16(const float* g, const float* z, const float* scale, const float* shift,
                                    const float* mean, const float* rstd, int act, int training, long M, int C,
                                    float dropout_p, long seed, double* sums, void* gz_bf16, float* dgamma,
                                    float* dbeta, int accumulate, fs2k_stream_t stream) {
    return bn_act_bwd_impl(g, z, scale, shift, mean, rstd, act, training, M, C, dropout_p, seed, sums, gz_bf16, 1, dgamma, dbeta, accumulate, stream);
}

FS2K_DEFINE_SEED_BASE_SETTER(norms_bwd)
