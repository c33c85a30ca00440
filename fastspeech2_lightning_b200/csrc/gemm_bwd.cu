// Backward pieces of the dense contractions (nn.Linear / Conv1d on channels-last data).
//
//   forward   y = (act(conv(x, W) + b)·alpha + residual)·row_mask
//   backward  gz = g·row_mask·alpha·act'(·)                       fs2k_act_bwd   (elementwise)
//             db = Σ_rows gz                                       fs2k_colsum
//             dx = conv_transpose(gz, W)                           forward GEMM on the flipped/transposed taps
//                                                                  (fs2k_repack_weight_t builds them)
//             dW[tap][n][k] = Σ_(b,l) gz[b,l,n]·x[b,l+tap−pad,k]   fs2k_gemm_wgrad  (reduction over rows)
#include <cuda_bf16.h>

#include "common.cuh"

namespace fs2k {

// mode: 0 none, 1 relu (aux = output y), 2 silu (aux = pre-activation z), 3 tanh (aux = output y)
__global__ void __launch_bounds__(256)
act_bwd_kernel(const float* __restrict__ g, const float* __restrict__ aux, int mode, float alpha,
               const uint8_t* __restrict__ row_mask, long M, int C, float drop_p, unsigned long long seed,
               float* __restrict__ gz) {
    pdl_prologue();
    seed = seed_with_base(seed);
    const uint32_t thr16 = drop_thr16(drop_p);
    const float inv_keep = drop_inv_keep(thr16);
    const int C4 = C >> 2;
    const long N = M * C4;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<const float4*>(g)[i];
        float s = alpha;
        if (row_mask) s *= row_mask[i / C4] ? 1.f : 0.f;
        float o[4] = {v.x * s, v.y * s, v.z * s, v.w * s};
        if (drop_p > 0.f) {  // the forward applied dropout after the activation: same mask on the gradient
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = drop_keep_g(seed, (unsigned long long)i, k, thr16) ? o[k] * inv_keep : 0.f;
        }
        if (mode != 0) {
            const float4 a = reinterpret_cast<const float4*>(aux)[i];
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (mode == 1) o[k] = av[k] > 0.f ? o[k] : 0.f;
                else if (mode == 2) {
                    const float sg = 1.0f / (1.0f + expf(-av[k]));
                    o[k] *= sg * (1.0f + av[k] * (1.0f - sg));
                } else o[k] *= 1.0f - av[k] * av[k];
            }
        }
        reinterpret_cast<float4*>(gz)[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// out[c] += Σ_m z[m,c]   (out zeroed by the launcher unless it accumulates; fp32 partials per CTA, one atomic per
// column per CTA; blockIdx.y selects the 128-channel block, blockIdx.x the row chunk)
template <bool IN16>
__global__ void __launch_bounds__(256)
colsum_kernel(const void* __restrict__ zv, long M, int C, long rows_per_cta, float* __restrict__ out) {
    const float* z = reinterpret_cast<const float*>(zv);
    const __nv_bfloat16* z16 = reinterpret_cast<const __nv_bfloat16*>(zv);
    pdl_prologue();
    __shared__ float s_part[8][128];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const long m0 = (long)blockIdx.x * rows_per_cta, m1 = min(M, m0 + rows_per_cta);
    {
        const int c0 = blockIdx.y * 128;
        const int c = c0 + lane * 4;
        float s[4] = {0, 0, 0, 0};
        if (c < C)
#pragma unroll 4
            for (long m = m0 + grp; m < m1; m += 8) {
                float4 v;
                if (IN16) {
                    const uint2 pk = *reinterpret_cast<const uint2*>(z16 + (size_t)m * C + c);
                    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.x));
                    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.y));
                    v = make_float4(a.x, a.y, b.x, b.y);
                } else {
                    v = *reinterpret_cast<const float4*>(z + (size_t)m * C + c);
                }
                s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
            }
#pragma unroll
        for (int k = 0; k < 4; ++k) s_part[grp][lane * 4 + k] = s[k];
        __syncthreads();
        if (threadIdx.x < 128 && c0 + threadIdx.x < C) {
            float a = 0.f;
#pragma unroll
            for (int g = 0; g < 8; ++g) a += s_part[g][threadIdx.x];
            atomicAdd(&out[c0 + threadIdx.x], a);
        }
        __syncthreads();
    }
}

// [taps][N][K] → [taps][K][N] with the taps reversed: the weights of the transposed convolution
__global__ void repack_weight_t_kernel(const float* __restrict__ w, int N, int K, int taps, float* __restrict__ out) {
    pdl_prologue();
    const long total = (long)N * K * taps;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int n = (int)(i % N);
        const int k = (int)((i / N) % K);
        const int t = (int)(i / ((long)N * K));
        out[i] = w[((size_t)(taps - 1 - t) * N + n) * K + k];
    }
}

// [taps][N][K] → PyTorch Conv1d layout [N][K][taps]
__global__ void unpack_conv_weight_kernel(const float* __restrict__ w, int N, int K, int taps, float* __restrict__ out) {
    pdl_prologue();
    const long total = (long)N * K * taps;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int t = (int)(i % taps);
        const int k = (int)((i / taps) % K);
        const int n = (int)(i / ((long)taps * K));
        out[i] = w[((size_t)t * N + n) * K + k];
    }
}

// dW[tap][n][k] += Σ_{m in this CTA's row range} G[m][n] · X[shift_tap(m)][k]      64×64 tile, 4×4 per thread
constexpr int WG_T = 64, WG_MK = 16;
__global__ void __launch_bounds__(256)
gemm_wgrad_kernel(const float* __restrict__ G, int ldg, const float* __restrict__ X, int ldx, int B, int L, int N,
                  int K, int taps, int pad, long rows_per_split, float* __restrict__ dW) {
    pdl_prologue();
    __shared__ __align__(16) float Gs[WG_MK][WG_T + 4];
    __shared__ __align__(16) float Xs[WG_MK][WG_T + 4];
    const int n0 = blockIdx.x * WG_T, k0 = blockIdx.y * WG_T;
    const int tap = blockIdx.z % taps, split = blockIdx.z / taps;
    const long M = (long)B * L;
    const long m_begin = (long)split * rows_per_split, m_end = min(M, m_begin + rows_per_split);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    // staging: 16 rows × 64 cols = 256 float4 → one float4 per thread per operand
    const int sr = tid >> 4, sc = (tid & 15) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (long mb = m_begin; mb < m_end; mb += WG_MK) {
        const long m = mb + sr;
        float4 gv = make_float4(0.f, 0.f, 0.f, 0.f), xv = gv;
        if (m < m_end) {
            if (n0 + sc < N) {
                const float* p = G + (size_t)m * ldg + n0 + sc;
                if (n0 + sc + 3 < N) gv = *reinterpret_cast<const float4*>(p);
                else { gv.x = p[0]; if (n0 + sc + 1 < N) gv.y = p[1]; if (n0 + sc + 2 < N) gv.z = p[2]; }
            }
            const int l = (int)(m % L);
            const int ls = l + tap - pad;
            if (ls >= 0 && ls < L && k0 + sc < K) {
                const float* p = X + (size_t)(m - l + ls) * ldx + k0 + sc;
                if (k0 + sc + 3 < K) xv = *reinterpret_cast<const float4*>(p);
                else { xv.x = p[0]; if (k0 + sc + 1 < K) xv.y = p[1]; if (k0 + sc + 2 < K) xv.z = p[2]; }
            }
        }
        __syncthreads();
        *reinterpret_cast<float4*>(&Gs[sr][sc]) = gv;
        *reinterpret_cast<float4*>(&Xs[sr][sc]) = xv;
        __syncthreads();
#pragma unroll
        for (int mm = 0; mm < WG_MK; ++mm) {
            const float4 a = *reinterpret_cast<const float4*>(&Gs[mm][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Xs[mm][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty * 4 + i;
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k < K) atomicAdd(&dW[((size_t)tap * N + n) * K + k], acc[i][j]);
        }
    }
}

}  // namespace fs2k

using namespace fs2k;

static inline int ew_grid2(long n) {
    long g = (n + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    return (int)(g < 1 ? 1 : g);
}

extern "C" int fs2k_act_bwd(const float* g, const float* aux, int mode, float alpha, const uint8_t* row_mask, long M,
                            int C, float dropout_p, long seed, float* gz, fs2k_stream_t stream) {
    FS2K_REQUIRE(M >= 0 && C > 0 && mode >= 0 && mode <= 3, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE((C & 3) == 0, FS2K_ERR_UNSUPPORTED);
    if (M == 0) return FS2K_OK;
    FS2K_REQUIRE(g && gz && (mode == 0 || aux), FS2K_ERR_NULL);
    fs2k_launch(act_bwd_kernel, dim3(ew_grid2(M * (C >> 2))), dim3(256), 0, (cudaStream_t)stream, g, aux, mode, alpha, row_mask, M, C, dropout_p, (unsigned long long)seed, gz);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_colsum(const float* z, long M, int C, float* out, int accumulate, fs2k_stream_t stream) {
    FS2K_REQUIRE(M >= 0 && C > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE((C & 3) == 0, FS2K_ERR_UNSUPPORTED);
    FS2K_REQUIRE(out && (z || M == 0), FS2K_ERR_NULL);
    cudaStream_t s = (cudaStream_t)stream;
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * C, s);
        if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    }
    if (M == 0) return FS2K_OK;
    long rows;
    const dim3 grid = col_reduce_grid(M, C, &rows);
    fs2k_launch(colsum_kernel<false>, dim3(grid), dim3(256), 0, s, (const void*)z, M, C, rows, out);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_colsum_bf16(const void* z_bf16, long M, int C, float* out, int accumulate, fs2k_stream_t stream) {
    FS2K_REQUIRE(M >= 0 && C > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE((C & 3) == 0, FS2K_ERR_UNSUPPORTED);
    FS2K_REQUIRE(out && (z_bf16 || M == 0), FS2K_ERR_NULL);
    cudaStream_t s = (cudaStream_t)stream;
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * C, s);
        if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    }
    if (M == 0) return FS2K_OK;
    long rows;
    const dim3 grid = col_reduce_grid(M, C, &rows);
    fs2k_launch(colsum_kernel<true>, dim3(grid), dim3(256), 0, s, z_bf16, M, C, rows, out);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_repack_weight_t(const float* w, int N, int K, int taps, float* out, fs2k_stream_t stream) {
    FS2K_REQUIRE(N > 0 && K > 0 && taps > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(w && out, FS2K_ERR_NULL);
    fs2k_launch(repack_weight_t_kernel, dim3(ew_grid2((long)N * K * taps)), dim3(256), 0, (cudaStream_t)stream, w, N, K, taps, out);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_unpack_conv_weight(const float* w, int N, int K, int taps, float* out, fs2k_stream_t stream) {
    FS2K_REQUIRE(N > 0 && K > 0 && taps > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(w && out, FS2K_ERR_NULL);
    fs2k_launch(unpack_conv_weight_kernel, dim3(ew_grid2((long)N * K * taps)), dim3(256), 0, (cudaStream_t)stream, w, N, K, taps, out);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_gemm_wgrad(const float* G, int ldg, const float* X, int ldx, int B, int L, int N, int K, int taps,
                               int pad, float* dW, fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && L >= 0 && N > 0 && K > 0 && taps >= 1 && pad >= 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE((ldg & 3) == 0 && (ldx & 3) == 0, FS2K_ERR_UNSUPPORTED);
    FS2K_REQUIRE(G && X && dW, FS2K_ERR_NULL);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)taps * N * K, s);
    if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    const long M = (long)B * L;
    if (M == 0) return FS2K_OK;
    const int gx = cdiv(N, WG_T), gy = cdiv(K, WG_T);
    // enough row splits to fill the chip (≈ 4 CTAs per SM), each a multiple of the 16-row staging step
    long splits = (148L * 4 + (long)gx * gy * taps - 1) / ((long)gx * gy * taps);
    long rows = (M + splits - 1) / splits;
    rows = ((rows + WG_MK - 1) / WG_MK) * WG_MK;
    if (rows < 64) rows = 64;
    splits = (M + rows - 1) / rows;
    FS2K_REQUIRE(splits * taps <= 65535, FS2K_ERR_UNSUPPORTED);
    dim3 grid(gx, gy, (unsigned)(splits * taps));
    fs2k_launch(gemm_wgrad_kernel, dim3(grid), dim3(256), 0, s, G, ldg, X, ldx, B, L, N, K, taps, pad, rows, dW);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

FS2K_DEFINE_SEED_BASE_SETTER(gemm_bwd)
