// Exact-fp32 (FFMA) GEMM / 1-D convolution with fused epilogue — the strict-fp32 path and the
// fallback for shapes the tensor-core kernel does not take (K or N not a multiple of 8).
//
//   C[(b,l), n] = epi( Σ_tap Σ_k A[b, l + tap − pad, k] · W[tap][n][k] )      (zero outside 0 ≤ l' < L)
//
// covers nn.Linear (taps = 1), 1×1 Conv1d and k-tap Conv1d on channels-last data:
//   Conformer FFN / in_proj / out_proj / pointwise convs   torchaudio conformer.py:42-75,103-108,151-153
//   variance-predictor pointwise conv, final Linear        fs2/blocks.py:7-16 ; fs2/variance_adaptor.py:53,58
//   aligner key/query projections                           fs2/attn/attention.py:122-151
//   mel_linear, PostNet k=5 convs                           fs2/model.py:121-123 ; fs2/layers.py:148-202
// Epilogue: + bias[n] → ·scale[n] + shift[n] (folded BatchNorm) → act → ·alpha + residual → ·row_mask.
//
// Tiling: 128×64 output tile per 256-thread CTA, BK = 16, 8×4 register tile per thread, operands
// staged k-major in shared memory (float4 reads), register double buffering of the next tile.
#include "common.cuh"

namespace fs2k {

constexpr int BM = 128, BN = 64, BK = 16;

struct GemmArgs {
    const float* A; int lda;
    int B, L, K;
    const float* W;  // [taps][N][K]
    int N, taps, pad;
    const float* bias; const float* scale; const float* shift;
    int act; float alpha;
    const float* residual; int ldr;
    const uint8_t* row_mask;
    float* C; int ldc;
};

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == FS2K_ACT_RELU) return fmaxf(v, 0.f);
    if (act == FS2K_ACT_SILU) return silu(v);
    if (act == FS2K_ACT_TANH) return tanhf(v);
    return v;
}

template <bool KVEC>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const GemmArgs g) {
    pdl_prologue();
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long M = (long)g.B * g.L;
    const long m0 = (long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int nk = (g.K + BK - 1) / BK;
    const int iters = g.taps * nk;

    // rows this thread stages for A (2 float4 per tile) and for W (1 float4 per tile)
    int a_l[2], a_row[2], a_kq[2];
    long a_base[2];
    bool a_ok[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int idx = tid + 256 * i;
        a_row[i] = idx >> 2;
        a_kq[i] = idx & 3;
        const long m = m0 + a_row[i];
        a_ok[i] = m < M;
        const long mm = a_ok[i] ? m : 0;
        a_l[i] = (int)(mm % g.L);
        a_base[i] = mm - a_l[i];  // b·L
    }
    const int w_row = tid >> 2, w_kq = tid & 3;
    const bool w_ok = (n0 + w_row) < g.N;

    float4 ra[2], rb;
    auto load_tile = [&](int it) {
        const int tap = it / nk, k0 = (it - tap * nk) * BK;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int lsrc = a_l[i] + tap - g.pad;
            const int k = k0 + a_kq[i] * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a_ok[i] && lsrc >= 0 && lsrc < g.L) {
                const float* p = g.A + (size_t)(a_base[i] + lsrc) * g.lda + k;
                if (KVEC) {
                    if (k < g.K) v = *reinterpret_cast<const float4*>(p);
                } else {
                    if (k + 0 < g.K) v.x = p[0];
                    if (k + 1 < g.K) v.y = p[1];
                    if (k + 2 < g.K) v.z = p[2];
                    if (k + 3 < g.K) v.w = p[3];
                }
            }
            ra[i] = v;
        }
        {
            const int k = k0 + w_kq * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (w_ok) {
                const float* p = g.W + ((size_t)tap * g.N + n0 + w_row) * g.K + k;
                if (KVEC) {
                    if (k < g.K) v = __ldg(reinterpret_cast<const float4*>(p));
                } else {
                    if (k + 0 < g.K) v.x = p[0];
                    if (k + 1 < g.K) v.y = p[1];
                    if (k + 2 < g.K) v.z = p[2];
                    if (k + 3 < g.K) v.w = p[3];
                }
            }
            rb = v;
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            As[buf][a_kq[i] * 4 + 0][a_row[i]] = ra[i].x;
            As[buf][a_kq[i] * 4 + 1][a_row[i]] = ra[i].y;
            As[buf][a_kq[i] * 4 + 2][a_row[i]] = ra[i].z;
            As[buf][a_kq[i] * 4 + 3][a_row[i]] = ra[i].w;
        }
        Bs[buf][w_kq * 4 + 0][w_row] = rb.x;
        Bs[buf][w_kq * 4 + 1][w_row] = rb.y;
        Bs[buf][w_kq * 4 + 2][w_row] = rb.z;
        Bs[buf][w_kq * 4 + 3][w_row] = rb.w;
    };

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    load_tile(0);
    store_tile(0);
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        const int buf = it & 1;
        if (it + 1 < iters) load_tile(it + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        if (it + 1 < iters) store_tile(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue ----
    const int n = n0 + tx * 4;
    if (n >= g.N) return;
    float bias[4] = {0, 0, 0, 0}, sc[4] = {1, 1, 1, 1}, sh[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (n + j < g.N) {
            if (g.bias) bias[j] = g.bias[n + j];
            if (g.scale) { sc[j] = g.scale[n + j]; sh[j] = g.shift[n + j]; }
        }
    }
    const bool vec = (n + 3 < g.N) && ((g.ldc & 3) == 0) && (!g.residual || (g.ldr & 3) == 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long m = m0 + ty * 8 + i;
        if (m >= M) break;
        const float rm = g.row_mask ? (g.row_mask[m] ? 1.f : 0.f) : 1.f;
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v = acc[i][j] + bias[j];
            if (g.scale) v = v * sc[j] + sh[j];
            v = apply_act(v, g.act);
            o[j] = v * g.alpha;
        }
        if (vec) {
            if (g.residual) {
                const float4 r = *reinterpret_cast<const float4*>(g.residual + (size_t)m * g.ldr + n);
                o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
            }
            if (g.row_mask) { o[0] *= rm; o[1] *= rm; o[2] *= rm; o[3] *= rm; }
            *reinterpret_cast<float4*>(g.C + (size_t)m * g.ldc + n) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (n + j < g.N) {
                    float v = o[j];
                    if (g.residual) v += g.residual[(size_t)m * g.ldr + n + j];
                    if (g.row_mask) v *= rm;
                    g.C[(size_t)m * g.ldc + n + j] = v;
                }
            }
        }
    }
}

// y[m] = (x[m,:]·w + b) · mask[m]   — nn.Linear(D → 1) + squeeze + mask (variance_adaptor.py:58-61)
__global__ void __launch_bounds__(256)
rowdot_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
              const uint8_t* __restrict__ mask, long M, int D, float* __restrict__ y) {
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    for (long m = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); m < M;
         m += (long)gridDim.x * (blockDim.x >> 5)) {
        float s = 0.f;
        for (int q = lane; q < (D >> 2); q += 32) {
            const float4 a = reinterpret_cast<const float4*>(x + (size_t)m * D)[q];
            const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + q);
            s += (a.x * ww.x + a.y * ww.y) + (a.z * ww.z + a.w * ww.w);
        }
        s = warp_sum(s);
        if (lane == 0) {
            float v = s + (b ? b[0] : 0.f);
            if (mask) v *= mask[m] ? 1.f : 0.f;
            y[m] = v;
        }
    }
}

// Conv1d weight [N][K][taps] (PyTorch layout) → [taps][N][K] (k contiguous per tap)
__global__ void repack_conv_weight_kernel(const float* __restrict__ w, int N, int K, int taps, float* __restrict__ out) {
    pdl_prologue();
    const long total = (long)N * K * taps;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int k = (int)(i % K);
        const int n = (int)((i / K) % N);
        const int t = (int)(i / ((long)K * N));
        out[i] = w[((size_t)n * K + k) * taps + t];
    }
}

}  // namespace fs2k

using namespace fs2k;

extern "C" int fs2k_gemm_f32(const float* A, int lda, int B, int L, int K, const float* W, int N, int taps, int pad,
                             const float* bias, const float* scale, const float* shift, int act, float alpha,
                             const float* residual, int ldr, const uint8_t* row_mask, float* C, int ldc,
                             fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && L >= 0 && K > 0 && N > 0 && taps >= 1 && pad >= 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(act >= 0 && act <= 3, FS2K_ERR_UNSUPPORTED);
    const long M = (long)B * L;
    if (M == 0) return FS2K_OK;
    FS2K_REQUIRE(A && W && C, FS2K_ERR_NULL);
    FS2K_REQUIRE(!scale || shift, FS2K_ERR_NULL);
    GemmArgs g{A, lda, B, L, K, W, N, taps, pad, bias, scale, shift, act, alpha, residual, ldr, row_mask, C, ldc};
    dim3 grid(cdiv(M, BM), cdiv(N, BN));
    const bool kvec = (K & 3) == 0 && (lda & 3) == 0 && ((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0;
    if (kvec) fs2k_launch(gemm_simt_kernel<true>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, g);
    else fs2k_launch(gemm_simt_kernel<false>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, g);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_rowdot(const float* x, const float* w, const float* b, const uint8_t* mask, long M, int D,
                           float* y, fs2k_stream_t stream) {
    FS2K_REQUIRE(M >= 0 && D > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE((D & 3) == 0, FS2K_ERR_UNSUPPORTED);
    if (M == 0) return FS2K_OK;
    FS2K_REQUIRE(x && w && y, FS2K_ERR_NULL);
    long g = (M + 7) / 8;
    if (g > 148 * 8) g = 148 * 8;
    fs2k_launch(rowdot_kernel, dim3((int)g), dim3(256), 0, (cudaStream_t)stream, x, w, b, mask, M, D, y);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_repack_conv_weight(const float* w, int N, int K, int taps, float* out, fs2k_stream_t stream) {
    FS2K_REQUIRE(N > 0 && K > 0 && taps > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(w && out, FS2K_ERR_NULL);
    long total = (long)N * K * taps;
    long g = (total + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    fs2k_launch(repack_conv_weight_kernel, dim3((int)g), dim3(256), 0, (cudaStream_t)stream, w, N, K, taps, out);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}
