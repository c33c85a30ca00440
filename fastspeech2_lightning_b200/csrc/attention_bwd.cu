// Backward of the masked multi-head self-attention (flash style: S and P are recomputed from the saved
// log-sum-exp, nothing of size L×L is stored).  Forward: attention_simt.cu.
//
//   delta[q] = Σ_d dO[q,d]·O[q,d]
//   P = exp(Q Kᵀ·scale − lse[q])  (0 on padded keys) ; dP = dO Vᵀ ; dS = P ⊙ (dP − delta)·scale
//   dQ = dS K ; dK = dSᵀ Q ; dV = Pᵀ dO
// Two kernels so that every output tile has a single owner (no atomics): `dq` walks the key tiles for a
// block of 64 queries; `dkv` walks the query tiles for a block of 64 keys.  Exact fp32 FFMA arithmetic.
#include "common.cuh"

namespace fs2k {

// delta[b,h,q] — one warp per (b,q,h)
__global__ void __launch_bounds__(256)
attn_delta_kernel(const float* __restrict__ o, const float* __restrict__ dO, int B, int L, int H, int HD,
                  float* __restrict__ delta) {
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    const long n = (long)B * L * H;
    for (long i = (long)blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += (long)gridDim.x * 8) {
        const int h = (int)(i % H);
        const long bq = i / H;
        const int b = (int)(bq / L), q = (int)(bq % L);
        const float* po = o + (size_t)bq * H * HD + h * HD;
        const float* pd = dO + (size_t)bq * H * HD + h * HD;
        float s = 0.f;
        for (int d = lane * 4; d < HD; d += 128) {
            const float4 a = *reinterpret_cast<const float4*>(po + d), c = *reinterpret_cast<const float4*>(pd + d);
            s += (a.x * c.x + a.y * c.y) + (a.z * c.z + a.w * c.w);
        }
        s = warp_sum(s);
        if (lane == 0) delta[((size_t)b * H + h) * L + q] = s;
    }
}

template <int HD>
__global__ void __launch_bounds__(256)
attn_bwd_dq_kernel(const float* __restrict__ qkv, const float* __restrict__ dO, const float* __restrict__ lse,
                   const float* __restrict__ delta, const int* __restrict__ lens, int L, int H, float scale,
                   float drop_p, unsigned long long seed, const int* __restrict__ order, float* __restrict__ dqkv) {
    pdl_prologue();
    seed = seed_with_base(seed);
    constexpr int BQ = 64, BKEY = 64, QS = HD + 4, PS = BKEY + 4, NJ = HD / 64;
    extern __shared__ __align__(16) float smem[];
    float* Qs = smem;               // [BQ][QS]
    float* Gs = Qs + BQ * QS;       // dO [BQ][QS]
    float* Ks = Gs + BQ * QS;       // [BKEY][QS]
    float* Vs = Ks + BKEY * QS;     // [BKEY][QS]
    float* Ss = Vs + BKEY * QS;     // dS [BQ][PS]
    const int b = order ? order[blockIdx.z] : blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;  // longest utterance first
    const int D = H * HD, ld = 3 * D;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int len = min(lens[b], L);
    const float inv_keep = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.f;
    const float* base = qkv + (size_t)b * L * ld + h * HD;
    for (int i = tid; i < BQ * (HD / 4); i += 256) {
        const int r = i / (HD / 4), c = i % (HD / 4);
        float4 qv = make_float4(0.f, 0.f, 0.f, 0.f), gv = qv;
        if (q0 + r < L) {
            qv = *reinterpret_cast<const float4*>(base + (size_t)(q0 + r) * ld + c * 4);
            gv = *reinterpret_cast<const float4*>(dO + ((size_t)b * L + q0 + r) * D + h * HD + c * 4);
        }
        *reinterpret_cast<float4*>(Qs + r * QS + c * 4) = qv;
        *reinterpret_cast<float4*>(Gs + r * QS + c * 4) = gv;
    }
    float lse_r[4], del_r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = q0 + ty * 4 + i;
        lse_r[i] = q < L ? lse[((size_t)b * H + h) * L + q] : 0.f;
        del_r[i] = q < L ? delta[((size_t)b * H + h) * L + q] : 0.f;
    }
    float4 acc[4][NJ];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int n_tiles = (len + BKEY - 1) / BKEY;
    for (int kt = 0; kt < n_tiles; ++kt) {
        const int k0 = kt * BKEY;
        __syncthreads();
        for (int i = tid; i < BKEY * (HD / 4); i += 256) {
            const int r = i / (HD / 4), c = i % (HD / 4);
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (k0 + r < L) {
                const float* p = base + (size_t)(k0 + r) * ld + c * 4;
                kv = *reinterpret_cast<const float4*>(p + D);
                vv = *reinterpret_cast<const float4*>(p + 2 * D);
            }
            *reinterpret_cast<float4*>(Ks + r * QS + c * 4) = kv;
            *reinterpret_cast<float4*>(Vs + r * QS + c * 4) = vv;
        }
        __syncthreads();
        // 4 queries × 4 keys per thread (24 shared-memory wavefronts per 128 FFMA; 4×2 over 32-key tiles was LDS-bound)
        float s[4][4], dp[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = dp[i][j] = 0.f;
#pragma unroll 2
        for (int c = 0; c < HD / 4; ++c) {
            float4 q4[4], g4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                q4[i] = *reinterpret_cast<const float4*>(Qs + (ty * 4 + i) * QS + c * 4);
                g4[i] = *reinterpret_cast<const float4*>(Gs + (ty * 4 + i) * QS + c * 4);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 k4 = *reinterpret_cast<const float4*>(Ks + (tx + 16 * j) * QS + c * 4);
                const float4 v4 = *reinterpret_cast<const float4*>(Vs + (tx + 16 * j) * QS + c * 4);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    s[i][j] = fmaf(q4[i].x, k4.x, s[i][j]); s[i][j] = fmaf(q4[i].y, k4.y, s[i][j]);
                    s[i][j] = fmaf(q4[i].z, k4.z, s[i][j]); s[i][j] = fmaf(q4[i].w, k4.w, s[i][j]);
                    dp[i][j] = fmaf(g4[i].x, v4.x, dp[i][j]); dp[i][j] = fmaf(g4[i].y, v4.y, dp[i][j]);
                    dp[i][j] = fmaf(g4[i].z, v4.z, dp[i][j]); dp[i][j] = fmaf(g4[i].w, v4.w, dp[i][j]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int key = k0 + tx + 16 * j;
                const float p = key < len ? expf(s[i][j] * scale - lse_r[i]) : 0.f;
                const float mk = attn_keep(seed, drop_p, inv_keep, b, h, q0 + ty * 4 + i, key, H, L);
                Ss[(ty * 4 + i) * PS + tx + 16 * j] = p * (dp[i][j] * mk - del_r[i]) * scale;
            }
        __syncthreads();
        // dQ += dS K
#pragma unroll 2
        for (int kk = 0; kk < BKEY; kk += 4) {
            float4 p4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) p4[i] = *reinterpret_cast<const float4*>(Ss + (ty * 4 + i) * PS + kk);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const float4 kv = *reinterpret_cast<const float4*>(Ks + (kk + u) * QS + tx * 4 + 64 * j);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float p = u == 0 ? p4[i].x : (u == 1 ? p4[i].y : (u == 2 ? p4[i].z : p4[i].w));
                        acc[i][j].x = fmaf(p, kv.x, acc[i][j].x); acc[i][j].y = fmaf(p, kv.y, acc[i][j].y);
                        acc[i][j].z = fmaf(p, kv.z, acc[i][j].z); acc[i][j].w = fmaf(p, kv.w, acc[i][j].w);
                    }
                }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = q0 + ty * 4 + i;
        if (q >= L) continue;
#pragma unroll
        for (int j = 0; j < NJ; ++j)
            *reinterpret_cast<float4*>(dqkv + ((size_t)b * L + q) * ld + h * HD + tx * 4 + 64 * j) = acc[i][j];
    }
}

template <int HD>
__global__ void __launch_bounds__(256)
attn_bwd_dkv_kernel(const float* __restrict__ qkv, const float* __restrict__ dO, const float* __restrict__ lse,
                    const float* __restrict__ delta, const int* __restrict__ lens, int L, int H, float scale,
                    float drop_p, unsigned long long seed, const int* __restrict__ order, float* __restrict__ dqkv) {
    pdl_prologue();
    seed = seed_with_base(seed);
    constexpr int BKEY = 64, BQ = 64, QS = HD + 4, PS = BKEY + 4, NJ = HD / 64;
    extern __shared__ __align__(16) float smem[];
    float* Ks = smem;               // [BKEY][QS]
    float* Vs = Ks + BKEY * QS;     // [BKEY][QS]
    float* Qs = Vs + BKEY * QS;     // [BQ][QS]
    float* Gs = Qs + BQ * QS;       // dO [BQ][QS]
    float* Ps = Gs + BQ * QS;       // P  [BQ][PS]
    float* Ss = Ps + BQ * PS;       // dS [BQ][PS]
    __shared__ float s_lse[BQ], s_del[BQ];
    const int b = order ? order[blockIdx.z] : blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * BKEY;  // longest utterance first
    const int D = H * HD, ld = 3 * D;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int len = min(lens[b], L);
    const float inv_keep = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.f;
    const float* base = qkv + (size_t)b * L * ld + h * HD;
    float4 dk[4][NJ], dv[4][NJ];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) dk[i][j] = dv[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k0 < len) {
        for (int i = tid; i < BKEY * (HD / 4); i += 256) {
            const int r = i / (HD / 4), c = i % (HD / 4);
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (k0 + r < L) {
                const float* p = base + (size_t)(k0 + r) * ld + c * 4;
                kv = *reinterpret_cast<const float4*>(p + D);
                vv = *reinterpret_cast<const float4*>(p + 2 * D);
            }
            *reinterpret_cast<float4*>(Ks + r * QS + c * 4) = kv;
            *reinterpret_cast<float4*>(Vs + r * QS + c * 4) = vv;
        }
        for (int q0 = 0; q0 < L; q0 += BQ) {
            __syncthreads();
            for (int i = tid; i < BQ * (HD / 4); i += 256) {
                const int r = i / (HD / 4), c = i % (HD / 4);
                float4 qv = make_float4(0.f, 0.f, 0.f, 0.f), gv = qv;
                if (q0 + r < L) {
                    qv = *reinterpret_cast<const float4*>(base + (size_t)(q0 + r) * ld + c * 4);
                    gv = *reinterpret_cast<const float4*>(dO + ((size_t)b * L + q0 + r) * D + h * HD + c * 4);
                }
                *reinterpret_cast<float4*>(Qs + r * QS + c * 4) = qv;
                *reinterpret_cast<float4*>(Gs + r * QS + c * 4) = gv;
            }
            if (tid < BQ) {
                const int q = q0 + tid;
                s_lse[tid] = q < L ? lse[((size_t)b * H + h) * L + q] : INFINITY;  // exp(−inf) = 0 for rows past L
                s_del[tid] = q < L ? delta[((size_t)b * H + h) * L + q] : 0.f;
            }
            __syncthreads();
            // S, dP for 4 queries (ty*4..+3) × 4 keys (tx + 16j): 24 shared-memory wavefronts per 128 FFMA (the 2×4
            // tile this replaces needed 20 per 64 and was bound by the LDS pipe)
            float s[4][4], dp[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) s[i][j] = dp[i][j] = 0.f;
#pragma unroll 2
            for (int c = 0; c < HD / 4; ++c) {
                float4 q4[4], g4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    q4[i] = *reinterpret_cast<const float4*>(Qs + (ty * 4 + i) * QS + c * 4);
                    g4[i] = *reinterpret_cast<const float4*>(Gs + (ty * 4 + i) * QS + c * 4);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 k4 = *reinterpret_cast<const float4*>(Ks + (tx + 16 * j) * QS + c * 4);
                    const float4 v4 = *reinterpret_cast<const float4*>(Vs + (tx + 16 * j) * QS + c * 4);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        s[i][j] = fmaf(q4[i].x, k4.x, s[i][j]); s[i][j] = fmaf(q4[i].y, k4.y, s[i][j]);
                        s[i][j] = fmaf(q4[i].z, k4.z, s[i][j]); s[i][j] = fmaf(q4[i].w, k4.w, s[i][j]);
                        dp[i][j] = fmaf(g4[i].x, v4.x, dp[i][j]); dp[i][j] = fmaf(g4[i].y, v4.y, dp[i][j]);
                        dp[i][j] = fmaf(g4[i].z, v4.z, dp[i][j]); dp[i][j] = fmaf(g4[i].w, v4.w, dp[i][j]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int key = k0 + tx + 16 * j, r = ty * 4 + i;
                    const float p = key < len ? expf(s[i][j] * scale - s_lse[r]) : 0.f;
                    const float mk = attn_keep(seed, drop_p, inv_keep, b, h, q0 + r, key, H, L);
                    Ps[r * PS + tx + 16 * j] = p * mk;
                    Ss[r * PS + tx + 16 * j] = p * (dp[i][j] * mk - s_del[r]) * scale;
                }
            __syncthreads();
            // dV[key][d] += Σ_q P[q][key]·dO[q][d] ; dK[key][d] += Σ_q dS[q][key]·Q[q][d]   keys ty*4..+3, d = tx*4 + 64j
#pragma unroll 2
            for (int q = 0; q < BQ; ++q) {
                const float4 p4 = *reinterpret_cast<const float4*>(Ps + q * PS + ty * 4);
                const float4 s4 = *reinterpret_cast<const float4*>(Ss + q * PS + ty * 4);
                const float pv[4] = {p4.x, p4.y, p4.z, p4.w}, sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const float4 gv = *reinterpret_cast<const float4*>(Gs + q * QS + tx * 4 + 64 * j);
                    const float4 qv = *reinterpret_cast<const float4*>(Qs + q * QS + tx * 4 + 64 * j);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        dv[i][j].x = fmaf(pv[i], gv.x, dv[i][j].x); dv[i][j].y = fmaf(pv[i], gv.y, dv[i][j].y);
                        dv[i][j].z = fmaf(pv[i], gv.z, dv[i][j].z); dv[i][j].w = fmaf(pv[i], gv.w, dv[i][j].w);
                        dk[i][j].x = fmaf(sv[i], qv.x, dk[i][j].x); dk[i][j].y = fmaf(sv[i], qv.y, dk[i][j].y);
                        dk[i][j].z = fmaf(sv[i], qv.z, dk[i][j].z); dk[i][j].w = fmaf(sv[i], qv.w, dk[i][j].w);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int key = k0 + ty * 4 + i;
        if (key >= L) continue;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            float* p = dqkv + ((size_t)b * L + key) * ld + h * HD + tx * 4 + 64 * j;
            *reinterpret_cast<float4*>(p + D) = dk[i][j];
            *reinterpret_cast<float4*>(p + 2 * D) = dv[i][j];
        }
    }
}

}  // namespace fs2k

using namespace fs2k;

extern "C" int fs2k_attention_bwd_f32(const float* qkv, const float* out, const float* lse, const float* dout,
                                      const int* lens, int B, int L, int H, int head_dim, float dropout_p, long seed,
                                      float* delta /* [B,H,L] */, float* dqkv, const int* order, fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && L >= 0 && H > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(head_dim == 64 || head_dim == 128, FS2K_ERR_UNSUPPORTED);
    if (B == 0 || L == 0) return FS2K_OK;
    FS2K_REQUIRE(qkv && out && lse && dout && lens && delta && dqkv, FS2K_ERR_NULL);
    FS2K_REQUIRE(B <= 65535, FS2K_ERR_UNSUPPORTED);
    cudaStream_t s = (cudaStream_t)stream;
    const float scale = 1.0f / sqrtf((float)head_dim);
    long g = ((long)B * L * H + 7) / 8;
    if (g > 148 * 8) g = 148 * 8;
    fs2k_launch(attn_delta_kernel, dim3((int)g), dim3(256), 0, s, out, dout, B, L, H, head_dim, delta);
    FS2K_CHECK_LAUNCH();
    const int QS = head_dim + 4;
    const int smem_dq = ((64 + 64 + 64 + 64) * QS + 64 * 68) * 4;
    const int smem_dkv = ((64 + 64 + 64 + 64) * QS + 2 * 64 * 68) * 4;
    cudaError_t e;
    if (head_dim == 128) {
        e = cudaFuncSetAttribute(attn_bwd_dq_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dq);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_dkv_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dkv);
        if (e != cudaSuccess) return fs2k_set_cuda_error(e);
        fs2k_launch(attn_bwd_dq_kernel<128>, dim3(dim3(cdiv(L, 64), H, B)), dim3(256), smem_dq, s, qkv, dout, lse, delta, lens, L, H, scale, dropout_p, (unsigned long long)seed, order, dqkv);
        FS2K_CHECK_LAUNCH();
        fs2k_launch(attn_bwd_dkv_kernel<128>, dim3(dim3(cdiv(L, 64), H, B)), dim3(256), smem_dkv, s, qkv, dout, lse, delta, lens, L, H, scale, dropout_p, (unsigned long long)seed, order, dqkv);
    } else {
        e = cudaFuncSetAttribute(attn_bwd_dq_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dq);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_dkv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dkv);
        if (e != cudaSuccess) return fs2k_set_cuda_error(e);
        fs2k_launch(attn_bwd_dq_kernel<64>, dim3(dim3(cdiv(L, 64), H, B)), dim3(256), smem_dq, s, qkv, dout, lse, delta, lens, L, H, scale, dropout_p, (unsigned long long)seed, order, dqkv);
        FS2K_CHECK_LAUNCH();
        fs2k_launch(attn_bwd_dkv_kernel<64>, dim3(dim3(cdiv(L, 64), H, B)), dim3(256), smem_dkv, s, qkv, dout, lse, delta, lens, L, H, scale, dropout_p, (unsigned long long)seed, order, dqkv);
    }
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

FS2K_DEFINE_SEED_BASE_SETTER(attention_bwd)
