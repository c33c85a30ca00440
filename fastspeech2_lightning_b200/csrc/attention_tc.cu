// Masked multi-head self-attention on the 5th-generation tensor cores (bf16 arithmetic mode):
//
//   nn.MultiheadAttention(q=k=v=LN(x), key_padding_mask, need_weights=False)   torchaudio conformer.py:151-153,193-202
//
// forward and backward, head_dim = 128.  Every product of the flash-attention recurrences is a tcgen05.mma kind::f16 with
// bf16 operands and an fp32 accumulator in TMEM; softmax statistics, the exponentials, the normaliser and dS stay fp32
// in registers.  The packed projection qkv [B, L, 3·H·128] is bf16 in HBM (written by the in-projection GEMM's
// epilogue) and every operand tile — Q, K, V, dO: 128 rows × 128 columns — is brought in by TMA as two 128-byte-swizzled
// boxes; the SAME tile serves as a K-major operand (contraction over its columns: Q·Kᵀ, dO·Vᵀ) and as an MN-major one
// (contraction over its rows: P·V, dS·K, Pᵀ·dO, dSᵀ·Q) just by changing the shared-memory descriptor, so nothing is
// ever transposed.  P and dS are written by the softmax warps (thread = query row = TMEM lane) straight into that
// layout.
//
// The phases of a tile visit — MMA (S, dP) → softmax warps → MMA (P·V / dQ / dV, dK) — are chained by mbarriers; one
// CTA per SM holds all its tiles in shared memory, TMA loads of the next tile overlap the softmax phase.  At these
// sizes (a 128 × 128 × 128 product is 512 tensor-pipe cycles) the kernels are bound by the exponentials, not by the
// pipeline depth, and replace FFMA kernels that needed 32 768 FMA-pipe cycles for the same tile.
//
//   forward : CTA = (q tile, h, b).  Two passes over the key tiles: row maxima first (S only), then P = exp2(S − m),
//             O += P·V, l += Σp — no rescaling of the TMEM accumulator is ever needed.  out = O / l, lse = m + ln l.
//   dq      : CTA = (q tile, h, b), loop over key tiles: dQ += dS·K.
//   dkv     : CTA = (key tile, h, b), loop over query tiles: dV += Pᵀ·dO, dK += dSᵀ·Q.   (deterministic: no atomics)
// Attention-probability dropout: one 64-bit counter hash decides four consecutive keys (16 bits each); the backward
// kernels regenerate the same mask.
#include "bf16_common.cuh"

namespace fs2k {

constexpr int AT = 128;                    // tile rows (queries / keys) == head_dim
constexpr uint32_t AT_BLK = 16384;         // one 64-column block of a tile: 128 rows × 128 B
constexpr uint32_t AT_TILE = 2 * AT_BLK;   // 128 × 128 bf16
constexpr int AT_THREADS = 192;            // warp 0 TMA, warp 1 MMA + TMEM owner, warps 2..5 softmax (thread = row)
constexpr float AT_LOG2E = 1.4426950408889634f, AT_LN2 = 0.6931471805599453f;

__device__ __forceinline__ void at_load_tile(uint32_t dst, const CUtensorMap* map, uint32_t bar, int col0, int row0, int b) {
    tma_load_3d(dst, map, bar, col0, row0, b);
    tma_load_3d(dst + AT_BLK, map, bar, col0 + 64, row0, b);
}
// operand rows = tile rows, contraction over the tile's 128 columns; k-step kk = 16 columns
__device__ __forceinline__ uint64_t at_desc_k(uint32_t base, int kk) { return hb_desc_k(base + (uint32_t)(kk >> 2) * AT_BLK + (uint32_t)(kk & 3) * 32u); }
// contraction over the tile's rows (k-step = 16 rows), operand M/N index = the tile's 128 columns
__device__ __forceinline__ uint64_t at_desc_mn(uint32_t base, int kk) { return hb_desc_mn(base + (uint32_t)kk * 2048u, AT_BLK); }

__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// four consecutive keys share one hash: 16 bits each against thr16 = p·65536
__device__ __forceinline__ unsigned long long at_hash4(unsigned long long seed, int b, int h, int q, int k4, int H, int L) {
    return hash_u64(seed, (((unsigned long long)b * H + h) * L + q) * (unsigned long long)((L + 3) >> 2) + k4);
}

// eight fp32 → one 16-byte chunk of bf16 at columns [c, c+8) of row r of a P / dS tile
__device__ __forceinline__ void at_store8(uint8_t* tile, int r, int c, const float* v) {
    uint4 o;
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    const __nv_bfloat162 c2 = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    o.x = *reinterpret_cast<const uint32_t*>(&a); o.y = *reinterpret_cast<const uint32_t*>(&b);
    o.z = *reinterpret_cast<const uint32_t*>(&c2); o.w = *reinterpret_cast<const uint32_t*>(&d);
    *reinterpret_cast<uint4*>(tile + (uint32_t)(c >> 6) * AT_BLK + (uint32_t)r * 128u + (uint32_t)((((c & 63) >> 3) ^ (r & 7)) << 4)) = o;
}

struct AtBars {
    uint64_t a_full, b_full, b_free, c_full, c_free, s_full, s_free, p_full, p_free, acc_full;
};

__device__ __forceinline__ void at_setup(AtBars* bars, uint32_t* tmem_slot, int warp, int lane, uint32_t tmem_cols) {
    if (warp == 0 && lane == 0) {
        mbar_init(smem_u32(&bars->a_full), 1);
        mbar_init(smem_u32(&bars->b_full), 1);
        mbar_init(smem_u32(&bars->b_free), 1);
        mbar_init(smem_u32(&bars->c_full), 1);
        mbar_init(smem_u32(&bars->c_free), 1);
        mbar_init(smem_u32(&bars->s_full), 1);
        mbar_init(smem_u32(&bars->s_free), 128);
        mbar_init(smem_u32(&bars->p_full), 128);
        mbar_init(smem_u32(&bars->p_free), 1);
        mbar_init(smem_u32(&bars->acc_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
}

__device__ __forceinline__ void at_teardown(uint32_t tmem_base, int warp, uint32_t tmem_cols) {
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ======================================================================================= forward
// smem tiles: Q, K, V, P.  TMEM: S [0,128), O [128,256).
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const int* __restrict__ lens, int L, int H, float scale,
                        float drop_p, unsigned long long seed_in, const int* __restrict__ order, float* __restrict__ out,
                        float* __restrict__ lse_out) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    __shared__ AtBars bars;
    __shared__ uint32_t s_tmem;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sQ = smem, *sK = smem + AT_TILE, *sV = smem + 2 * AT_TILE, *sP = smem + 3 * AT_TILE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    at_setup(&bars, &s_tmem, warp, lane, 256);
    const uint32_t tmem = s_tmem, tmem_S = tmem, tmem_O = tmem + 128;
    pdl_wait();
    lens = pdl_acquire(lens);
    order = pdl_acquire(order);

    const int b = order ? order[blockIdx.z] : blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AT;
    const int D = H * AT;
    const int len = min(lens[b], L);
    const int n = (len + AT - 1) / AT;  // key tiles holding at least one valid key

    if (warp == 0) {
        if (lane == 0 && n > 0) {
            mbar_expect_tx(smem_u32(&bars.a_full), AT_TILE);
            at_load_tile(smem_u32(sQ), &tmQKV, smem_u32(&bars.a_full), h * AT, q0, b);
            for (int t = 0; t < 2 * n; ++t) {
                const int j = t % n;
                mbar_wait(smem_u32(&bars.b_free), (t & 1) ^ 1);
                mbar_expect_tx(smem_u32(&bars.b_full), AT_TILE);
                at_load_tile(smem_u32(sK), &tmQKV, smem_u32(&bars.b_full), D + h * AT, j * AT, b);
                if (t >= n) {
                    const int u = t - n;
                    mbar_wait(smem_u32(&bars.c_free), (u & 1) ^ 1);
                    mbar_expect_tx(smem_u32(&bars.c_full), AT_TILE);
                    at_load_tile(smem_u32(sV), &tmQKV, smem_u32(&bars.c_full), 2 * D + h * AT, j * AT, b);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && n > 0) {
            const uint32_t id_kk = hb_idesc(AT, 0, 0), id_kmn = hb_idesc(AT, 0, 1);
            mbar_wait(smem_u32(&bars.a_full), 0);
            for (int t = 0; t < 2 * n; ++t) {
                mbar_wait(smem_u32(&bars.b_full), t & 1);
                mbar_wait(smem_u32(&bars.s_free), (t & 1) ^ 1);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)  // S = Q·Kᵀ
                    hb_mma(tmem_S, at_desc_k(smem_u32(sQ), kk), at_desc_k(smem_u32(sK), kk), id_kk, kk ? 1u : 0u);
                tc_commit(smem_u32(&bars.s_full));
                tc_commit(smem_u32(&bars.b_free));
                if (t >= n) {
                    const int u = t - n;
                    mbar_wait(smem_u32(&bars.c_full), u & 1);
                    mbar_wait(smem_u32(&bars.p_full), u & 1);
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)  // O += P·V
                        hb_mma(tmem_O, at_desc_k(smem_u32(sP), kk), at_desc_mn(smem_u32(sV), kk), id_kmn, (u | kk) ? 1u : 0u);
                    tc_commit(smem_u32(&bars.c_free));
                    tc_commit(smem_u32(&bars.p_free));
                }
            }
            tc_commit(smem_u32(&bars.acc_full));
        }
    } else {
        // ================= softmax warps: thread = query row = TMEM lane =================
        const int quad = warp & 3, r = quad * 32 + lane, q = q0 + r;
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        const float sl2 = scale * AT_LOG2E;
        const unsigned long long seed = drop_p > 0.f ? seed_with_base(seed_in) : 0ull;
        const uint32_t thr16 = (uint32_t)(drop_p * 65536.0f);
        const float keep_scale = 65536.0f / (float)(65536u - thr16);
        float m2 = -INFINITY, l = 0.f;
        for (int t = 0; t < 2 * n; ++t) {
            const int j = t % n, kvalid = min(AT, len - j * AT);
            mbar_wait(smem_u32(&bars.s_full), t & 1);
            tc_fence_after();
            if (t < n) {  // pass 1: row maximum of the scaled scores
#pragma unroll 1
                for (int c0 = 0; c0 < AT; c0 += 32) {
                    float v[32];
                    tc_ld32(tmem_S + lane_off + (uint32_t)c0, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c0 + i < kvalid) m2 = fmaxf(m2, v[i] * sl2);
                }
            } else {  // pass 2: P = exp2(s − m), l += Σ p, dropped P → shared memory (bf16)
                const int u = t - n;
                mbar_wait(smem_u32(&bars.p_free), (u & 1) ^ 1);
#pragma unroll 1
                for (int c0 = 0; c0 < AT; c0 += 32) {
                    float v[32];
                    tc_ld32(tmem_S + lane_off + (uint32_t)c0, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float p = (c0 + i < kvalid) ? exp2f(v[i] * sl2 - m2) : 0.f;
                        l += p;
                        v[i] = p;
                    }
                    if (thr16) {
#pragma unroll
                        for (int g = 0; g < 8; ++g) {
                            const unsigned long long hsh = at_hash4(seed, b, h, q, (j * AT + c0) / 4 + g, H, L);
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                v[g * 4 + e] = ((uint32_t)(hsh >> (16 * e)) & 0xFFFFu) >= thr16 ? v[g * 4 + e] * keep_scale : 0.f;
                        }
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g) at_store8(sP, r, c0 + 8 * g, v + 8 * g);
                }
                fence_async_smem();
                mbar_arrive(smem_u32(&bars.p_full));
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&bars.s_free));
        }
        // ---- out = O / l, lse = m + ln l ----
        if (n > 0) {
            mbar_wait(smem_u32(&bars.acc_full), 0);
            tc_fence_after();
            const float inv_l = 1.0f / l;
            float* dst = out + ((size_t)b * L + q) * D + h * AT;
#pragma unroll 1
            for (int c0 = 0; c0 < AT; c0 += 32) {
                float v[32];
                tc_ld32(tmem_O + lane_off + (uint32_t)c0, v);
                if (q < L) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *reinterpret_cast<float4*>(dst + c0 + i) = make_float4(v[i] * inv_l, v[i + 1] * inv_l, v[i + 2] * inv_l, v[i + 3] * inv_l);
                }
            }
            if (lse_out && q < L) lse_out[((size_t)b * H + h) * L + q] = m2 * AT_LN2 + logf(l);
        } else if (q < L) {
            float* dst = out + ((size_t)b * L + q) * D + h * AT;
            for (int c = 0; c < AT; c += 4) *reinterpret_cast<float4*>(dst + c) = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lse_out) lse_out[((size_t)b * H + h) * L + q] = 0.f;
        }
    }
    at_teardown(tmem, warp, 256);
}

// P (dropped) and dS of one 128 × 128 visit, row r of the softmax thread: reads S [tmem+0) and dP [tmem+128)
template <bool WRITE_P>
__device__ __forceinline__ void at_bwd_rows(uint32_t tmem, uint32_t lane_off, uint8_t* sP, uint8_t* sdS, int r, bool row_ok, float lse2,
                                            float delta, int kvalid, float sl2, float scale, uint32_t thr16, float keep_scale,
                                            unsigned long long seed, int b, int h, int q, int key0, int H, int L) {
#pragma unroll 1
    for (int c0 = 0; c0 < AT; c0 += 32) {
        float s[32], dp[32];
        tc_ld32(tmem + lane_off + (uint32_t)c0, s);
        tc_ld32(tmem + 128 + lane_off + (uint32_t)c0, dp);
#pragma unroll
        for (int i = 0; i < 32; ++i) s[i] = (row_ok && c0 + i < kvalid) ? exp2f(s[i] * sl2 - lse2) : 0.f;  // P
        if (thr16) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const unsigned long long hsh = at_hash4(seed, b, h, q, (key0 + c0) / 4 + g, H, L);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float k = ((uint32_t)(hsh >> (16 * e)) & 0xFFFFu) >= thr16 ? keep_scale : 0.f;
                    dp[g * 4 + e] *= k;                       // dP = dP_dropped · keep
                    const float p = s[g * 4 + e];
                    s[g * 4 + e] = p * k;                     // dropped P (feeds dV)
                    dp[g * 4 + e] = p * (dp[g * 4 + e] - delta) * scale;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) dp[i] = s[i] * (dp[i] - delta) * scale;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            if (WRITE_P) at_store8(sP, r, c0 + 8 * g, s + 8 * g);
            at_store8(sdS, r, c0 + 8 * g, dp + 8 * g);
        }
    }
}

// ======================================================================================= dQ
// smem tiles: Q, dO (resident), K, V (per visit), dS.  TMEM: S [0,128), dP [128,256), dQ [256,384).
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc_dq_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                       const float* __restrict__ lse, const float* __restrict__ delta, const int* __restrict__ lens, int L, int H,
                       float scale, float drop_p, unsigned long long seed_in, const int* __restrict__ order, float* __restrict__ dqkv) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    __shared__ AtBars bars;
    __shared__ uint32_t s_tmem;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sQ = smem, *sdO = smem + AT_TILE, *sK = smem + 2 * AT_TILE, *sV = smem + 3 * AT_TILE, *sdS = smem + 4 * AT_TILE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    at_setup(&bars, &s_tmem, warp, lane, 512);
    const uint32_t tmem = s_tmem;
    pdl_wait();
    lens = pdl_acquire(lens);
    order = pdl_acquire(order);
    lse = pdl_acquire(lse);
    delta = pdl_acquire(delta);

    const int b = order ? order[blockIdx.z] : blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AT;
    const int D = H * AT;
    const int len = min(lens[b], L);
    const int n = (len + AT - 1) / AT;

    if (warp == 0) {
        if (lane == 0 && n > 0) {
            mbar_expect_tx(smem_u32(&bars.a_full), 2 * AT_TILE);
            at_load_tile(smem_u32(sQ), &tmQKV, smem_u32(&bars.a_full), h * AT, q0, b);
            at_load_tile(smem_u32(sdO), &tmDO, smem_u32(&bars.a_full), h * AT, q0, b);
            for (int t = 0; t < n; ++t) {
                mbar_wait(smem_u32(&bars.b_free), (t & 1) ^ 1);
                mbar_expect_tx(smem_u32(&bars.b_full), 2 * AT_TILE);
                at_load_tile(smem_u32(sK), &tmQKV, smem_u32(&bars.b_full), D + h * AT, t * AT, b);
                at_load_tile(smem_u32(sV), &tmQKV, smem_u32(&bars.b_full), 2 * D + h * AT, t * AT, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && n > 0) {
            const uint32_t id_kk = hb_idesc(AT, 0, 0), id_kmn = hb_idesc(AT, 0, 1);
            mbar_wait(smem_u32(&bars.a_full), 0);
            for (int t = 0; t < n; ++t) {
                mbar_wait(smem_u32(&bars.b_full), t & 1);
                mbar_wait(smem_u32(&bars.s_free), (t & 1) ^ 1);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)  // S = Q·Kᵀ
                    hb_mma(tmem, at_desc_k(smem_u32(sQ), kk), at_desc_k(smem_u32(sK), kk), id_kk, kk ? 1u : 0u);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)  // dP = dO·Vᵀ
                    hb_mma(tmem + 128, at_desc_k(smem_u32(sdO), kk), at_desc_k(smem_u32(sV), kk), id_kk, kk ? 1u : 0u);
                tc_commit(smem_u32(&bars.s_full));
                mbar_wait(smem_u32(&bars.p_full), t & 1);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)  // dQ += dS·K
                    hb_mma(tmem + 256, at_desc_k(smem_u32(sdS), kk), at_desc_mn(smem_u32(sK), kk), id_kmn, (t | kk) ? 1u : 0u);
                tc_commit(smem_u32(&bars.b_free));
                tc_commit(smem_u32(&bars.p_free));
            }
            tc_commit(smem_u32(&bars.acc_full));
        }
    } else {
        const int quad = warp & 3, r = quad * 32 + lane, q = q0 + r;
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        const float sl2 = scale * AT_LOG2E;
        const unsigned long long seed = drop_p > 0.f ? seed_with_base(seed_in) : 0ull;
        const uint32_t thr16 = (uint32_t)(drop_p * 65536.0f);
        const float keep_scale = 65536.0f / (float)(65536u - thr16);
        const bool row_ok = q < L;
        const float lse2 = row_ok ? lse[((size_t)b * H + h) * L + q] * AT_LOG2E : 0.f;
        const float dl = row_ok ? delta[((size_t)b * H + h) * L + q] : 0.f;
        for (int t = 0; t < n; ++t) {
            mbar_wait(smem_u32(&bars.s_full), t & 1);
            tc_fence_after();
            mbar_wait(smem_u32(&bars.p_free), (t & 1) ^ 1);
            at_bwd_rows<false>(tmem, lane_off, nullptr, sdS, r, row_ok, lse2, dl, min(AT, len - t * AT), sl2, scale, thr16, keep_scale, seed,
                               b, h, q, t * AT, H, L);
            fence_async_smem();
            mbar_arrive(smem_u32(&bars.p_full));
            tc_fence_before();
            mbar_arrive(smem_u32(&bars.s_free));
        }
        float* dst = dqkv + ((size_t)b * L + q) * 3 * D + h * AT;
        if (n > 0) {
            mbar_wait(smem_u32(&bars.acc_full), 0);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < AT; c0 += 32) {
                float v[32];
                tc_ld32(tmem + 256 + lane_off + (uint32_t)c0, v);
                if (row_ok) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
            }
        } else if (row_ok) {
            for (int c = 0; c < AT; c += 4) *reinterpret_cast<float4*>(dst + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    at_teardown(tmem, warp, 512);
}

// ======================================================================================= dK, dV
// smem tiles: K, V (resident), Q, dO (per visit), P, dS.  TMEM: S [0,128), dP [128,256), dV [256,384), dK [384,512).
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc_dkv_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                        const float* __restrict__ lse, const float* __restrict__ delta, const int* __restrict__ lens, int L, int H,
                        float scale, float drop_p, unsigned long long seed_in, const int* __restrict__ order, float* __restrict__ dqkv) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    __shared__ AtBars bars;
    __shared__ uint32_t s_tmem;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sK = smem, *sV = smem + AT_TILE, *sQ = smem + 2 * AT_TILE, *sdO = smem + 3 * AT_TILE, *sP = smem + 4 * AT_TILE, *sdS = smem + 5 * AT_TILE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    at_setup(&bars, &s_tmem, warp, lane, 512);
    const uint32_t tmem = s_tmem;
    pdl_wait();
    lens = pdl_acquire(lens);
    order = pdl_acquire(order);
    lse = pdl_acquire(lse);
    delta = pdl_acquire(delta);

    const int b = order ? order[blockIdx.z] : blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * AT;
    const int D = H * AT;
    const int len = min(lens[b], L);
    const int n = (k0 < len) ? (L + AT - 1) / AT : 0;  // query tiles (all rows < L: padded queries carry gradient too)
    const int kvalid = min(AT, len - k0);

    if (warp == 0) {
        if (lane == 0 && n > 0) {
            mbar_expect_tx(smem_u32(&bars.a_full), 2 * AT_TILE);
            at_load_tile(smem_u32(sK), &tmQKV, smem_u32(&bars.a_full), D + h * AT, k0, b);
            at_load_tile(smem_u32(sV), &tmQKV, smem_u32(&bars.a_full), 2 * D + h * AT, k0, b);
            for (int t = 0; t < n; ++t) {
                mbar_wait(smem_u32(&bars.b_free), (t & 1) ^ 1);
                mbar_expect_tx(smem_u32(&bars.b_full), 2 * AT_TILE);
                at_load_tile(smem_u32(sQ), &tmQKV, smem_u32(&bars.b_full), h * AT, t * AT, b);
                at_load_tile(smem_u32(sdO), &tmDO, smem_u32(&bars.b_full), h * AT, t * AT, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && n > 0) {
            const uint32_t id_kk = hb_idesc(AT, 0, 0), id_mm = hb_idesc(AT, 1, 1);
            mbar_wait(smem_u32(&bars.a_full), 0);
            for (int t = 0; t < n; ++t) {
                mbar_wait(smem_u32(&bars.b_full), t & 1);
                mbar_wait(smem_u32(&bars.s_free), (t & 1) ^ 1);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)  // S = Q·Kᵀ   [q × keys]
                    hb_mma(tmem, at_desc_k(smem_u32(sQ), kk), at_desc_k(smem_u32(sK), kk), id_kk, kk ? 1u : 0u);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)  // dP = dO·Vᵀ
                    hb_mma(tmem + 128, at_desc_k(smem_u32(sdO), kk), at_desc_k(smem_u32(sV), kk), id_kk, kk ? 1u : 0u);
                tc_commit(smem_u32(&bars.s_full));
                mbar_wait(smem_u32(&bars.p_full), t & 1);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)  // dV += Pᵀ·dO   (contraction over the query rows of both tiles)
                    hb_mma(tmem + 256, at_desc_mn(smem_u32(sP), kk), at_desc_mn(smem_u32(sdO), kk), id_mm, (t | kk) ? 1u : 0u);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)  // dK += dSᵀ·Q
                    hb_mma(tmem + 384, at_desc_mn(smem_u32(sdS), kk), at_desc_mn(smem_u32(sQ), kk), id_mm, (t | kk) ? 1u : 0u);
                tc_commit(smem_u32(&bars.b_free));
                tc_commit(smem_u32(&bars.p_free));
            }
            tc_commit(smem_u32(&bars.acc_full));
        }
    } else {
        const int quad = warp & 3, r = quad * 32 + lane;
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        const float sl2 = scale * AT_LOG2E;
        const unsigned long long seed = drop_p > 0.f ? seed_with_base(seed_in) : 0ull;
        const uint32_t thr16 = (uint32_t)(drop_p * 65536.0f);
        const float keep_scale = 65536.0f / (float)(65536u - thr16);
        for (int t = 0; t < n; ++t) {
            const int q = t * AT + r;
            const bool row_ok = q < L;
            const float lse2 = row_ok ? lse[((size_t)b * H + h) * L + q] * AT_LOG2E : 0.f;
            const float dl = row_ok ? delta[((size_t)b * H + h) * L + q] : 0.f;
            mbar_wait(smem_u32(&bars.s_full), t & 1);
            tc_fence_after();
            mbar_wait(smem_u32(&bars.p_free), (t & 1) ^ 1);
            at_bwd_rows<true>(tmem, lane_off, sP, sdS, r, row_ok, lse2, dl, kvalid, sl2, scale, thr16, keep_scale, seed, b, h, q, k0, H, L);
            fence_async_smem();
            mbar_arrive(smem_u32(&bars.p_full));
            tc_fence_before();
            mbar_arrive(smem_u32(&bars.s_free));
        }
        // TMEM lane = key row of this tile
        const int key = k0 + r;
        float* dK = dqkv + ((size_t)b * L + key) * 3 * D + D + h * AT;
        float* dV = dK + D;
        if (n > 0) {
            mbar_wait(smem_u32(&bars.acc_full), 0);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < AT; c0 += 32) {
                float v[32];
                tc_ld32(tmem + 256 + lane_off + (uint32_t)c0, v);
                if (key < L) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dV + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
                tc_ld32(tmem + 384 + lane_off + (uint32_t)c0, v);
                if (key < L) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dK + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
            }
        } else if (key < L) {
            for (int c = 0; c < AT; c += 4) {
                *reinterpret_cast<float4*>(dK + c) = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(dV + c) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    at_teardown(tmem, warp, 512);
}

__global__ void __launch_bounds__(256)
attn_delta_kernel_tc(const float* __restrict__ o, const float* __restrict__ dO, int B, int L, int H, int HD, float* __restrict__ delta) {
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

}  // namespace fs2k

using namespace fs2k;

static bool at_make_map(CUtensorMap* tm, const void* base, int cols, int L, int B) {
    EncodeTiledFn encode = get_encode();
    if (!encode) return false;
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)L, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)L * cols * 2};
    cuuint32_t box[3] = {64, AT, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

extern "C" int fs2k_attention_bf16(const void* qkv_bf16, const int* lens, int B, int L, int H, int head_dim, float dropout_p,
                                   long seed, float* out, float* lse_out, const int* order, fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && L >= 0 && H > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(head_dim == AT, FS2K_ERR_UNSUPPORTED);
    FS2K_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, FS2K_ERR_BAD_SHAPE);
    if (B == 0 || L == 0) return FS2K_OK;
    FS2K_REQUIRE(qkv_bf16 && lens && out, FS2K_ERR_NULL);
    FS2K_REQUIRE(B <= 65535 && H <= 65535, FS2K_ERR_UNSUPPORTED);
    CUtensorMap tm;
    if (!at_make_map(&tm, qkv_bf16, 3 * H * AT, L, B)) return fs2k_set_cuda_error(cudaErrorInvalidValue);
    const size_t smem = 4 * AT_TILE + 1024;
    cudaError_t e = cudaFuncSetAttribute(attention_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    const float scale = 1.0f / sqrtf((float)head_dim);
    fs2k_launch(attention_tc_fwd_kernel, dim3(cdiv(L, AT), H, B), dim3(AT_THREADS), smem, (cudaStream_t)stream, tm, lens, L, H, scale,
                dropout_p, (unsigned long long)seed, order, out, lse_out);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_attention_bwd_bf16(const void* qkv_bf16, const float* out, const float* lse, const float* dout,
                                       const void* dout_bf16, const int* lens, int B, int L, int H, int head_dim, float dropout_p,
                                       long seed, float* delta, float* dqkv, const int* order, fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && L >= 0 && H > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(head_dim == AT, FS2K_ERR_UNSUPPORTED);
    if (B == 0 || L == 0) return FS2K_OK;
    FS2K_REQUIRE(qkv_bf16 && out && lse && dout && dout_bf16 && lens && delta && dqkv, FS2K_ERR_NULL);
    FS2K_REQUIRE(B <= 65535 && H <= 65535, FS2K_ERR_UNSUPPORTED);
    cudaStream_t s = (cudaStream_t)stream;
    CUtensorMap tmQ, tmDO;
    if (!at_make_map(&tmQ, qkv_bf16, 3 * H * AT, L, B) || !at_make_map(&tmDO, dout_bf16, H * AT, L, B))
        return fs2k_set_cuda_error(cudaErrorInvalidValue);
    long g = ((long)B * L * H + 7) / 8;
    if (g > 148 * 8) g = 148 * 8;
    fs2k_launch(attn_delta_kernel_tc, dim3((int)g), dim3(256), 0, s, out, dout, B, L, H, head_dim, delta);
    FS2K_CHECK_LAUNCH();
    const float scale = 1.0f / sqrtf((float)head_dim);
    const size_t smem_dq = 5 * AT_TILE + 1024, smem_dkv = 6 * AT_TILE + 1024;
    cudaError_t e = cudaFuncSetAttribute(attention_tc_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dq);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_tc_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dkv);
    if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    fs2k_launch(attention_tc_dq_kernel, dim3(cdiv(L, AT), H, B), dim3(AT_THREADS), smem_dq, s, tmQ, tmDO, lse, delta, lens, L, H, scale,
                dropout_p, (unsigned long long)seed, order, dqkv);
    FS2K_CHECK_LAUNCH();
    fs2k_launch(attention_tc_dkv_kernel, dim3(cdiv(L, AT), H, B), dim3(AT_THREADS), smem_dkv, s, tmQ, tmDO, lse, delta, lens, L, H, scale,
                dropout_p, (unsigned long long)seed, order, dqkv);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

FS2K_DEFINE_SEED_BASE_SETTER(attention_tc)
