// Tensor-core GEMM / 1-D convolution for sm_100a: TMA → shared memory → tcgen05.mma (kind::tf32,
// accumulators in TMEM) → fused epilogue.  Same contract as fs2k_gemm_f32 (gemm_simt.cu):
//
//   C[(b,l), n] = (dropout(act((Σ_tap Σ_k A[b, l+tap−pad, k] · W[tap][n][k] + bias[n])·scale[n] + shift[n])·alpha)
//                 + residual[(b,l), n]) · row_mask[(b,l)]          (+ optional LayerNorm of the finished row)
//
// One 128 × BLOCK_N output tile per CTA (BLOCK_N = N when N ≤ 256, else 256).  Operands stay fp32 in
// HBM: TMA brings 128-row × 32-float boxes (one 128-byte swizzle atom per row) into a ring of 2 (3×TF32) to 4 stages; the
// convolution taps are time-shifted boxes of a 3-D tensor map [B][L][K] whose out-of-range rows TMA
// zero-fills, so a tap never crosses an utterance boundary and no im2col buffer exists.  One elected
// thread issues four K=8 tf32 MMAs per stage.  `passes = 3` adds the split-accumulate 3×TF32 scheme
// (big·big + big·small + small·big, small = x − tf32(x)) that restores fp32-level accuracy.
//
// Warp roles (12 warps): 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4-11 = operand splitter
// during the main loop (3×TF32 only), then TMEM → registers → raw staging tile in shared memory; finally
// all twelve warps run the store pass: per-column math (bias, folded BatchNorm, activation, alpha),
// residual, row mask and, when requested, the LayerNorm of the row (a row lives in one warp, so mean and
// variance are two shuffle reductions), with fully coalesced 16-byte global accesses.  Dropout and the LayerNorm tail are
// template flags: anything living in the store loop costs every launch.
#include "tc_common.cuh"

namespace fs2k {

constexpr int TC_BM = 128;      // rows per tile == TMEM lanes
// K elements per pipeline stage: 32 floats = one 128-byte swizzle atom per row.  3×TF32 keeps a second ("small")
// copy of every tile, so a stage costs 96 KB and only two fit; the ring is then latency-bound (stage cycle = TMA
// latency + split + MMA ≈ 2.7 µs against 0.85 µs of MMA work).  Half-width stages (16 floats, 64-byte swizzle —
// set value = 16 below, the descriptors and tensor maps follow) give four stages in the same shared memory, but were
// measured SLOWER (main loop 10.9 µs vs 9.7 µs at K = 256, N = 256): the per-stage fixed costs (barrier round trips,
// splitter hand-off) do not shrink with the stage.
template <int PASSES> struct TcBK { static constexpr int value = 32; };
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_THREADS = 384;

struct TcEpilogue {
    const float* bias; const float* scale; const float* shift;
    int act; float alpha;
    const float* residual; int ldr;
    const uint8_t* row_mask;
    float* C; int ldc;
    // optional fused LayerNorm of the finished row (requires BLOCK_N == N)
    const float* ln_gamma; const float* ln_beta; float ln_eps; float* ln_out; int ld_ln;
    const float* ln2_gamma; const float* ln2_beta; float* ln2_out;  // second LN chained on ln_out
    float drop_p; unsigned long long seed;  // dropout of the value before the residual add (mask = hash(seed, m·N + n))
};

// K-major operand whose rows are one swizzle atom wide (BK·4 = 128 or 64 bytes): 8-row groups are 8·row bytes apart
// (SBO), descriptor version 1, layout type SWIZZLE_128B (2) or SWIZZLE_64B (4)
template <int BK>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr) {
    static_assert(BK == 32 || BK == 16, "row = one 128-byte or 64-byte swizzle atom");
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);          // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                           // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)((8 * BK * 4) >> 4) << 32;         // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
    d |= (uint64_t)(BK == 32 ? 2 : 4) << 61;          // SWIZZLE_128B / SWIZZLE_64B
    return d;
}
// instruction descriptor: D = f32, A = B = tf32, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t make_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

__device__ __forceinline__ float tc_act(float v, int act) {
    if (act == FS2K_ACT_RELU) return fmaxf(v, 0.f);
    if (act == FS2K_ACT_SILU) return silu(v);
    if (act == FS2K_ACT_TANH) return tanhf(v);
    return v;
}

// finished-row math shared by the store pass: v = act((acc + bias)·scale + shift)·alpha
__device__ __forceinline__ float4 tc_colmath(float4 a, const float4& bias, const float4& sc, const float4& sh, int act, float alpha) {
    a.x = tc_act((a.x + bias.x) * sc.x + sh.x, act) * alpha;
    a.y = tc_act((a.y + bias.y) * sc.y + sh.y, act) * alpha;
    a.z = tc_act((a.z + bias.z) * sc.z + sh.z, act) * alpha;
    a.w = tc_act((a.w + bias.w) * sc.w + sh.w, act) * alpha;
    return a;
}
__device__ __forceinline__ float4 tc_ln_apply(const float4& v, float mean, float rstd, const float4& g, const float4& b) {
    float4 o;
    o.x = (v.x - mean) * rstd * g.x + b.x;
    o.y = (v.y - mean) * rstd * g.y + b.y;
    o.z = (v.z - mean) * rstd * g.z + b.z;
    o.w = (v.w - mean) * rstd * g.w + b.w;
    return o;
}
__device__ __forceinline__ float tc_sum4(const float4& v) { return (v.x + v.y) + (v.z + v.w); }
__device__ __forceinline__ float tc_sq4(const float4& v, float m) {
    const float a = v.x - m, b = v.y - m, c = v.z - m, d = v.w - m;
    return (a * a + b * b) + (c * c + d * d);
}

constexpr int TC_EPI_WARPS = 8;                 // warps 4..11: operand split (3×TF32) and TMEM → staging
constexpr int TC_WARPS = TC_THREADS / 32;       // 12
template <bool LNORM> struct TcRowsPerIter { static constexpr int value = 1; };  // rows a warp keeps in flight in the store pass (measured: 1 beats 2, 3, 4, 6, 8 — 17.6 vs 18.0 (4) vs 19.5 us (8) per GEMM)

template <int PASSES, bool DROPOUT, bool LNORM>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmBs, int presplit_b, int L, long M_total, int K, int N, int block_n, int taps, int pad, int tiles_per_b, int n_stages,
               TcEpilogue ep) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    constexpr int TC_BK = TcBK<PASSES>::value;
    const int TC_STAGES = n_stages;  // 2..8, as many as fit in shared memory (host decides)
    __shared__ __align__(8) uint64_t s_full[TC_MAX_STAGES], s_empty[TC_MAX_STAGES], s_split[TC_MAX_STAGES], s_tmem_full;
    __shared__ uint32_t s_tmem_base;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t a_bytes = TC_BM * TC_BK * 4, b_bytes = block_n * TC_BK * 4;
    const uint32_t stage_bytes = (a_bytes + b_bytes) * (PASSES == 3 ? 2 : 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int tile_m = blockIdx.x, n0 = blockIdx.y * block_n;
    const int b_idx = tile_m / tiles_per_b, l0 = (tile_m % tiles_per_b) * TC_BM;
    const int nk = (K + TC_BK - 1) / TC_BK;
    const int iters = taps * nk;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(smem_u32(&s_full[s]), 1);
            mbar_init(smem_u32(&s_empty[s]), 1);
            mbar_init(smem_u32(&s_split[s]), TC_EPI_WARPS * 32);
        }
        mbar_init(smem_u32(&s_tmem_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        // 256 accumulator columns cover every BLOCK_N ≤ 256 (power of two ≥ 32 required)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&s_tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;
    pdl_wait();  // barrier init + TMEM allocation above overlapped the previous kernel's tail; memory is touched below

    if (warp == 0) {
        if (lane == 0) {
            // ================= TMA producer =================
            for (int it = 0; it < iters; ++it) {
                const int s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
                mbar_wait(smem_u32(&s_empty[s]), ph ^ 1);
                const int tap = it / nk, k0 = (it - tap * nk) * TC_BK;
                const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes), sb = sa + a_bytes;
                const uint32_t bar = smem_u32(&s_full[s]);
                // presplit_b (3×TF32): the weights' small parts were computed once per weight version in global memory;
                // TMA brings them straight into the small slot and the splitter warps only handle the A tile
                const bool pre = PASSES == 3 && presplit_b;
                mbar_expect_tx(bar, a_bytes + b_bytes + (pre ? b_bytes : 0u));
                tma_load_3d(sa, &tmA, bar, k0, l0 + tap - pad, b_idx);
                tma_load_2d(sb, &tmB, bar, k0, tap * N + n0);
                if (pre) tma_load_2d(sb + a_bytes + b_bytes, &tmBs, bar, k0, tap * N + n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ================= MMA issuer =================
            const uint32_t idesc = make_idesc_tf32(block_n);
            for (int it = 0; it < iters; ++it) {
                const int s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
                mbar_wait(smem_u32(PASSES == 3 ? &s_split[s] : &s_full[s]), ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes), sb = sa + a_bytes;
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k) {
                    const uint64_t ad = make_smem_desc<TC_BK>(sa + k * 32), bd = make_smem_desc<TC_BK>(sb + k * 32);
                    tc_mma_tf32(tmem_base, ad, bd, idesc, (it | k) ? 1u : 0u);
                    if (PASSES == 3) {
                        const uint64_t ads = make_smem_desc<TC_BK>(sa + a_bytes + b_bytes + k * 32);
                        const uint64_t bds = make_smem_desc<TC_BK>(sb + a_bytes + b_bytes + k * 32);
                        tc_mma_tf32(tmem_base, ad, bds, idesc, 1u);
                        tc_mma_tf32(tmem_base, ads, bd, idesc, 1u);
                    }
                }
                tc_commit(smem_u32(&s_empty[s]));  // frees the stage when these MMAs have read it
            }
            tc_commit(smem_u32(&s_tmem_full));
        }
    } else if (warp >= 4) {
        const int et = threadIdx.x - 128;  // 0..255 among the epilogue warps
        if (PASSES == 3) {
            // ===== operand splitter (3×TF32): small parts of A and W written next to the originals =====
            for (int it = 0; it < iters; ++it) {
                const int s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
                mbar_wait(smem_u32(&s_full[s]), ph);
                const float4* src = reinterpret_cast<const float4*>(smem + (size_t)s * stage_bytes);
                float4* dst = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes + a_bytes + b_bytes);
                split_small(src, dst, (presplit_b ? a_bytes : a_bytes + b_bytes) / 16, et, TC_EPI_WARPS * 32);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes → visible to the MMA proxy
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_split[s])) : "memory");
            }
        }
        // ===== epilogue stage 1: raw accumulators TMEM → registers → staging tile (row = TMEM lane) =====
        mbar_wait(smem_u32(&s_tmem_full), 0);
        tc_fence_after();
        const int q = warp & 3;              // TMEM lane quadrant this warp may read
        const int half = (warp - 4) >> 2;    // two warps share a quadrant and alternate 16-column chunks
        const int row = q * 32 + lane;
        const int pitch = block_n + 4;
        float* stag = reinterpret_cast<float*>(smem);
        for (int c0 = half * 16; c0 < block_n; c0 += 32) {
            float v[16];
            tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
            for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(stag + (size_t)row * pitch + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        tc_fence_before();
    }
    __syncthreads();

    // ===== epilogue stage 2 (all warps): column math, residual, row mask, LayerNorm, coalesced 16-byte stores =====
    {
        const int pitch = block_n + 4;
        const float* stag = reinterpret_cast<const float*>(smem);
        const int nv = block_n >> 2;  // float4 per row (≤ 64): lane owns float4 columns lane and lane+32
        const bool has[2] = {lane < nv, lane + 32 < nv};
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f), one4 = make_float4(1.f, 1.f, 1.f, 1.f);
        float4 bias4[2], sc4[2], sh4[2], g1[2], b1[2], g2[2], b2[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int qv = lane + 32 * j;
            bias4[j] = (has[j] && ep.bias) ? __ldg(reinterpret_cast<const float4*>(ep.bias + n0) + qv) : zero4;
            sc4[j] = (has[j] && ep.scale) ? __ldg(reinterpret_cast<const float4*>(ep.scale + n0) + qv) : one4;
            sh4[j] = (has[j] && ep.scale) ? __ldg(reinterpret_cast<const float4*>(ep.shift + n0) + qv) : zero4;
            g1[j] = (LNORM && has[j]) ? __ldg(reinterpret_cast<const float4*>(ep.ln_gamma) + qv) : one4;
            b1[j] = (LNORM && has[j]) ? __ldg(reinterpret_cast<const float4*>(ep.ln_beta) + qv) : zero4;
            g2[j] = (LNORM && has[j] && ep.ln2_out) ? __ldg(reinterpret_cast<const float4*>(ep.ln2_gamma) + qv) : one4;
            b2[j] = (LNORM && has[j] && ep.ln2_out) ? __ldg(reinterpret_cast<const float4*>(ep.ln2_beta) + qv) : zero4;
        }
        const int rows_valid = (int)min((long)TC_BM, min((long)L - l0, M_total - ((long)b_idx * L + l0)));
        const float inv_n = 1.0f / (float)block_n;
        // (DROPOUT is a template flag: the counter hash must not sit in the store loop of the launches that never drop)
        const float drop_p = DROPOUT ? ep.drop_p : 0.f;
        const uint32_t thr16 = drop_thr16(drop_p);
        const float inv_keep = drop_inv_keep(thr16);
        const unsigned long long seed = DROPOUT ? seed_with_base(ep.seed) : 0ull;
        constexpr int TC_ROWS_PER_ITER = TcRowsPerIter<LNORM>::value;
        for (int rb = warp; rb < rows_valid; rb += TC_WARPS * TC_ROWS_PER_ITER) {
            float4 val[TC_ROWS_PER_ITER][2], res[TC_ROWS_PER_ITER][2];
            float rm[TC_ROWS_PER_ITER];
            // issue every global read of the row group before using any of them
#pragma unroll
            for (int u = 0; u < TC_ROWS_PER_ITER; ++u) {
                const int r = rb + u * TC_WARPS;
                const long m = (long)b_idx * L + l0 + r;
                rm[u] = 1.f;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    res[u][j] = zero4;
                    if (r < rows_valid && has[j] && ep.residual)
                        res[u][j] = *reinterpret_cast<const float4*>(ep.residual + (size_t)m * ep.ldr + n0 + (lane + 32 * j) * 4);
                }
                if (r < rows_valid && ep.row_mask) rm[u] = ep.row_mask[m] ? 1.f : 0.f;
            }
#pragma unroll
            for (int u = 0; u < TC_ROWS_PER_ITER; ++u) {
                const int r = rb + u * TC_WARPS;
                if (r >= rows_valid) break;  // warp-uniform
                const long m = (long)b_idx * L + l0 + r;
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (has[j]) {
                        const int qv = lane + 32 * j;
                        float4 v = *reinterpret_cast<const float4*>(stag + (size_t)r * pitch + qv * 4);
                        v = tc_colmath(v, bias4[j], sc4[j], sh4[j], ep.act, ep.alpha);
                        if (DROPOUT) {
                            const unsigned long long e = (unsigned long long)m * N + n0 + qv * 4;
                            drop_apply4(v, seed, e, thr16, inv_keep);
                        }
                        v.x = (v.x + res[u][j].x) * rm[u]; v.y = (v.y + res[u][j].y) * rm[u];
                        v.z = (v.z + res[u][j].z) * rm[u]; v.w = (v.w + res[u][j].w) * rm[u];
                        if (!LNORM || ep.C) *reinterpret_cast<float4*>(ep.C + (size_t)m * ep.ldc + n0 + qv * 4) = v;
                        if (LNORM) {
                            val[u][j] = v;
                            sum += tc_sum4(v);
                        }
                    } else if (LNORM) {
                        val[u][j] = zero4;
                    }
                }
                if (LNORM) {  // LayerNorm over the block_n == N columns of this row (one warp holds the row)
                    float mean = warp_sum(sum) * inv_n;
                    float ss = 0.f;
#pragma unroll
                    for (int j = 0; j < 2; ++j) if (has[j]) ss += tc_sq4(val[u][j], mean);
                    float rstd = 1.0f / sqrtf(warp_sum(ss) * inv_n + ep.ln_eps);
                    float sum2 = 0.f;
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        if (has[j]) {
                            const float4 o = tc_ln_apply(val[u][j], mean, rstd, g1[j], b1[j]);
                            *reinterpret_cast<float4*>(ep.ln_out + (size_t)m * ep.ld_ln + (lane + 32 * j) * 4) = o;
                            val[u][j] = o;
                            sum2 += tc_sum4(o);
                        }
                    if (ep.ln2_out) {  // second LayerNorm chained on the first one's output
                        mean = warp_sum(sum2) * inv_n;
                        ss = 0.f;
#pragma unroll
                        for (int j = 0; j < 2; ++j) if (has[j]) ss += tc_sq4(val[u][j], mean);
                        rstd = 1.0f / sqrtf(warp_sum(ss) * inv_n + ep.ln_eps);
#pragma unroll
                        for (int j = 0; j < 2; ++j)
                            if (has[j])
                                *reinterpret_cast<float4*>(ep.ln2_out + (size_t)m * ep.ld_ln + (lane + 32 * j) * 4) =
                                    tc_ln_apply(val[u][j], mean, rstd, g2[j], b2[j]);
                    }
                }
            }
        }
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
    }
}

// ------------------------------------------------------------------------------------- host side
}  // namespace fs2k

using namespace fs2k;

extern "C" int fs2k_gemm_tc_supported(int K, int N, int lda, int taps) {
    if (K <= 0 || N <= 0 || taps < 1) return 0;
    if ((lda & 3) || (K & 3)) return 0;                     // TMA: 16-byte global strides
    if (N <= 256) return (N % 16) == 0;
    return (N % 256) == 0 || (N % 128) == 0;
}

extern "C" int fs2k_gemm_tc(const float* A, int lda, int B, int L, int K, const float* W, int N, int taps, int pad,
                            const float* bias, const float* scale, const float* shift, int act, float alpha,
                            const float* residual, int ldr, const uint8_t* row_mask, float* C, int ldc,
                            const float* ln_gamma, const float* ln_beta, float ln_eps, float* ln_out,
                            const float* ln2_gamma, const float* ln2_beta, float* ln2_out, float dropout_p, long seed,
                            int passes, const float* W_small, fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && L >= 0 && K > 0 && N > 0 && taps >= 1 && pad >= 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(act >= 0 && act <= 3 && (passes == 1 || passes == 3), FS2K_ERR_UNSUPPORTED);
    FS2K_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(fs2k_gemm_tc_supported(K, N, lda, taps), FS2K_ERR_UNSUPPORTED);
    const long M = (long)B * L;
    if (M == 0) return FS2K_OK;
    FS2K_REQUIRE(A && W && (C || ln_out), FS2K_ERR_NULL);
    FS2K_REQUIRE(!scale || shift, FS2K_ERR_NULL);
    FS2K_REQUIRE((ldc & 3) == 0 && (!residual || (ldr & 3) == 0), FS2K_ERR_UNSUPPORTED);
    int block_n = N <= 256 ? N : ((N % 256) == 0 ? 256 : 128);
    // (measured: narrowing 3×TF32 tiles to 128 columns for a 3-stage ring shortens the main loop 11 → 7.6 µs per tile
    // but doubles the number of ≈5 µs epilogues — slower overall on these 1-wave problems, so tiles stay 256 wide)
    if (!ln_out) {
        // small-M problems (encoder, predictors: ≤ 20 row tiles) are latency-bound per CTA: narrow the
        // column tile until the grid covers most of the 148 SMs
        const long tiles_m = (taps == 1) ? (M + TC_BM - 1) / TC_BM : (long)B * ((L + TC_BM - 1) / TC_BM);
        while (block_n > 32 && (block_n / 2) % 16 == 0 && N % (block_n / 2) == 0 && tiles_m * (N / block_n) < 100)
            block_n /= 2;
    }
    FS2K_REQUIRE(!ln_out || (block_n == N && ln_gamma && ln_beta), FS2K_ERR_UNSUPPORTED);
    FS2K_REQUIRE(!ln2_out || (ln_out && ln2_gamma && ln2_beta), FS2K_ERR_UNSUPPORTED);
    EncodeTiledFn encode = get_encode();
    FS2K_REQUIRE(encode != nullptr, FS2K_ERR_ARCH);

    // taps == 1: flatten to one "utterance" of M rows so tiles pack densely; otherwise tiles stay inside one b
    const int Bm = taps == 1 ? 1 : B;
    const long Lm = taps == 1 ? M : L;
    FS2K_REQUIRE(Lm < (1L << 31), FS2K_ERR_UNSUPPORTED);
    const int tiles_per_b = (int)((Lm + TC_BM - 1) / TC_BM);

    const int bk = passes == 3 ? TcBK<3>::value : TcBK<1>::value;
    const CUtensorMapSwizzle swz = bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUtensorMap tmA, tmB, tmBs;
    const int presplit_b = (passes == 3 && W_small != nullptr) ? 1 : 0;
    {
        cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)Lm, (cuuint64_t)Bm};
        cuuint64_t strides[2] = {(cuuint64_t)lda * 4, (cuuint64_t)Lm * lda * 4};
        cuuint32_t box[3] = {(cuuint32_t)bk, TC_BM, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)A, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fs2k_set_cuda_error(cudaErrorInvalidValue);
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)taps * N};
        cuuint64_t strides[1] = {(cuuint64_t)K * 4};
        cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)block_n};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)W, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fs2k_set_cuda_error(cudaErrorInvalidValue);
        r = encode(&tmBs, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)(presplit_b ? W_small : W), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fs2k_set_cuda_error(cudaErrorInvalidValue);
    }
    TcEpilogue ep{bias, scale, shift, act, alpha, residual, ldr, row_mask, C, ldc,
                  ln_gamma, ln_beta, ln_eps, ln_out, N, ln2_gamma, ln2_beta, ln2_out, dropout_p, (unsigned long long)seed};
    const size_t stage = (size_t)(TC_BM + block_n) * bk * 4 * (passes == 3 ? 2 : 1);
    int n_stages = (int)((226 * 1024) / stage);
    const int max_stages = passes == 3 ? TC_MAX_STAGES : 4;
    if (n_stages > max_stages) n_stages = max_stages;
    FS2K_REQUIRE(n_stages >= 2, FS2K_ERR_UNSUPPORTED);
    size_t smem = stage * n_stages;
    const size_t staging = (size_t)TC_BM * (block_n + 4) * 4;
    if (smem < staging) smem = staging;
    smem += 1024;  // manual 1024-byte alignment of the swizzled tiles
    FS2K_REQUIRE(smem <= 227 * 1024, FS2K_ERR_UNSUPPORTED);
    dim3 grid(Bm * tiles_per_b, N / block_n);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e;
    auto launch = [&](auto kernel) -> cudaError_t {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        fs2k_launch(kernel, dim3(grid), dim3(TC_THREADS), smem, s, tmA, tmB, tmBs, presplit_b, (int)Lm, M, K, N, block_n, taps, pad, tiles_per_b, n_stages, ep);
        return cudaSuccess;
    };
    const bool drop = dropout_p > 0.f;
    FS2K_REQUIRE(!(drop && ln_out), FS2K_ERR_UNSUPPORTED);
    if (ln_out) e = passes == 3 ? launch(gemm_tc_kernel<3, false, true>) : launch(gemm_tc_kernel<1, false, true>);
    else if (passes == 3) e = drop ? launch(gemm_tc_kernel<3, true, false>) : launch(gemm_tc_kernel<3, false, false>);
    else e = drop ? launch(gemm_tc_kernel<1, true, false>) : launch(gemm_tc_kernel<1, false, false>);
    if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

namespace fs2k {
__global__ void __launch_bounds__(256) split_small_kernel(const float* __restrict__ x, long n, float* __restrict__ out) {
    pdl_prologue();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float v = x[i];
        out[i] = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    }
}
}  // namespace fs2k

extern "C" int fs2k_split_small(const float* x, long n, float* out, fs2k_stream_t stream) {
    FS2K_REQUIRE(n >= 0, FS2K_ERR_BAD_SHAPE);
    if (n == 0) return FS2K_OK;
    FS2K_REQUIRE(x && out, FS2K_ERR_NULL);
    long g = (n + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    fs2k_launch(fs2k::split_small_kernel, dim3((unsigned)g), dim3(256), 0, (cudaStream_t)stream, x, n, out);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

FS2K_DEFINE_SEED_BASE_SETTER(gemm_tc)
