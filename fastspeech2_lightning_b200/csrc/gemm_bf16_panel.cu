// Row-panel variant of the bf16 GEMM for the model's dominant shape class: K ≤ 256 (the Conformer width), fp32
// activations in HBM, N up to 1024.  These contractions are HBM-bound (≈ 40 FLOP/B), so the kernel is organised
// around the memory streams, not the MMA:
//
//   * one CTA owns a 128-row panel of A.  The panel is read from HBM exactly once (256 threads, coalesced 16-byte
//     loads, all of them in flight together), rounded to bf16 and parked in shared memory as the K-major operand
//     (64 KB for K = 256) for the whole life of the CTA;
//   * the CTA then walks over the N tiles of 128 columns: a TMA ring streams the bf16 weight tiles (L2-resident:
//     ≤ 512 KB per layer), one elected thread issues the tcgen05.mma's into one of TWO TMEM accumulators;
//   * eight epilogue warps drain the other accumulator meanwhile: tcgen05.ld → private 32 × 64 staging tile →
//     bias / activation / dropout / residual / row mask → coalesced 128-byte-line stores (fp32 and / or bf16 final
//     value, fp32 and / or bf16 pre-activation).  The accumulator is handed back as soon as it is in registers, so the
//     tensor pipe never waits for the stores.
// Traffic per launch = the algorithmic minimum (A once, outputs once, residual once); weights come from L2.
// w_mn = 1 reads the forward layer's [K][N] weight array as an MN-major operand (data-gradient GEMM).
// Shapes outside this class (K > 256, convolution taps, bf16 A, 3-term split, folded BatchNorm) use gemm_bf16.cu.
#include <cstring>

#include "bf16_common.cuh"

namespace fs2k {

constexpr int PB_BN = 128;          // columns per N tile == TMEM columns per accumulator
constexpr int PB_BK = 64;
constexpr int PB_THREADS = 384;     // warp 0 TMA, 1 MMA, 2 TMEM owner, 3 idle, 4..11 A converters then epilogue
constexpr int PB_STAGES = 4;
constexpr int PB_EPI_WARPS = 8;
constexpr int PB_STAG_PITCH = 68;   // floats per staging row (64 + 4: conflict-free 16-byte accesses)

struct PbEpilogue {
    const float* bias; int act; float alpha;
    const float* residual; int ldr;
    const uint8_t* row_mask;
    float* C; int ldc;
    __nv_bfloat16* C16; int ldc16;
    float* P32; __nv_bfloat16* P16; int ldp;
    float drop_p; unsigned long long seed;
    const __nv_bfloat16* dact_pre16;   // backward fusion: value *= act'(pre) with pre read here (stride ldp) — then dropout mask
};

__device__ __forceinline__ float pb_act_grad(float u, int act) {
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

__device__ __forceinline__ float pb_act(float v, int act) {
    if (act == FS2K_ACT_RELU) return fmaxf(v, 0.f);
    if (act == FS2K_ACT_SILU) return silu(v);
    if (act == FS2K_ACT_TANH) return tanhf(v);
    return v;
}

// epilogue flavours (template): the store loop is instruction-bound (2 epilogue warps per scheduler), so the generic
// runtime-`act` path — three compares and an IEEE division per element — is kept out of the common launches
enum { PB_EPI_PLAIN = 0, PB_EPI_SILU = 1, PB_EPI_GENERIC = 2, PB_EPI_DACT = 3 };
// SiLU through the SFU: ex2.approx + rcp.approx (≈ 2^-21 relative — two orders below the bf16 rounding of the operands)
__device__ __forceinline__ float pb_silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ float pb_silu_grad_fast(float u) {
    const float sg = __fdividef(1.0f, 1.0f + __expf(-u));
    return sg * fmaf(u, 1.0f - sg, 1.0f);
}

template <bool B_MN, bool DROPOUT, int EPI>
__global__ void __launch_bounds__(PB_THREADS, 1)
gemm_bf16_panel_kernel(const __grid_constant__ CUtensorMap tmB, const float* __restrict__ A32, int lda, long M, int K, int N,
                       int tiles_per_cta, PbEpilogue ep) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_full[PB_STAGES], s_empty[PB_STAGES], s_a_full, s_tmem_full[2], s_tmem_empty[2];
    __shared__ uint32_t s_tmem_base;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                                   // [K/64] blocks of 128 rows × 128 B
    uint8_t* sB = smem + 4 * 16384;                       // ring of 16 KB weight tiles
    float* sStag = reinterpret_cast<float*>(smem + 4 * 16384 + PB_STAGES * 16384);  // 8 warps × 32 rows × 68 floats
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long m0 = (long)blockIdx.x * HB_BM;
    const int nk = K / PB_BK;
    const int n_tiles_total = N / PB_BN;
    const int j_begin = blockIdx.y * tiles_per_cta, j_end = min(n_tiles_total, j_begin + tiles_per_cta);
    const int n_tiles = j_end - j_begin;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < PB_STAGES; ++s) {
            mbar_init(smem_u32(&s_full[s]), 1);
            mbar_init(smem_u32(&s_empty[s]), 1);
        }
        mbar_init(smem_u32(&s_a_full), PB_EPI_WARPS * 32);
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&s_tmem_full[i]), 1);
            mbar_init(smem_u32(&s_tmem_empty[i]), PB_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&s_tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            // ================= TMA producer: weight tiles =================
            int it = 0;
            for (int j = j_begin; j < j_end; ++j) {
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = it % PB_STAGES, ph = (it / PB_STAGES) & 1;
                    mbar_wait(smem_u32(&s_empty[s]), ph ^ 1);
                    const uint32_t dst = smem_u32(sB + (size_t)s * 16384);
                    const uint32_t bar = smem_u32(&s_full[s]);
                    mbar_expect_tx(bar, 16384);
                    if (B_MN) {
                        tma_load_2d(dst, &tmB, bar, j * PB_BN, kb * PB_BK);
                        tma_load_2d(dst + 8192, &tmB, bar, j * PB_BN + 64, kb * PB_BK);
                    } else {
                        tma_load_2d(dst, &tmB, bar, kb * PB_BK, j * PB_BN);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ================= MMA issuer =================
            const uint32_t idesc = hb_idesc(PB_BN, 0, B_MN ? 1 : 0);
            mbar_wait(smem_u32(&s_a_full), 0);
            tc_fence_after();
            int it = 0;
            for (int jj = 0; jj < n_tiles; ++jj) {
                const int buf = jj & 1;
                mbar_wait(smem_u32(&s_tmem_empty[buf]), ((jj >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)buf * PB_BN;
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = it % PB_STAGES, ph = (it / PB_STAGES) & 1;
                    mbar_wait(smem_u32(&s_full[s]), ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(sA + (size_t)kb * 16384), sb = smem_u32(sB + (size_t)s * 16384);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t bd = B_MN ? hb_desc_mn(sb + k * 2048, 8192) : hb_desc_k(sb + k * 32);
                        hb_mma(acc, hb_desc_k(sa + k * 32), bd, idesc, (kb | k) ? 1u : 0u);
                    }
                    tc_commit(smem_u32(&s_empty[s]));
                }
                tc_commit(smem_u32(&s_tmem_full[buf]));
            }
        }
    } else if (warp >= 4) {
        // ================= A panel: fp32 HBM → bf16 K-major tile, once =================
        {
            const int ct = threadIdx.x - 128;            // 0..255
            const int c4 = ct & 63, r_in = ct >> 6;      // 64 float4 per 256-float row; 4 rows per pass
            const bool col_ok = c4 * 4 < K;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                float4 v[16];
#pragma unroll
                for (int p = 0; p < 16; ++p) {
                    const long m = m0 + (half * 16 + p) * 4 + r_in;
                    v[p] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (col_ok && m < M) v[p] = ld_stream(reinterpret_cast<const float4*>(A32 + (size_t)m * lda + c4 * 4));
                }
                if (col_ok) {
#pragma unroll
                    for (int p = 0; p < 16; ++p)
                        *reinterpret_cast<uint2*>(sA + hb_tile_off((half * 16 + p) * 4 + r_in, c4 * 4, 16384)) = hb_pack4(v[p]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_a_full)) : "memory");
        }
        // ================= epilogue: one accumulator at a time, the MMA warp fills the other =================
        const int ew = warp - 4;
        const int quad = warp & 3, chalf = ew >> 2;       // TMEM lane quadrant, 64-column half of the tile
        float* stag = sStag + (size_t)ew * 32 * PB_STAG_PITCH;
        const int c4 = lane & 15, rsub = lane >> 4;       // store pass: 16 lanes cover a 64-float row, 2 rows per instruction
        const float drop_p = DROPOUT ? ep.drop_p : 0.f;
        const uint32_t thr16 = drop_thr16(drop_p);
        const float inv_keep = drop_inv_keep(thr16);
        const unsigned long long seed = DROPOUT ? seed_with_base(ep.seed) : 0ull;
        const int act = ep.act;
        const float alpha = ep.alpha;
        // everything the store loop needs, in registers: base pointers of this lane's first row (rows advance by 2 per step)
        const long mrow0 = m0 + quad * 32 + rsub;
        const int rows_left = (int)min((long)32, M - (m0 + quad * 32));   // rows of this warp that exist
        const float* p_res = ep.residual;
        const uint8_t* p_mask = ep.row_mask;
        float* p_C = ep.C;
        __nv_bfloat16* p_C16 = ep.C16;
        float* p_P32 = ep.P32;
        __nv_bfloat16* p_P16 = ep.P16;
        const __nv_bfloat16* p_pre = ep.dact_pre16;
        const int ldr = ep.ldr, ldc = ep.ldc, ldc16 = ep.ldc16, ldp = ep.ldp;
        for (int jj = 0; jj < n_tiles; ++jj) {
            const int buf = jj & 1;
            const int ncol = (j_begin + jj) * PB_BN + chalf * 64 + c4 * 4;   // this lane's 4 output columns
            const float4 bias4 = ep.bias ? __ldg(reinterpret_cast<const float4*>(ep.bias + ncol)) : make_float4(0.f, 0.f, 0.f, 0.f);
            // software-pipelined global reads: trip t+1's residual rows / row masks are in flight while trip t is processed,
            // and trip 0's are issued before the wait for the accumulator
            float4 res_n[4];
            float rm_n[4];
            // DACT: the saved bf16 pre-activations of ALL 32 rows of this lane's columns are requested before the wait for the
            // accumulator (16 loads in flight behind the MMAs of the block).  Read inline they exposed one global-load latency per
            // trip (119 µs per decoder FFN dgrad against 55 µs for the forward of the same shape); one trip ahead still 104 µs.
            constexpr bool PIPE_PRE = EPI == PB_EPI_DACT;
            uint2 pre_all[PIPE_PRE ? 16 : 1];
            if (PIPE_PRE) {
#pragma unroll
                for (int t = 0; t < 4; ++t)
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = t * 8 + u * 2 + rsub;
                        pre_all[t * 4 + u] = r < rows_left ? *reinterpret_cast<const uint2*>(p_pre + (size_t)(mrow0 + t * 8 + u * 2) * ldp + ncol)
                                                           : make_uint2(0u, 0u);
                    }
            }
            auto load_trip = [&](int rb) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = rb + u * 2 + rsub;
                    res_n[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    rm_n[u] = 1.f;
                    if (r < rows_left) {
                        if (p_res) res_n[u] = *reinterpret_cast<const float4*>(p_res + (size_t)(mrow0 + rb + u * 2) * ldr + ncol);
                        if (p_mask) rm_n[u] = p_mask[mrow0 + rb + u * 2] ? 1.f : 0.f;
                    }
                }
            };
            if (p_res || p_mask) load_trip(0);
            mbar_wait(smem_u32(&s_tmem_full[buf]), (jj >> 1) & 1);
            tc_fence_after();
            {
                float v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * PB_BN + chalf * 64);
                tc_ld32(taddr, v);
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4*>(stag + lane * PB_STAG_PITCH + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                tc_ld32(taddr + 32, v);
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4*>(stag + lane * PB_STAG_PITCH + 32 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_tmem_empty[buf])) : "memory");
            // rows of this warp: quad*32 + [0,32); 4 row pairs per trip
#pragma unroll 1
            for (int rb = 0; rb < 32; rb += 8) {
                float4 res[4];
                float rm[4];
                uint2 pre[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    res[u] = (p_res || p_mask) ? res_n[u] : make_float4(0.f, 0.f, 0.f, 0.f);
                    rm[u] = (p_res || p_mask) ? rm_n[u] : 1.f;
                    pre[u] = PIPE_PRE ? pre_all[u] : make_uint2(0u, 0u);
                }
                if (PIPE_PRE) {  // next trip's values move to the front (the trip loop is not unrolled: static register indices)
#pragma unroll
                    for (int i = 0; i < 12; ++i) pre_all[i] = pre_all[i + 4];
                }
                if ((p_res || p_mask) && rb + 8 < 32) load_trip(rb + 8);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = rb + u * 2 + rsub;
                    if (r >= rows_left) continue;
                    const size_t m = (size_t)(mrow0 + rb + u * 2);
                    float4 v = *reinterpret_cast<const float4*>(stag + r * PB_STAG_PITCH + c4 * 4);
                    v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w;
                    if (p_P32) *reinterpret_cast<float4*>(p_P32 + m * ldp + ncol) = v;
                    if (p_P16) *reinterpret_cast<uint2*>(p_P16 + m * ldp + ncol) = hb_pack4(v);
                    if (EPI == PB_EPI_DACT) {  // data-gradient GEMM fused with the activation's derivative (pre-activation saved in bf16)
                        const uint2 pk = pre[u];
                        const float2 p0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.x));
                        const float2 p1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.y));
                        if (act == FS2K_ACT_SILU) {
                            v.x *= pb_silu_grad_fast(p0.x); v.y *= pb_silu_grad_fast(p0.y);
                            v.z *= pb_silu_grad_fast(p1.x); v.w *= pb_silu_grad_fast(p1.y);
                        } else {
                            v.x *= pb_act_grad(p0.x, act); v.y *= pb_act_grad(p0.y, act);
                            v.z *= pb_act_grad(p1.x, act); v.w *= pb_act_grad(p1.y, act);
                        }
                    } else if (EPI == PB_EPI_SILU) {
                        v.x = pb_silu_fast(v.x); v.y = pb_silu_fast(v.y); v.z = pb_silu_fast(v.z); v.w = pb_silu_fast(v.w);
                    } else if (EPI == PB_EPI_GENERIC) {
                        v.x = pb_act(v.x, act); v.y = pb_act(v.y, act); v.z = pb_act(v.z, act); v.w = pb_act(v.w, act);
                    }
                    if (alpha != 1.0f) { v.x *= alpha; v.y *= alpha; v.z *= alpha; v.w *= alpha; }
                    if (DROPOUT) drop_apply4(v, seed, (unsigned long long)m * N + ncol, thr16, inv_keep);
                    if (p_res) { v.x += res[u].x; v.y += res[u].y; v.z += res[u].z; v.w += res[u].w; }
                    if (p_mask) { v.x *= rm[u]; v.y *= rm[u]; v.z *= rm[u]; v.w *= rm[u]; }
                    if (p_C) *reinterpret_cast<float4*>(p_C + m * ldc + ncol) = v;
                    if (p_C16) *reinterpret_cast<uint2*>(p_C16 + m * ldc16 + ncol) = hb_pack4(v);
                }
            }
            __syncwarp();  // the staging tile is rewritten by the next accumulator
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
    }
}

}  // namespace fs2k

using namespace fs2k;

// shapes the row-panel kernel takes (the caller — fs2k_gemm_bf16 — has already validated the generic contract)
bool fs2k_gemm_bf16_panel_ok(int K, int N, int taps, int a_is_bf16, bool has_lo, bool has_scale) {
    return !a_is_bf16 && !has_lo && !has_scale && taps == 1 && K % PB_BK == 0 && K <= 256 && N % PB_BN == 0;
}

int fs2k_gemm_bf16_panel_launch(const float* A, int lda, long M, int K, const void* W, int w_mn, int N, const float* bias, int act,
                                float alpha, const float* residual, int ldr, const uint8_t* row_mask, float* C, int ldc, void* C16,
                                int ldc16, float* P32, void* P16, int ldp, float dropout_p, long seed, const void* dact_pre16,
                                cudaStream_t s) {
    EncodeTiledFn encode = get_encode();
    FS2K_REQUIRE(encode != nullptr, FS2K_ERR_ARCH);
    CUtensorMap tmB;
    {
        // w_mn: W is [K][N] (row = contraction index): boxes of 64 columns × 64 rows; else [N][K]: 64 k × 128 rows
        cuuint64_t dims[2] = {(cuuint64_t)(w_mn ? N : K), (cuuint64_t)(w_mn ? K : N)};
        cuuint64_t strides[1] = {(cuuint64_t)(w_mn ? N : K) * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)(w_mn ? PB_BK : PB_BN)};
        cuuint32_t estr[2] = {1, 1};
        if (encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)W, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return fs2k_set_cuda_error(cudaErrorInvalidValue);
    }
    const int panels = (int)((M + HB_BM - 1) / HB_BM);
    const int n_tiles = N / PB_BN;
    int groups = 148 / panels;           // split the N tiles over CTAs only while that does not add a wave
    if (groups < 1) groups = 1;
    if (groups > n_tiles) groups = n_tiles;
    const int tiles_per_cta = (n_tiles + groups - 1) / groups;
    groups = (n_tiles + tiles_per_cta - 1) / tiles_per_cta;
    PbEpilogue ep{bias, act, alpha, residual, ldr, row_mask, C, ldc, (__nv_bfloat16*)C16, ldc16, P32, (__nv_bfloat16*)P16, ldp,
                  dropout_p, (unsigned long long)seed, (const __nv_bfloat16*)dact_pre16};
    const size_t smem = 4 * 16384 + PB_STAGES * 16384 + (size_t)PB_EPI_WARPS * 32 * PB_STAG_PITCH * 4 + 1024;
    dim3 grid(panels, groups);
    cudaError_t e = cudaSuccess;
    auto launch = [&](auto kernel) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return;
        fs2k_launch(kernel, dim3(grid), dim3(PB_THREADS), smem, s, tmB, A, lda, M, K, N, tiles_per_cta, ep);
    };
    const bool drop = dropout_p > 0.f;
    const int epi = dact_pre16 ? PB_EPI_DACT : act == FS2K_ACT_NONE ? PB_EPI_PLAIN : act == FS2K_ACT_SILU ? PB_EPI_SILU : PB_EPI_GENERIC;
#define PB_CASE(MN, DR, EP) if ((w_mn != 0) == MN && drop == DR && epi == EP) launch(gemm_bf16_panel_kernel<MN, DR, EP>);
#define PB_CASES(EP) PB_CASE(false, false, EP) PB_CASE(false, true, EP) PB_CASE(true, false, EP) PB_CASE(true, true, EP)
    PB_CASES(PB_EPI_PLAIN) PB_CASES(PB_EPI_SILU) PB_CASES(PB_EPI_GENERIC) PB_CASES(PB_EPI_DACT)
#undef PB_CASES
#undef PB_CASE
    if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

FS2K_DEFINE_SEED_BASE_SETTER(gemm_bf16_panel)
