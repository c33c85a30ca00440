// Weight gradients of Linear / Conv1d in the bf16 arithmetic mode:
//
//   dW[tap][n][k] = Σ_{b,l} G[b,l,n] · X[b, l+tap−pad, k]
//
// The contraction runs over the ROWS of both operands, so both are MN-major for tcgen05.mma: a stage holds 64
// contraction rows of G (128 n) and X (tile_k k) as blocks of 64 columns × 64 rows × 128 B, 16-byte chunks swizzled
// with the row index (the canonical MN-major SWIZZLE_128B layout: LBO = 8192 B between 64-wide blocks, SBO = 1024 B
// between groups of 8 contraction rows).  Both operands are fp32 in HBM (gradients and saved activations of the
// fp32 residual stream): eight producer warps read them with coalesced 16-byte loads, round to bf16 in registers and
// lay the tiles down; zero rows stand in for the frames a convolution tap reaches outside its utterance.  Products
// accumulate in fp32 in TMEM.  The row range is split over CTAs; partial tiles go to a workspace that
// wgrad_reduce_kernel (gemm_wgrad_tc.cu) sums in a fixed order.  N and K need not be multiples of the tile: the
// 80-channel mel projections run here too (rows / columns past the edge are zero-filled and never written).
#include <cstring>

#include "bf16_common.cuh"

namespace fs2k {

constexpr int WB_N = 128;        // dW rows per tile = MMA M = TMEM lanes
constexpr int WB_ROWS = 64;      // contraction rows per stage (4 MMAs of K = 16)
constexpr int WB_THREADS = 384;
constexpr int WB_CONV_THREADS = 256;
constexpr int WB_STAGES = 4;

// G16 / X16: that operand is bf16 in HBM and comes in by TMA (boxes of 64 columns × 64 rows of a 3-D map [B][L][C],
// rows outside the utterance zero-filled) instead of through the converting producer warps.
template <bool G16, bool X16>
__global__ void __launch_bounds__(WB_THREADS, 1)
gemm_wgrad_bf16_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX,
                       const float* __restrict__ G, int ldg, const float* __restrict__ X, int ldx, int L, int N, int K,
                       int tile_k, int taps, int pad, int chunks_per_b, int n_chunks, int chunks_per_split,
                       float* __restrict__ ws, float* __restrict__ red_out) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_full[WB_STAGES], s_empty[WB_STAGES], s_tmem_full;
    __shared__ uint32_t s_tmem_base;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t a_bytes = (WB_N / 64) * 8192u;
    const uint32_t b_bytes = (uint32_t)((tile_k + 63) / 64) * 8192u;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * WB_N, k0 = blockIdx.y * tile_k;
    const int tap = blockIdx.z % taps, split = blockIdx.z / taps;
    const int c_begin = split * chunks_per_split, c_end = min(n_chunks, c_begin + chunks_per_split);
    const int iters = max(c_end - c_begin, 0);

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < WB_STAGES; ++s) {
            mbar_init(smem_u32(&s_full[s]), ((G16 && X16) ? 0 : WB_CONV_THREADS) + ((G16 || X16) ? 1 : 0));
            mbar_init(smem_u32(&s_empty[s]), 1);
        }
        mbar_init(smem_u32(&s_tmem_full), 1);
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
        if ((G16 || X16) && lane == 0) {
            // ================= TMA producer for the bf16 operand(s) =================
            for (int it = 0; it < iters; ++it) {
                const int s = it % WB_STAGES, ph = (it / WB_STAGES) & 1;
                mbar_wait(smem_u32(&s_empty[s]), ph ^ 1);
                const int c = c_begin + it;
                const int b = c / chunks_per_b, l0 = (c % chunks_per_b) * WB_ROWS;
                const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes), sb = sa + a_bytes;
                const uint32_t bar = smem_u32(&s_full[s]);
                mbar_expect_tx(bar, (G16 ? a_bytes : 0u) + (X16 ? b_bytes : 0u));
                if (G16)
                    for (int j = 0; j < WB_N / 64; ++j) tma_load_3d(sa + j * 8192, &tmG, bar, n0 + 64 * j, l0, b);
                if (X16)
                    for (int j = 0; j * 64 < tile_k; ++j) tma_load_3d(sb + j * 8192, &tmX, bar, k0 + 64 * j, l0 + tap - pad, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = hb_idesc(tile_k, 1, 1);  // both operands MN-major
            for (int it = 0; it < iters; ++it) {
                const int s = it % WB_STAGES, ph = (it / WB_STAGES) & 1;
                mbar_wait(smem_u32(&s_full[s]), ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes), sb = sa + a_bytes;
#pragma unroll
                for (int kk = 0; kk < WB_ROWS / 16; ++kk)
                    hb_mma(tmem_base, hb_desc_mn(sa + kk * 2048, 8192), hb_desc_mn(sb + kk * 2048, 8192), idesc, (it | kk) ? 1u : 0u);
                tc_commit(smem_u32(&s_empty[s]));
            }
            tc_commit(smem_u32(&s_tmem_full));
        }
    } else if (warp >= 4) {
        // ================= producers: fp32 rows → bf16 MN-major tiles =================
        const int ct = threadIdx.x - 128;                  // 0..255
        const int g_c4 = ct & 31, g_r = ct >> 5;           // G: 32 float4 per 128-column row, 8 rows per pass
        const int x_c4 = ct & 63, x_r = ct >> 6;           // X: up to 64 float4 per row, 4 rows per pass
        const bool g_col_ok = n0 + g_c4 * 4 < N;
        const bool x_col_ok = x_c4 * 4 < tile_k && k0 + x_c4 * 4 < K;
        for (int it = 0; it < ((G16 && X16) ? 0 : iters); ++it) {
            const int s = it % WB_STAGES, ph = (it / WB_STAGES) & 1;
            const int c = c_begin + it;
            const int b = c / chunks_per_b, l0 = (c % chunks_per_b) * WB_ROWS;
            float4 gv[8], xv[16];
            if (!G16) {
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const int l = l0 + p * 8 + g_r;
                    gv[p] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (g_col_ok && l < L) gv[p] = ld_stream(reinterpret_cast<const float4*>(G + ((size_t)b * L + l) * ldg + n0 + g_c4 * 4));
                }
            }
            if (!X16) {
#pragma unroll
                for (int p = 0; p < 16; ++p) {
                    const int l = l0 + p * 4 + x_r, lx = l + tap - pad;
                    xv[p] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (x_col_ok && l < L && lx >= 0 && lx < L)
                        xv[p] = ld_stream(reinterpret_cast<const float4*>(X + ((size_t)b * L + lx) * ldx + k0 + x_c4 * 4));
                }
            }
            mbar_wait(smem_u32(&s_empty[s]), ph ^ 1);
            uint8_t* sa = smem + (size_t)s * stage_bytes;
            uint8_t* sb = sa + a_bytes;
            if (!G16) {
#pragma unroll
                for (int p = 0; p < 8; ++p)
                    *reinterpret_cast<uint2*>(sa + hb_tile_off(p * 8 + g_r, g_c4 * 4, 8192)) = hb_pack4(gv[p]);
            }
            if (!X16 && x_c4 * 4 < tile_k) {
#pragma unroll
                for (int p = 0; p < 16; ++p)
                    *reinterpret_cast<uint2*>(sb + hb_tile_off(p * 4 + x_r, x_c4 * 4, 8192)) = hb_pack4(xv[p]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_full[s])) : "memory");
        }
        // ================= partial tile → workspace [split][tap][N][K] =================
        mbar_wait(smem_u32(&s_tmem_full), 0);
        tc_fence_after();
        const int q = warp & 3, half = (warp - 4) >> 2;
        const int n = n0 + q * 32 + lane;
        if (red_out) {
            // taps == 1: the partial tile is added straight into dW[N][K] with 16-byte reductions resolved in L2 — no workspace
            // round trip, no reduce launch (profiles/red_probe.cu: 12 µs vs 16–20 µs for the ≈ 19 MB of partial tiles of one launch)
            float* dst = red_out + (size_t)n * K + k0;
            if (iters > 0)
                for (int c0 = half * 16; c0 < tile_k; c0 += 32) {
                    float v[16];
                    tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                    if (n < N) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) red_add4(dst + c0 + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                }
        } else {
            float* dst = ws + (((size_t)split * taps + tap) * N + n) * K + k0;
            for (int c0 = half * 16; c0 < tile_k; c0 += 32) {
                float v[16];
                if (iters > 0) {
                    tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0.f;
                }
                if (n < N) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
    }
}

}  // namespace fs2k

using namespace fs2k;

static void wgrad_bf16_plan(int B, int L, int N, int K, int taps, int* tile_k, int* splits, int* chunks_per_split) {
    *tile_k = K >= 256 ? 256 : K;
    const int chunks_per_b = (L + WB_ROWS - 1) / WB_ROWS;
    const long n_chunks = (long)B * chunks_per_b;
    const long tiles = (long)((N + WB_N - 1) / WB_N) * (K / *tile_k) * taps;
    long s = 148 / tiles;  // one CTA per SM (192 KB of shared memory each) and ONE wave: rounding up (152 CTAs for 8 tiles) left a second
                           // wave of 4 CTAs that doubled the kernel's duration
    if (s > n_chunks) s = n_chunks;
    if (s < 1) s = 1;
    long cps = (n_chunks + s - 1) / s;
    if (cps < 1) cps = 1;
    s = (n_chunks + cps - 1) / cps;
    *splits = (int)(s < 1 ? 1 : s);
    *chunks_per_split = (int)cps;
}

extern "C" int fs2k_gemm_wgrad_bf16_supported(int N, int K, int ldg, int ldx) {
    if (N <= 0 || K <= 0) return 0;
    if ((N & 3) || (K & 15)) return 0;           // float4 rows of G; MMA N (= K tile) in steps of 16
    if (K > 256 && (K % 256)) return 0;
    return (ldg % 4) == 0 && (ldx % 4) == 0;
}

extern "C" size_t fs2k_gemm_wgrad_bf16_workspace_bytes(int B, int L, int N, int K, int taps) {
    int tile_k, splits, cps;
    wgrad_bf16_plan(B, L, N, K, taps, &tile_k, &splits, &cps);
    return (size_t)splits * taps * N * K * sizeof(float);
}

static bool wb_make_map(CUtensorMap* tm, const void* base, int cols, int ld, int L, int B) {
    EncodeTiledFn encode = get_encode();
    if (!encode) return false;
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)L, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)L * ld * 2};
    cuuint32_t box[3] = {64, WB_ROWS, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

extern "C" int fs2k_gemm_wgrad_bf16_ex(const void* G, int g_is_bf16, int ldg, const void* X, int x_is_bf16, int ldx, int B, int L,
                                       int N, int K, int taps, int pad, void* workspace, size_t workspace_bytes,
                                       float* dW_param_layout, int accumulate, fs2k_stream_t stream) {
    FS2K_REQUIRE(B > 0 && L > 0 && N > 0 && K > 0 && taps >= 1 && pad >= 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(fs2k_gemm_wgrad_bf16_supported(N, K, ldg, ldx), FS2K_ERR_UNSUPPORTED);
    FS2K_REQUIRE((!g_is_bf16 || (ldg % 8) == 0) && (!x_is_bf16 || (ldx % 8) == 0), FS2K_ERR_UNSUPPORTED);
    FS2K_REQUIRE(G && X && workspace && dW_param_layout, FS2K_ERR_NULL);
    FS2K_REQUIRE(workspace_bytes >= fs2k_gemm_wgrad_bf16_workspace_bytes(B, L, N, K, taps), FS2K_ERR_WORKSPACE);
    int tile_k, splits, cps;
    wgrad_bf16_plan(B, L, N, K, taps, &tile_k, &splits, &cps);
    const int chunks_per_b = (L + WB_ROWS - 1) / WB_ROWS;
    const int n_chunks = B * chunks_per_b;
    FS2K_REQUIRE((long)splits * taps <= 65535, FS2K_ERR_UNSUPPORTED);
    CUtensorMap tmG, tmX;
    memset(&tmG, 0, sizeof(tmG));
    memset(&tmX, 0, sizeof(tmX));
    if (g_is_bf16 && !wb_make_map(&tmG, G, N, ldg, L, B)) return fs2k_set_cuda_error(cudaErrorInvalidValue);
    if (x_is_bf16 && !wb_make_map(&tmX, X, K, ldx, L, B)) return fs2k_set_cuda_error(cudaErrorInvalidValue);
    const size_t stage = (size_t)(WB_N / 64) * 8192 + (size_t)((tile_k + 63) / 64) * 8192;
    const size_t smem = stage * WB_STAGES + 1024;
    FS2K_REQUIRE(smem <= 227 * 1024, FS2K_ERR_UNSUPPORTED);
    dim3 grid((N + WB_N - 1) / WB_N, K / tile_k, (unsigned)(splits * taps));
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
    // taps == 1 and more than one split: partial tiles are reduced in L2 (red.global.add.v4.f32) straight into dW
    float* red_out = (taps == 1 && splits > 1 && wgrad_atomic_enabled()) ? dW_param_layout : nullptr;
    if (red_out && !accumulate) {
        e = cudaMemsetAsync(dW_param_layout, 0, (size_t)N * K * sizeof(float), s);
        if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    }
    auto launch = [&](auto kernel) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return;
        fs2k_launch(kernel, dim3(grid), dim3(WB_THREADS), smem, s, tmG, tmX, g_is_bf16 ? nullptr : (const float*)G, ldg,
                    x_is_bf16 ? nullptr : (const float*)X, ldx, L, N, K, tile_k, taps, pad, chunks_per_b, n_chunks, cps, (float*)workspace, red_out);
    };
    if (g_is_bf16 && x_is_bf16) launch(gemm_wgrad_bf16_kernel<true, true>);
    else if (g_is_bf16) launch(gemm_wgrad_bf16_kernel<true, false>);
    else if (x_is_bf16) launch(gemm_wgrad_bf16_kernel<false, true>);
    else launch(gemm_wgrad_bf16_kernel<false, false>);
    if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    FS2K_CHECK_LAUNCH();
    if (!red_out) {
        wgrad_reduce_launch((const float*)workspace, splits, taps, N, K, accumulate, dW_param_layout, s);
        FS2K_CHECK_LAUNCH();
    }
    return FS2K_OK;
}

extern "C" int fs2k_gemm_wgrad_bf16(const float* G, int ldg, const float* X, int ldx, int B, int L, int N, int K, int taps,
                                    int pad, void* workspace, size_t workspace_bytes, float* dW_param_layout, int accumulate,
                                    fs2k_stream_t stream) {
    return fs2k_gemm_wgrad_bf16_ex(G, 0, ldg, X, 0, ldx, B, L, N, K, taps, pad, workspace, workspace_bytes, dW_param_layout, accumulate, stream);
}
