// Weight gradients of Linear / Conv1d on the tensor cores:
//
//   dW[tap][n][k] = Σ_{b,l} G[b,l,n] · X[b, l+tap−pad, k]
//
// i.e. a GEMM whose contraction runs over the ROWS of both operands, so both are "MN-major" for
// tcgen05.mma: the TMA boxes are [32 rows (contraction)] × [32 floats (128-byte span along n or k)], laid down
// as 4 (G: 128 n) and TILE_K/32 (X) consecutive 4 KB blocks — the canonical MN-major SWIZZLE_128B_BASE32B
// layout (LBO = 4096 B between 32-wide atoms, SBO = 512 B between 4-row groups).  The time shift of the
// convolution taps is again a shifted box of a 3-D map [B][L][K] with zero-filled out-of-range rows.
// The row range is split over CTAs; each split writes its 128 × TILE_K partial tile to a workspace and a
// second kernel sums the splits in a fixed order (deterministic gradients, no atomics) while converting to
// the parameter's layout.  `passes = 3` = 3×TF32 split accumulation, as in gemm_tc.cu.
#include "tc_common.cuh"

namespace fs2k {

constexpr int WG_N = 128;        // dW rows per tile = MMA M = TMEM lanes
constexpr int WG_ROWS = 32;      // contraction rows per stage (4 MMAs of K = 8)
constexpr int WG_THREADS = 384;
constexpr int WG_EPI_WARPS = 8;
template <int PASSES> struct WgStages { static constexpr int value = PASSES == 3 ? 2 : 4; };

// MN-major tf32 operand: the only legal shared-memory layout is SWIZZLE_128B_BASE32B (32-byte swizzle chunks,
// atoms of 32 MN-elements × 4 contraction rows = 512 B); TMA writes it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
// 32-wide atoms along MN are 4096 B apart (one TMA box), groups of 4 contraction rows 512 B apart.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(4096 >> 4) << 16;   // leading byte offset: next 32-wide atom along MN
    d |= (uint64_t)(512 >> 4) << 32;    // stride byte offset: next group of 4 contraction rows
    d |= (uint64_t)1 << 46;             // descriptor version (Blackwell)
    d |= (uint64_t)1 << 61;             // SWIZZLE_128B_BASE32B
    return d;
}

template <int PASSES>
__global__ void __launch_bounds__(WG_THREADS, 1)
gemm_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX, int L, int N,
                     int K, int tile_k, int taps, int pad, int chunks_per_b, int n_chunks, int chunks_per_split,
                     float* __restrict__ ws, float* __restrict__ red_out) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    constexpr int STAGES = WgStages<PASSES>::value;
    __shared__ __align__(8) uint64_t s_full[STAGES], s_empty[STAGES], s_split[STAGES], s_tmem_full;
    __shared__ uint32_t s_tmem_base;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t a_bytes = WG_N * WG_ROWS * 4, b_bytes = tile_k * WG_ROWS * 4;
    const uint32_t stage_bytes = (a_bytes + b_bytes) * (PASSES == 3 ? 2 : 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * WG_N, k0 = blockIdx.y * tile_k;
    const int tap = blockIdx.z % taps, split = blockIdx.z / taps;
    const int c_begin = split * chunks_per_split, c_end = min(n_chunks, c_begin + chunks_per_split);
    const int iters = max(c_end - c_begin, 0);

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&s_full[s]), 1);
            mbar_init(smem_u32(&s_empty[s]), 1);
            mbar_init(smem_u32(&s_split[s]), WG_EPI_WARPS * 32);
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
    pdl_wait();  // setup above overlapped the previous kernel; global memory is touched from here on

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(smem_u32(&s_empty[s]), ph ^ 1);
                const int c = c_begin + it;
                const int b = c / chunks_per_b, l0 = (c % chunks_per_b) * WG_ROWS;
                const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes), sb = sa + a_bytes;
                const uint32_t bar = smem_u32(&s_full[s]);
                mbar_expect_tx(bar, a_bytes + b_bytes);
                for (int j = 0; j < WG_N / 32; ++j) tma_load_3d(sa + j * 4096, &tmG, bar, n0 + 32 * j, l0, b);
                for (int j = 0; j < tile_k / 32; ++j) tma_load_3d(sb + j * 4096, &tmX, bar, k0 + 32 * j, l0 + tap - pad, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // D = f32, A = B = tf32, both MN-major (bits 15, 16), M = 128, N = tile_k
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(tile_k >> 3) << 17) | ((uint32_t)(WG_N >> 4) << 24);
            for (int it = 0; it < iters; ++it) {
                const int s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(smem_u32(PASSES == 3 ? &s_split[s] : &s_full[s]), ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes), sb = sa + a_bytes;
#pragma unroll
                for (int kk = 0; kk < WG_ROWS / 8; ++kk) {
                    const uint64_t ad = make_smem_desc_mn(sa + kk * 1024), bd = make_smem_desc_mn(sb + kk * 1024);
                    tc_mma_tf32(tmem_base, ad, bd, idesc, (it | kk) ? 1u : 0u);
                    if (PASSES == 3) {
                        const uint64_t ads = make_smem_desc_mn(sa + a_bytes + b_bytes + kk * 1024);
                        const uint64_t bds = make_smem_desc_mn(sb + a_bytes + b_bytes + kk * 1024);
                        tc_mma_tf32(tmem_base, ad, bds, idesc, 1u);
                        tc_mma_tf32(tmem_base, ads, bd, idesc, 1u);
                    }
                }
                tc_commit(smem_u32(&s_empty[s]));
            }
            tc_commit(smem_u32(&s_tmem_full));
        }
    } else if (warp >= 4) {
        const int et = threadIdx.x - 128;
        if (PASSES == 3) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(smem_u32(&s_full[s]), ph);
                const float4* src = reinterpret_cast<const float4*>(smem + (size_t)s * stage_bytes);
                float4* dst = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes + a_bytes + b_bytes);
                split_small(src, dst, (a_bytes + b_bytes) / 16, et, WG_EPI_WARPS * 32);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_split[s])) : "memory");
            }
        }
        // partial tile → workspace [split][tap][N][K]
        mbar_wait(smem_u32(&s_tmem_full), 0);
        tc_fence_after();
        const int q = warp & 3, half = (warp - 4) >> 2;
        const int n = n0 + q * 32 + lane;
        if (red_out) {  // taps == 1: partial tile added straight into dW[N][K] by 16-byte reductions in L2 (see gemm_wgrad_bf16.cu)
            float* dst = red_out + (size_t)n * K + k0;
            if (iters > 0)
                for (int c0 = half * 16; c0 < tile_k; c0 += 32) {
                    float v[16];
                    tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
                    for (int j = 0; j < 16; j += 4) red_add4(dst + c0 + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
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
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
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

// out (parameter layout: [N][K][taps] or [N][K]) (+)= Σ_split ws[split][tap][n][k], fixed summation order.
// CTA = 32 float4 lanes × 8 split lanes: every thread sums its share of the splits (independent 16-byte loads), the
// eight partial sums are added in lane order through shared memory.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ ws, int splits, int taps, int N, int K, int accumulate,
                    float* __restrict__ out) {
    pdl_prologue();
    __shared__ float4 s_part[8][32];
    const long nk4 = ((long)N * K) >> 2, total4 = nk4 * taps, plane4 = total4;
    const int lane = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const long i4 = (long)blockIdx.x * 32 + lane;  // float4 index inside one split plane [tap][n][k]
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i4 < total4) {
        const float4* p = reinterpret_cast<const float4*>(ws) + i4;
#pragma unroll 4
        for (int sp = sl; sp < splits; sp += 8) {
            const float4 v = p[(size_t)sp * plane4];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    s_part[sl][lane] = acc;
    __syncthreads();
    if (sl == 0 && i4 < total4) {
#pragma unroll
        for (int g = 1; g < 8; ++g) {
            const float4 v = s_part[g][lane];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        const int t = (int)(i4 / nk4);
        const long nk = (i4 - (long)t * nk4) << 2;  // n·K + k of the first of the four elements
        if (taps == 1) {
            float4* o = reinterpret_cast<float4*>(out) + i4;
            if (accumulate) { const float4 c = *o; acc.x += c.x; acc.y += c.y; acc.z += c.z; acc.w += c.w; }
            *o = acc;
        } else {
            const float a[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float* o = out + (nk + e) * taps + t;
                *o = accumulate ? *o + a[e] : a[e];
            }
        }
    }
}

}  // namespace fs2k

using namespace fs2k;

static int g_wgrad_atomic = 1;
// 1 (default): single-tap weight gradients fold their split-K partial tiles into dW with red.global.add.v4.f32 (summation order
// not fixed: run-to-run differences in the last bits); 0: partial tiles to the workspace + deterministic reduce kernel
extern "C" int fs2k_wgrad_set_atomic(int enabled) {
    g_wgrad_atomic = enabled ? 1 : 0;
    return FS2K_OK;
}
bool fs2k::wgrad_atomic_enabled() { return g_wgrad_atomic != 0; }

// shared with gemm_wgrad_bf16.cu (declared in tc_common.cuh)
int fs2k::wgrad_reduce_launch(const float* ws, int splits, int taps, int N, int K, int accumulate, float* out, cudaStream_t s) {
    const long total4 = ((long)N * K * taps) >> 2;
    fs2k_launch(wgrad_reduce_kernel, dim3((int)((total4 + 31) / 32)), dim3(256), 0, s, ws, splits, taps, N, K, accumulate, out);
    return 0;
}

static void wgrad_tc_plan(int B, int L, int N, int K, int taps, int* tile_k, int* splits, int* chunks_per_split) {
    *tile_k = K >= 256 ? 256 : K;
    const int chunks_per_b = (L + WG_ROWS - 1) / WG_ROWS;
    const long n_chunks = (long)B * chunks_per_b;
    const long tiles = (long)(N / WG_N) * (K / *tile_k) * taps;
    long s = 148 / tiles;  // one CTA per SM (196 KB of shared memory), one wave; a second wave only doubles the partial tiles to reduce
    if (s > n_chunks) s = n_chunks;
    if (s < 1) s = 1;
    long cps = (n_chunks + s - 1) / s;
    if (cps < 1) cps = 1;
    s = (n_chunks + cps - 1) / cps;
    *splits = (int)(s < 1 ? 1 : s);
    *chunks_per_split = (int)cps;
}

extern "C" int fs2k_gemm_wgrad_tc_supported(int N, int K, int ldg, int ldx) {
    if (N <= 0 || K <= 0) return 0;
    if (N % WG_N) return 0;
    if (K >= 256 ? (K % 256) : (K % 32)) return 0;
    return (ldg % 4) == 0 && (ldx % 4) == 0;
}

extern "C" size_t fs2k_gemm_wgrad_tc_workspace_bytes(int B, int L, int N, int K, int taps) {
    int tile_k, splits, cps;
    wgrad_tc_plan(B, L, N, K, taps, &tile_k, &splits, &cps);
    return (size_t)splits * taps * N * K * sizeof(float);
}

extern "C" int fs2k_gemm_wgrad_tc(const float* G, int ldg, const float* X, int ldx, int B, int L, int N, int K, int taps,
                                  int pad, int passes, void* workspace, size_t workspace_bytes, float* dW_param_layout,
                                  int accumulate, fs2k_stream_t stream) {
    FS2K_REQUIRE(B > 0 && L > 0 && N > 0 && K > 0 && taps >= 1 && pad >= 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(passes == 1 || passes == 3, FS2K_ERR_UNSUPPORTED);
    FS2K_REQUIRE(fs2k_gemm_wgrad_tc_supported(N, K, ldg, ldx), FS2K_ERR_UNSUPPORTED);
    FS2K_REQUIRE(G && X && workspace && dW_param_layout, FS2K_ERR_NULL);
    FS2K_REQUIRE(workspace_bytes >= fs2k_gemm_wgrad_tc_workspace_bytes(B, L, N, K, taps), FS2K_ERR_WORKSPACE);
    EncodeTiledFn encode = get_encode();
    FS2K_REQUIRE(encode != nullptr, FS2K_ERR_ARCH);
    int tile_k, splits, cps;
    wgrad_tc_plan(B, L, N, K, taps, &tile_k, &splits, &cps);
    const int chunks_per_b = (L + WG_ROWS - 1) / WG_ROWS;
    const int n_chunks = B * chunks_per_b;
    FS2K_REQUIRE((long)splits * taps <= 65535, FS2K_ERR_UNSUPPORTED);
    CUtensorMap tmG, tmX;
    auto make_map = [&](CUtensorMap* tm, const float* base, int cols, int ld) -> bool {
        cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)L, (cuuint64_t)B};
        cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)L * ld * 4};
        cuuint32_t box[3] = {32, WG_ROWS, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    if (!make_map(&tmG, G, N, ldg) || !make_map(&tmX, X, K, ldx)) return fs2k_set_cuda_error(cudaErrorInvalidValue);
    const size_t stage = (size_t)(WG_N + tile_k) * WG_ROWS * 4 * (passes == 3 ? 2 : 1);
    const size_t smem = stage * (passes == 3 ? WgStages<3>::value : WgStages<1>::value) + 1024;
    FS2K_REQUIRE(smem <= 227 * 1024, FS2K_ERR_UNSUPPORTED);
    dim3 grid(N / WG_N, K / tile_k, (unsigned)(splits * taps));
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e;
    float* red_out = (taps == 1 && splits > 1 && wgrad_atomic_enabled()) ? dW_param_layout : nullptr;
    if (red_out && !accumulate) {
        e = cudaMemsetAsync(dW_param_layout, 0, (size_t)N * K * sizeof(float), s);
        if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    }
    if (passes == 3) {
        e = cudaFuncSetAttribute(gemm_wgrad_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fs2k_set_cuda_error(e);
        fs2k_launch(gemm_wgrad_tc_kernel<3>, dim3(grid), dim3(WG_THREADS), smem, s, tmG, tmX, L, N, K, tile_k, taps, pad, chunks_per_b, n_chunks, cps, (float*)workspace, red_out);
    } else {
        e = cudaFuncSetAttribute(gemm_wgrad_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fs2k_set_cuda_error(e);
        fs2k_launch(gemm_wgrad_tc_kernel<1>, dim3(grid), dim3(WG_THREADS), smem, s, tmG, tmX, L, N, K, tile_k, taps, pad, chunks_per_b, n_chunks, cps, (float*)workspace, red_out);
    }
    FS2K_CHECK_LAUNCH();
    if (!red_out) {
        const long total4 = ((long)N * K * taps) >> 2;  // N % 128 == 0, so N·K is a multiple of 4
        fs2k_launch(wgrad_reduce_kernel, dim3((int)((total4 + 31) / 32)), dim3(256), 0, s, (const float*)workspace, splits, taps, N, K, accumulate,
                                                                     dW_param_layout);
        FS2K_CHECK_LAUNCH();
    }
    return FS2K_OK;
}
