// bf16 tensor-core GEMM / 1-D convolution for sm_100a (the `bf16` arithmetic mode, BASELINE configs[2]):
// operands rounded to bf16, products accumulated in fp32 in TMEM (tcgen05.mma kind::f16), everything around the
// contraction — bias, folded BatchNorm, activation, dropout, residual stream, row mask — in fp32.
//
//   C[(b,l), n] = (dropout(act((Σ_tap Σ_k A[b, l+tap−pad, k] · W[tap][n][k] + bias[n])·scale[n] + shift[n])·alpha)
//                 + residual[(b,l), n]) · row_mask[(b,l)]
//
// Operand sources
//   A fp32 in HBM (A_F32): four producer warps read 128 × 64 fp32 with coalesced 16-byte loads, round to bf16 in
//       registers and store the tile in the 128-byte-swizzled K-major layout the MMA reads — the activations of the
//       model stay fp32 in HBM (residual stream, LayerNorm / BatchNorm inputs) and no separate cast pass exists.
//       NSPLIT = 3 additionally stores lo = x − bf16(x) and issues hi·hi + hi·lo + lo·hi (fp32-level accuracy for
//       the HBM-bound K = 256 contractions, where two more MMAs per k-step are free).
//   A bf16 in HBM: TMA boxes of a 3-D map [B][L][K] (taps = time-shifted boxes, out-of-range rows zero-filled).
//   W bf16 [taps][N][K] (K-major, forward) by TMA, or — B_MN, the data-gradient GEMM dX = G·W — the SAME weight
//       array read as an MN-major operand ([taps][N][K] with the contraction over N): no transposed copy of the
//       weights exists in this mode.
// One 128 × block_n tile per CTA, 256 threads; tiles of 128 columns need < 113 KB of shared memory so two CTAs share
// an SM and one's epilogue overlaps the other's main loop.
#include <cstring>

#include "bf16_common.cuh"

namespace fs2k {

constexpr int HB_BK = 64;        // bf16 per 128-byte swizzle row
constexpr int HB_THREADS = 256;
constexpr int HB_WARPS = HB_THREADS / 32;
constexpr int HB_MAX_STAGES = 8;
constexpr int HB_CONV_THREADS = 128;  // warps 4..7 convert the fp32 A tile
constexpr int HB_ROWS = 2;            // rows a warp keeps in flight in the store pass

struct HbEpilogue {
    const float* bias; const float* scale; const float* shift;
    int act; float alpha;
    const float* residual; int ldr;
    const uint8_t* row_mask;
    float* C; int ldc;                 // final value, fp32 (optional)
    __nv_bfloat16* C16; int ldc16;     // final value, bf16 (optional)
    float* P32; __nv_bfloat16* P16; int ldp;  // pre-activation acc + bias (optional; saved for act' in the backward)
    float drop_p; unsigned long long seed;
};

template <bool A_F32, int NSPLIT, bool B_MN, bool DROPOUT>
__global__ void __launch_bounds__(HB_THREADS, 2)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmBl, const float* __restrict__ A32, int lda, int L, long M_total,
                 int K, int N, int block_n, int taps, int pad, int tiles_per_b, int n_stages, int tmem_cols, HbEpilogue ep) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_full[HB_MAX_STAGES], s_empty[HB_MAX_STAGES], s_tmem_full;
    __shared__ uint32_t s_tmem_base;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t a_bytes = HB_BM * HB_BK * 2;                                            // 16 KB
    const uint32_t b_bytes = B_MN ? (uint32_t)((block_n + 63) / 64) * 8192u : (uint32_t)block_n * 128u;
    const uint32_t stage_bytes = (a_bytes + b_bytes) * (NSPLIT == 3 ? 2 : 1);
    // stage layout: [A hi][A lo (NSPLIT 3)][B hi][B lo (NSPLIT 3)]
    const uint32_t off_alo = a_bytes, off_b = a_bytes * (NSPLIT == 3 ? 2 : 1), off_blo = off_b + b_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int tile_m = blockIdx.x, n0 = blockIdx.y * block_n;
    const int b_idx = tile_m / tiles_per_b, l0 = (tile_m % tiles_per_b) * HB_BM;
    const int nk = (K + HB_BK - 1) / HB_BK;
    const int iters = taps * nk;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(smem_u32(&s_full[s]), 1 + (A_F32 ? HB_CONV_THREADS : 0));
            mbar_init(smem_u32(&s_empty[s]), 1);
        }
        mbar_init(smem_u32(&s_tmem_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;
    pdl_wait();  // setup above overlapped the previous kernel's tail; global memory is touched below

    if (warp == 0) {
        if (lane == 0) {
            // ================= TMA producer =================
            for (int it = 0; it < iters; ++it) {
                const int s = it % n_stages, ph = (it / n_stages) & 1;
                mbar_wait(smem_u32(&s_empty[s]), ph ^ 1);
                const int tap = it / nk, k0 = (it - tap * nk) * HB_BK;
                const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t bar = smem_u32(&s_full[s]);
                mbar_expect_tx(bar, (A_F32 ? 0u : a_bytes) + b_bytes * (NSPLIT == 3 ? 2u : 1u));
                if (!A_F32) tma_load_3d(st, &tmA, bar, k0, l0 + tap - pad, b_idx);
                if (B_MN) {
                    // weights [taps][Kc][N] read as an MN-major operand; the transposed convolution visits the taps in reverse
                    const int wt = taps - 1 - tap;
                    for (int jb = 0; jb * 64 < block_n; ++jb) {
                        tma_load_3d(st + off_b + jb * 8192, &tmB, bar, n0 + jb * 64, k0, wt);
                        if (NSPLIT == 3) tma_load_3d(st + off_blo + jb * 8192, &tmBl, bar, n0 + jb * 64, k0, wt);
                    }
                } else {
                    tma_load_2d(st + off_b, &tmB, bar, k0, tap * N + n0);
                    if (NSPLIT == 3) tma_load_2d(st + off_blo, &tmBl, bar, k0, tap * N + n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ================= MMA issuer =================
            const uint32_t idesc = hb_idesc(block_n, 0, B_MN ? 1 : 0);
            for (int it = 0; it < iters; ++it) {
                const int s = it % n_stages, ph = (it / n_stages) & 1;
                mbar_wait(smem_u32(&s_full[s]), ph);
                tc_fence_after();
                const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
#pragma unroll
                for (int k = 0; k < HB_BK / 16; ++k) {
                    const uint64_t ad = hb_desc_k(st + k * 32);
                    const uint64_t bd = B_MN ? hb_desc_mn(st + off_b + k * 2048, 8192) : hb_desc_k(st + off_b + k * 32);
                    hb_mma(tmem_base, ad, bd, idesc, (it | k) ? 1u : 0u);
                    if (NSPLIT == 3) {
                        const uint64_t adl = hb_desc_k(st + off_alo + k * 32);
                        const uint64_t bdl = B_MN ? hb_desc_mn(st + off_blo + k * 2048, 8192) : hb_desc_k(st + off_blo + k * 32);
                        hb_mma(tmem_base, ad, bdl, idesc, 1u);
                        hb_mma(tmem_base, adl, bd, idesc, 1u);
                    }
                }
                tc_commit(smem_u32(&s_empty[s]));  // frees the stage when these MMAs have read it
            }
            tc_commit(smem_u32(&s_tmem_full));
        }
    } else if (A_F32 && warp >= 4) {
        // ================= A converter: fp32 HBM → bf16 swizzled K-major tile =================
        const int ct = threadIdx.x - 128;          // 0..127
        const int c4 = ct & 15, r_in = ct >> 4;    // 16 threads cover one 64-float row; 8 rows per pass
        const long row_base = (long)b_idx * L;
        for (int it = 0; it < iters; ++it) {
            const int s = it % n_stages, ph = (it / n_stages) & 1;
            const int tap = it / nk, k0 = (it - tap * nk) * HB_BK;
            const int kc = k0 + c4 * 4;
            float4 v[16];
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                const int r = p * 8 + r_in;
                const long l = (long)l0 + r + tap - pad;
                v[p] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (l >= 0 && l < L && row_base + l < M_total && kc < K)
                    v[p] = ld_stream(reinterpret_cast<const float4*>(A32 + (size_t)(row_base + l) * lda + kc));
            }
            mbar_wait(smem_u32(&s_empty[s]), ph ^ 1);
            uint8_t* st = smem + (size_t)s * stage_bytes;
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                const int r = p * 8 + r_in;
                const uint32_t off = (uint32_t)r * 128u + (uint32_t)(((c4 >> 1) ^ (r & 7)) << 4) + (uint32_t)(c4 & 1) * 8u;
                if (NSPLIT == 3) {
                    uint2 hi, lo;
                    hb_split4(v[p], hi, lo);
                    *reinterpret_cast<uint2*>(st + off) = hi;
                    *reinterpret_cast<uint2*>(st + off_alo + off) = lo;
                } else {
                    *reinterpret_cast<uint2*>(st + off) = hb_pack4(v[p]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes → visible to the MMA proxy
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_full[s])) : "memory");
        }
    }

    // ===== epilogue stage 1 (all warps): raw accumulators TMEM → registers → staging tile (row = TMEM lane) =====
    const int pitch = block_n + 4;
    __syncwarp();
    {
        mbar_wait(smem_u32(&s_tmem_full), 0);
        tc_fence_after();
        const int q = warp & 3;          // TMEM lane quadrant this warp may read
        const int half = warp >> 2;      // two warps share a quadrant and alternate 16-column chunks
        const int row = q * 32 + lane;
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

    // ===== epilogue stage 2 (all warps): column math, residual, row mask, coalesced stores =====
    {
        const float* stag = reinterpret_cast<const float*>(smem);
        const int nv = block_n >> 2;  // float4 per row (≤ 64): lane owns float4 columns lane and lane+32
        const bool has[2] = {lane < nv, lane + 32 < nv};
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f), one4 = make_float4(1.f, 1.f, 1.f, 1.f);
        float4 bias4[2], sc4[2], sh4[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int qv = lane + 32 * j;
            bias4[j] = (has[j] && ep.bias) ? __ldg(reinterpret_cast<const float4*>(ep.bias + n0) + qv) : zero4;
            sc4[j] = (has[j] && ep.scale) ? __ldg(reinterpret_cast<const float4*>(ep.scale + n0) + qv) : one4;
            sh4[j] = (has[j] && ep.scale) ? __ldg(reinterpret_cast<const float4*>(ep.shift + n0) + qv) : zero4;
        }
        const int rows_valid = (int)min((long)HB_BM, min((long)L - l0, M_total - ((long)b_idx * L + l0)));
        const float drop_p = DROPOUT ? ep.drop_p : 0.f;
        const uint32_t thr16 = drop_thr16(drop_p);
        const float inv_keep = drop_inv_keep(thr16);
        const unsigned long long seed = DROPOUT ? seed_with_base(ep.seed) : 0ull;
        const int act = ep.act;
        const float alpha = ep.alpha;
        // HB_ROWS rows per warp trip: every global read of the group (residual, row mask) is issued before any is used
        for (int rb = warp; rb < rows_valid; rb += HB_WARPS * HB_ROWS) {
            float4 res[HB_ROWS][2];
            float rm[HB_ROWS];
#pragma unroll
            for (int u = 0; u < HB_ROWS; ++u) {
                const int r = rb + u * HB_WARPS;
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
            for (int u = 0; u < HB_ROWS; ++u) {
                const int r = rb + u * HB_WARPS;
                if (r >= rows_valid) break;  // warp-uniform
                const long m = (long)b_idx * L + l0 + r;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (!has[j]) continue;
                    const int qv = lane + 32 * j;
                    float4 v = *reinterpret_cast<const float4*>(stag + (size_t)r * pitch + qv * 4);
                    v.x += bias4[j].x; v.y += bias4[j].y; v.z += bias4[j].z; v.w += bias4[j].w;
                    if (ep.P32) *reinterpret_cast<float4*>(ep.P32 + (size_t)m * ep.ldp + n0 + qv * 4) = v;
                    if (ep.P16) *reinterpret_cast<uint2*>(ep.P16 + (size_t)m * ep.ldp + n0 + qv * 4) = hb_pack4(v);
                    v.x = hb_act(v.x * sc4[j].x + sh4[j].x, act) * alpha;
                    v.y = hb_act(v.y * sc4[j].y + sh4[j].y, act) * alpha;
                    v.z = hb_act(v.z * sc4[j].z + sh4[j].z, act) * alpha;
                    v.w = hb_act(v.w * sc4[j].w + sh4[j].w, act) * alpha;
                    if (DROPOUT) {
                        const unsigned long long e = (unsigned long long)m * N + n0 + qv * 4;
                        drop_apply4(v, seed, e, thr16, inv_keep);
                    }
                    v.x = (v.x + res[u][j].x) * rm[u]; v.y = (v.y + res[u][j].y) * rm[u];
                    v.z = (v.z + res[u][j].z) * rm[u]; v.w = (v.w + res[u][j].w) * rm[u];
                    if (ep.C) *reinterpret_cast<float4*>(ep.C + (size_t)m * ep.ldc + n0 + qv * 4) = v;
                    if (ep.C16) *reinterpret_cast<uint2*>(ep.C16 + (size_t)m * ep.ldc16 + n0 + qv * 4) = hb_pack4(v);
                }
            }
        }
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// x → hi = bf16(x) (and lo = bf16(x − hi) when lo != nullptr)
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ x, long n, __nv_bfloat16* __restrict__ hi,
                                                        __nv_bfloat16* __restrict__ lo) {
    pdl_prologue();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float v = x[i];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[i] = h;
        if (lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

// bf16 → fp32 (the widening side of the bf16 gradient exchange: optim.FusedAdamW.allreduce_range)
__global__ void __launch_bounds__(256) cast_f32_kernel(const __nv_bfloat16* __restrict__ x, long n, float* __restrict__ y) {
    pdl_prologue();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) y[i] = __bfloat162float(x[i]);
}

}  // namespace fs2k

using namespace fs2k;

bool fs2k_gemm_bf16_panel_ok(int K, int N, int taps, int a_is_bf16, bool has_lo, bool has_scale);
int fs2k_gemm_bf16_panel_launch(const float* A, int lda, long M, int K, const void* W, int w_mn, int N, const float* bias, int act,
                                float alpha, const float* residual, int ldr, const uint8_t* row_mask, float* C, int ldc, void* C16,
                                int ldc16, float* P32, void* P16, int ldp, float dropout_p, long seed, const void* dact_pre16,
                                cudaStream_t s);

extern "C" int fs2k_cast_f32(const void* x_bf16, long n, float* y, fs2k_stream_t stream) {
    FS2K_REQUIRE(n >= 0, FS2K_ERR_BAD_SHAPE);
    if (n == 0) return FS2K_OK;
    FS2K_REQUIRE(x_bf16 && y, FS2K_ERR_NULL);
    long g = (n + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    fs2k_launch(cast_f32_kernel, dim3((unsigned)g), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)x_bf16, n, y);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_cast_bf16(const float* x, long n, void* hi, void* lo, fs2k_stream_t stream) {
    FS2K_REQUIRE(n >= 0, FS2K_ERR_BAD_SHAPE);
    if (n == 0) return FS2K_OK;
    FS2K_REQUIRE(x && hi, FS2K_ERR_NULL);
    long g = (n + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    fs2k_launch(cast_bf16_kernel, dim3((unsigned)g), dim3(256), 0, (cudaStream_t)stream, x, n, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

// K = contraction length, N = output columns.  w_mn = 0: W is [taps][N][K]; w_mn = 1: W is [taps][K][N] … i.e. the
// forward layer's own [taps][N_fwd = K][K_fwd = N] array, read MN-major (data-gradient GEMM).
extern "C" int fs2k_gemm_bf16_supported(int K, int N, int lda, int taps, int a_is_bf16, int w_mn) {
    if (K <= 0 || N <= 0 || taps < 1) return 0;
    if (a_is_bf16 ? ((lda & 7) || (K & 7)) : ((lda & 3) || (K & 3))) return 0;
    if (w_mn) {
        if (N & 7) return 0;                       // TMA: 16-byte global strides of the [K][N] rows
    } else if (K & 7) {
        return 0;
    }
    if (N <= 256) return (N % 16) == 0;
    return (N % 128) == 0;
}

static int gemm_bf16_impl(const void* A, int a_is_bf16, int lda, int B, int L, int K, const void* W_hi, const void* W_lo,
                          int w_mn, int N, int taps, int pad, const float* bias, const float* scale, const float* shift,
                          int act, float alpha, const float* residual, int ldr, const uint8_t* row_mask, float* C, int ldc,
                          void* C16, int ldc16, float* P32, void* P16, int ldp, float dropout_p, long seed, int block_n_hint,
                          const void* dact_pre16, fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && L >= 0 && K > 0 && N > 0 && taps >= 1 && pad >= 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(act >= 0 && act <= 3, FS2K_ERR_UNSUPPORTED);
    FS2K_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(fs2k_gemm_bf16_supported(K, N, lda, taps, a_is_bf16, w_mn), FS2K_ERR_UNSUPPORTED);
    const long M = (long)B * L;
    if (M == 0) return FS2K_OK;
    FS2K_REQUIRE(A && W_hi && (C || C16 || P32 || P16), FS2K_ERR_NULL);
    FS2K_REQUIRE(!scale || shift, FS2K_ERR_NULL);
    FS2K_REQUIRE((!C || (ldc & 3) == 0) && (!residual || (ldr & 3) == 0) && (!C16 || (ldc16 & 3) == 0) &&
                     (!(P32 || P16) || (ldp & 3) == 0), FS2K_ERR_UNSUPPORTED);
    const int nsplit = W_lo ? 3 : 1;
    FS2K_REQUIRE(nsplit == 1 || !a_is_bf16, FS2K_ERR_UNSUPPORTED);  // the lo part of A comes from its fp32 source
    // the model's dominant class (fp32 activations, K <= 256, one tap): row-panel kernel, gemm_bf16_panel.cu
    // (block_n_hint < 0 forces the tile-per-CTA kernel below — A/B measurements)
    if ((block_n_hint >= 0 || dact_pre16) && fs2k_gemm_bf16_panel_ok(K, N, taps, a_is_bf16, W_lo != nullptr, scale != nullptr))
        return fs2k_gemm_bf16_panel_launch((const float*)A, lda, M, K, W_hi, w_mn, N, bias, act, alpha, residual, ldr, row_mask, C, ldc,
                                           C16, ldc16, P32, P16, ldp, dropout_p, seed, dact_pre16, (cudaStream_t)stream);
    FS2K_REQUIRE(!dact_pre16, FS2K_ERR_UNSUPPORTED);  // the fused activation-derivative epilogue lives in the row-panel kernel
    int block_n = N <= 256 ? N : 128;
    if (N > 128 && N % 128 == 0) block_n = 128;            // two CTAs per SM: epilogue of one overlaps the other's main loop
    if (block_n_hint == 256 && N % 256 == 0) block_n = 256;
    if (block_n_hint > 0 && block_n_hint < block_n && N % block_n_hint == 0 && block_n_hint % 16 == 0) block_n = block_n_hint;
    {
        // small-M problems (encoder, predictors): narrow the column tile until the grid covers most of the SMs
        const long tiles_m = (taps == 1) ? (M + HB_BM - 1) / HB_BM : (long)B * ((L + HB_BM - 1) / HB_BM);
        while (block_n > 32 && (block_n / 2) % 16 == 0 && N % (block_n / 2) == 0 && tiles_m * (N / block_n) < 100) block_n /= 2;
    }
    EncodeTiledFn encode = get_encode();
    FS2K_REQUIRE(encode != nullptr, FS2K_ERR_ARCH);

    const int Bm = taps == 1 ? 1 : B;
    const long Lm = taps == 1 ? M : L;
    FS2K_REQUIRE(Lm < (1L << 31), FS2K_ERR_UNSUPPORTED);
    const int tiles_per_b = (int)((Lm + HB_BM - 1) / HB_BM);

    CUtensorMap tmA, tmB, tmBl;
    memset(&tmA, 0, sizeof(tmA));
    if (a_is_bf16) {
        cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)Lm, (cuuint64_t)Bm};
        cuuint64_t strides[2] = {(cuuint64_t)lda * 2, (cuuint64_t)Lm * lda * 2};
        cuuint32_t box[3] = {HB_BK, HB_BM, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)A, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fs2k_set_cuda_error(cudaErrorInvalidValue);
    }
    auto make_w = [&](CUtensorMap* tm, const void* w) -> bool {
        if (w_mn) {
            cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)K, (cuuint64_t)taps};
            cuuint64_t strides[2] = {(cuuint64_t)N * 2, (cuuint64_t)K * N * 2};
            cuuint32_t box[3] = {64, HB_BK, 1};
            cuuint32_t estr[3] = {1, 1, 1};
            return encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)taps * N};
        cuuint64_t strides[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {HB_BK, (cuuint32_t)block_n};
        cuuint32_t estr[2] = {1, 1};
        return encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    if (!make_w(&tmB, W_hi) || !make_w(&tmBl, W_lo ? W_lo : W_hi)) return fs2k_set_cuda_error(cudaErrorInvalidValue);

    HbEpilogue ep{bias, scale, shift, act, alpha, residual, ldr, row_mask, C, ldc, (__nv_bfloat16*)C16, ldc16,
                  P32, (__nv_bfloat16*)P16, ldp, dropout_p, (unsigned long long)seed};
    const size_t b_bytes = w_mn ? (size_t)((block_n + 63) / 64) * 8192 : (size_t)block_n * 128;
    const size_t stage = ((size_t)HB_BM * HB_BK * 2 + b_bytes) * (nsplit == 3 ? 2 : 1);
    const size_t staging = (size_t)HB_BM * (block_n + 4) * 4;
    // ring depth: enough to cover the K loop (taps·K/64 iterations), at most what two co-resident CTAs can hold
    const int iters = taps * ((K + HB_BK - 1) / HB_BK);
    int n_stages = (int)((110 * 1024) / stage);
    if (n_stages < 3) n_stages = (int)((225 * 1024) / stage);   // wide tiles: one CTA per SM
    if (n_stages > HB_MAX_STAGES) n_stages = HB_MAX_STAGES;
    if (n_stages > iters) n_stages = iters < 2 ? 2 : iters;
    FS2K_REQUIRE(n_stages >= 2, FS2K_ERR_UNSUPPORTED);
    size_t smem = stage * n_stages;
    if (smem < staging) smem = staging;
    smem += 1024;  // manual 1024-byte alignment of the swizzled tiles
    FS2K_REQUIRE(smem <= 227 * 1024, FS2K_ERR_UNSUPPORTED);
    int tmem_cols = 32;
    while (tmem_cols < block_n) tmem_cols <<= 1;
    dim3 grid(Bm * tiles_per_b, N / block_n);
    cudaStream_t s = (cudaStream_t)stream;
    const bool drop = dropout_p > 0.f;
    cudaError_t e = cudaSuccess;
    auto launch = [&](auto kernel) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return;
        fs2k_launch(kernel, dim3(grid), dim3(HB_THREADS), smem, s, tmA, tmB, tmBl, (const float*)A, lda, (int)Lm, M, K, N, block_n,
                    taps, pad, tiles_per_b, n_stages, tmem_cols, ep);
    };
#define HB_DISPATCH(AF, NS, MN)                                             \
    do {                                                                    \
        if (drop) launch(gemm_bf16_kernel<AF, NS, MN, true>);               \
        else launch(gemm_bf16_kernel<AF, NS, MN, false>);                   \
    } while (0)
    if (a_is_bf16) {
        if (w_mn) HB_DISPATCH(false, 1, true); else HB_DISPATCH(false, 1, false);
    } else if (nsplit == 3) {
        if (w_mn) HB_DISPATCH(true, 3, true); else HB_DISPATCH(true, 3, false);
    } else {
        if (w_mn) HB_DISPATCH(true, 1, true); else HB_DISPATCH(true, 1, false);
    }
#undef HB_DISPATCH
    if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_gemm_bf16(const void* A, int a_is_bf16, int lda, int B, int L, int K, const void* W_hi, const void* W_lo,
                              int w_mn, int N, int taps, int pad, const float* bias, const float* scale, const float* shift,
                              int act, float alpha, const float* residual, int ldr, const uint8_t* row_mask, float* C, int ldc,
                              void* C16, int ldc16, float* P32, void* P16, int ldp, float dropout_p, long seed, int block_n_hint,
                              fs2k_stream_t stream) {
    return gemm_bf16_impl(A, a_is_bf16, lda, B, L, K, W_hi, W_lo, w_mn, N, taps, pad, bias, scale, shift, act, alpha, residual, ldr, row_mask,
                          C, ldc, C16, ldc16, P32, P16, ldp, dropout_p, seed, block_n_hint, nullptr, stream);
}

// Data-gradient GEMM fused with the derivative of the activation that followed the forward layer (and its dropout mask):
//   out[m, n] = (Σ_k G[m, k] · W[k][n]) · act'(pre[m, n]) · keep(seed, m·N + n) / (1 − p)
// `pre_bf16` [M, N] is the forward GEMM's saved pre-activation (its P16 output).  K <= 256, fp32 G, one tap (row-panel kernel).
extern "C" int fs2k_gemm_bf16_dact(const float* G, int ldg, long M, int K, const void* W_hi, int w_mn, int N, const void* pre_bf16,
                                   int act, float alpha, float dropout_p, long seed, float* C, void* C16, fs2k_stream_t stream) {
    FS2K_REQUIRE(M >= 0 && M < (1L << 31) && K > 0 && N > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(pre_bf16 != nullptr, FS2K_ERR_NULL);
    FS2K_REQUIRE(fs2k_gemm_bf16_panel_ok(K, N, 1, 0, false, false), FS2K_ERR_UNSUPPORTED);
    return gemm_bf16_impl(G, 0, ldg, 1, (int)M, K, W_hi, nullptr, w_mn, N, 1, 0, nullptr, nullptr, nullptr, act, alpha, nullptr, 0, nullptr,
                          C, N, C16, N, nullptr, nullptr, N, dropout_p, seed, 0, pre_bf16, stream);
}

FS2K_DEFINE_SEED_BASE_SETTER(gemm_bf16)
