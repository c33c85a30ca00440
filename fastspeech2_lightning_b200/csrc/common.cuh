// Shared device/host helpers for libfs2k (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fs2k.h"

#define FS2K_CHECK_LAUNCH()                                   \
    do {                                                      \
        cudaError_t e__ = cudaGetLastError();                 \
        if (e__ != cudaSuccess) return fs2k_set_cuda_error(e__); \
    } while (0)

#define FS2K_REQUIRE(cond, code) \
    do {                         \
        if (!(cond)) return (code); \
    } while (0)

int fs2k_set_cuda_error(cudaError_t e);

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------
// Every kernel of the library is launched through fs2k_launch with programmatic stream serialization allowed:
// the next kernel of the stream (or the next kernel node of a captured graph) may be scheduled as soon as every
// CTA of this one has executed griddepcontrol.launch_dependents — which every kernel does first thing — and
// then blocks in griddepcontrol.wait until this grid has completed and its writes are visible.  Launch latency
// and kernel prologues (barrier init, TMEM allocation, parameter loads) overlap the tail of the previous kernel;
// nothing that reads or writes global memory runs before the wait.  Exception: the small, long, latency-bound
// grids (MAS DP, CTC recursions: one CTA per utterance) use fs2k_launch_serial and never trigger early (see below).
// fs2k_set_pdl(0) = attribute off everywhere.
extern int g_fs2k_pdl_enabled;  // lib.cu

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// A kernel that does work between launch and pdl_wait() (barrier init, TMEM allocation) must re-acquire, AFTER the wait,
// every pointer to data an earlier kernel may have produced: loads through `const __restrict__` pointers are ld.global.nc —
// immutable for the kernel's lifetime as far as the compiler knows — and ptxas hoists them above griddepcontrol.wait to hide
// their latency (seen in SASS: LDG.E.CONSTANT ahead of ACQBULK).  Under a deep chain of programmatically launched kernels
// that read happens before the producer has run.  The empty volatile asm makes the pointer value unknown until after the wait.
// tests/test_cabi.py scans the SASS of every kernel for a global access ahead of ACQBULK.
template <typename T>
__device__ __forceinline__ T* pdl_acquire(T* p) {
    asm volatile("" : "+l"(p));
    return p;
}
__device__ __forceinline__ void pdl_prologue() {
    pdl_launch_dependents();
    pdl_wait();
}

template <typename... KArgs, typename... Args>
static inline void fs2k_launch_impl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                    cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && g_fs2k_pdl_enabled) ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);  // errors surface in FS2K_CHECK_LAUNCH
}
template <typename... KArgs, typename... Args>
static inline void fs2k_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                               Args&&... args) {
    fs2k_launch_impl(true, kernel, grid, block, smem, stream, static_cast<Args&&>(args)...);
}
// Plain stream-ordered launch for the small latency-bound grids (one CTA per utterance: MAS DP, CTC recursions).
// Launched programmatically they are placed while the previous kernel still fills the machine, several CTAs end up on
// the same SM instead of one per SM, and the serial recursion runs 2.6x slower (measured on the MAS DP at B = 32).
template <typename... KArgs, typename... Args>
static inline void fs2k_launch_serial(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                      Args&&... args) {
    fs2k_launch_impl(false, kernel, grid, block, smem, stream, static_cast<Args&&>(args)...);
}

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// grid of the column-reduction kernels (colsum / colstats / bn_bwd_stats): blockIdx.y = 128-channel block,
// blockIdx.x = row chunk, ≈ 4 CTAs per SM in total, at least 32 rows per CTA
static inline dim3 col_reduce_grid(long M, int C, long* rows_per_cta) {
    const int cb = cdiv(C, 128);
    long target = (148L * 4 + cb - 1) / cb;
    long rows = (M + target - 1) / target;
    if (rows < 32) rows = 32;
    rows = (rows + 7) / 8 * 8;
    *rows_per_cta = rows;
    return dim3((unsigned)((M + rows - 1) / rows), (unsigned)cb);
}

namespace fs2k {

constexpr int kWarp = 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 128-bit accesses (read-once / write-once data: keep it out of L1)
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w));
}

__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }

// counter-based uniform in [0,1): splitmix64 finaliser of (seed, index).  Dropout masks are regenerated
// from the seed in the backward pass instead of being stored.
__device__ __forceinline__ unsigned long long hash_u64(unsigned long long seed, unsigned long long i) {
    unsigned long long z = seed * 0x9E3779B97F4A7C15ull + i + 0x632BE59BD9B4E019ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float hash_uniform(unsigned long long seed, unsigned long long i) {
    return (float)(hash_u64(seed, i) >> 40) * (1.0f / 16777216.0f);
}
// Dropout decisions: ONE 64-bit hash decides four consecutive elements (group = element index >> 2, 16 bits each against
// thr16 = round(p·65536)); the keep probability is exactly 1 − thr16/65536 and the survivors are scaled by its inverse.
// Every kernel that drops (forward) or re-applies the mask (backward) goes through these helpers, so masks always agree.
__device__ __forceinline__ uint32_t drop_thr16(float p) { return p > 0.f ? (uint32_t)(p * 65536.0f + 0.5f) : 0u; }
__device__ __forceinline__ float drop_inv_keep(uint32_t thr16) { return 65536.0f / (float)(65536u - thr16); }
__device__ __forceinline__ bool drop_keep_g(unsigned long long seed, unsigned long long group, int lane, uint32_t thr16) {
    return ((uint32_t)(hash_u64(seed, group) >> (16 * lane)) & 0xFFFFu) >= thr16;
}
__device__ __forceinline__ bool drop_keep(unsigned long long seed, unsigned long long e, uint32_t thr16) {
    return drop_keep_g(seed, e >> 2, (int)(e & 3), thr16);
}
// elements e .. e+3 with e % 4 == 0
__device__ __forceinline__ void drop_apply4(float4& v, unsigned long long seed, unsigned long long e, uint32_t thr16, float inv_keep) {
    const unsigned long long h = hash_u64(seed, e >> 2);
    v.x = ((uint32_t)h & 0xFFFFu) >= thr16 ? v.x * inv_keep : 0.f;
    v.y = ((uint32_t)(h >> 16) & 0xFFFFu) >= thr16 ? v.y * inv_keep : 0.f;
    v.z = ((uint32_t)(h >> 32) & 0xFFFFu) >= thr16 ? v.z * inv_keep : 0.f;
    v.w = ((uint32_t)(h >> 48) & 0xFFFFu) >= thr16 ? v.w * inv_keep : 0.f;
}
// Dropout under CUDA-graph replay: the by-value seeds of a captured launch are frozen, so every dropout kernel
// adds *g_seed_base (a device counter the host rewrites before each replay; null → 0) to its seed on entry.
// One copy of the pointer per translation unit; FS2K_DEFINE_SEED_BASE_SETTER(tag) exports its setter and
// fs2k_set_dropout_seed_base (lib.cu) fans out to all of them.
static __device__ const unsigned long long* g_seed_base = nullptr;
__device__ __forceinline__ unsigned long long seed_with_base(unsigned long long seed) {
    const unsigned long long* p = g_seed_base;
    return p ? seed + *p : seed;
}
#define FS2K_DEFINE_SEED_BASE_SETTER(tag)                                                                  \
    extern "C" int fs2k_seed_base_set_##tag(const void* dev_ptr) {                                        \
        cudaError_t e = cudaMemcpyToSymbol(fs2k::g_seed_base, &dev_ptr, sizeof(dev_ptr));                  \
        return e == cudaSuccess ? FS2K_OK : fs2k_set_cuda_error(e);                                        \
    }

// scaled keep mask of attention-probability dropout for element (b,h,q,k)
__device__ __forceinline__ float attn_keep(unsigned long long seed, float p, float inv_keep, int b, int h, int q, int k,
                                           int H, int L) {
    if (p <= 0.f) return 1.f;
    const unsigned long long idx = (((unsigned long long)b * H + h) * L + q) * L + k;
    return hash_uniform(seed, idx) >= p ? inv_keep : 0.f;
}

}  // namespace fs2k
