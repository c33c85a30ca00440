// Batched monotonic alignment search on the device.
//
// Replaces the reference's device→host→numba→device round trip
//   VarianceAdaptor.binarize_attention   fs2/variance_adaptor.py:160-181
//   mas_width1 / b_mas                   fs2/attn/alignment.py:48-85
// Result is bit-exact w.r.t. the numba code for the same fp32 log-probabilities: the forward
// recurrence does the same fp32 add of max(left-up, up); the direction taken by the backtrack
// (`log_p[i-1,j-1] >= log_p[i-1,j]`, tie → diagonal) is recorded as one bit per cell while the
// row is live in registers, so the backtrack never re-reads log_p.
//
// Layout: one CTA per utterance; thread t owns column(s) t + c·blockDim of the current row
// (`prev[c]` lives in a register), the left neighbour comes from `__shfl_up` and, across warp
// and chunk boundaries, from a double-buffered shared-memory word; one `__syncthreads` per mel
// frame.  The DP is latency bound (F sequential steps), so
//   * rows are streamed PF frames ahead with `cp.async` into a shared-memory ring: each thread copies
//     exactly the elements it will read, so completion is tracked per thread with `cp.async.wait_group`
//     and costs no extra barrier (a register ring does not work: ptxas rotates it with moves that wait on
//     the load just issued);
//   * the direction words of DB frames are collected in shared memory and written out in one coalesced
//     burst, so the per-frame barrier never has a global store in flight.
// (One column per thread is kept on purpose: four columns per thread was measured 2.5× slower at T = 80 —
// fewer warps to overlap the `logf` chains — and no faster at T = 1000.)
// The backtrack is done by warp 0, 32 frames per step: lane r fetches the two direction words that
// can hold frame (top−r)'s column, then the dependent chain runs through shuffles only.
#include "common.cuh"

namespace fs2k {

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int MAS_DB = 32;  // frames of direction words buffered in shared memory between flushes

template <int CHUNKS, int PF, bool TAKE_LOG>
__global__ void __launch_bounds__(1024, 1)
mas_dp_kernel(const float* __restrict__ attn,   // [B,F,T] log-probs (or probs if TAKE_LOG)
              const int* __restrict__ in_lens,   // [B] text lengths
              const int* __restrict__ out_lens,  // [B] mel lengths
              int F, int T, int W,               // W = ceil(T/32) direction words per frame
              uint32_t* __restrict__ dirs,       // [B,F,W] workspace
              int* __restrict__ path,            // [B,F] column of frame f, −1 on padding
              int* __restrict__ durations)       // [B,T]
{
    pdl_wait();  // launched with fs2k_launch_serial (one CTA per SM matters here) and never triggers its dependents early
    extern __shared__ float ring[];  // [PF][CHUNKS·blockDim] then uint32 sdir[MAS_DB][W]
    const int b = blockIdx.x;
    const int n_text = min(in_lens[b], T);
    const int n_mel = min(out_lens[b], F);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int ncols = CHUNKS * blockDim.x;
    uint32_t* sdir = reinterpret_cast<uint32_t*>(ring + (size_t)PF * ncols);
    const float NEG_INF = -INFINITY;
    __shared__ float bnd[2][CHUNKS * 32];

    const float* x = attn + (size_t)b * F * T;
    uint32_t* d = dirs + (size_t)b * F * W;
    int* p = path + (size_t)b * F;
    int* dur = durations + (size_t)b * T;

    for (int t = tid; t < T; t += blockDim.x) dur[t] = 0;
    for (int f = max(n_mel, 0) + tid; f < F; f += blockDim.x) p[f] = -1;
    if (n_mel <= 0 || n_text <= 0) {
        for (int f = tid; f < F; f += blockDim.x) p[f] = -1;
        return;
    }

    int col[CHUNKS];
    bool live[CHUNKS];
    float prev[CHUNKS];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        col[c] = c * blockDim.x + tid;
        live[c] = col[c] < n_text;
        // row 0: log_p[0,0] = x[0,0]; log_p[0,1:] = -inf        (alignment.py:53-54)
        float v = NEG_INF;
        if (col[c] == 0) {
            v = x[0];
            if (TAKE_LOG) v = logf(v);
        }
        prev[c] = v;
        if (lane == 31) bnd[0][c * nwarps + warp] = v;
    }
    auto issue = [&](int r) {  // stream row r into ring slot r % PF (one commit group per row, even when empty)
        if (r < n_mel) {
            float* slot = ring + (size_t)(r % PF) * ncols;
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c)
                if (live[c]) cp_async4(slot + c * blockDim.x + tid, x + (size_t)r * T + col[c]);
        }
        cp_async_commit();
    };
    for (int r = 1; r <= PF; ++r) issue(r);
    __syncthreads();

    for (int i = 1; i < n_mel; ++i) {
        cp_async_wait<PF - 1>();  // this thread's copies of row i have landed
        const float* slot = ring + (size_t)(i % PF) * ncols;
        const int par = (i - 1) & 1;
        uint32_t* drow = sdir + (size_t)((i - 1) % MAS_DB) * W;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            float xv = live[c] ? slot[c * blockDim.x + tid] : 0.f;
            if (TAKE_LOG) xv = logf(xv);
            float left = __shfl_up_sync(0xffffffffu, prev[c], 1);
            const int g = c * nwarps + warp;  // global warp slot == direction word index
            if (lane == 0) left = (g == 0) ? NEG_INF : bnd[par][g - 1];
            const float up = prev[c];
            // backtrack predicate of alignment.py:68 evaluated for cell (i, col); column 0 never moves
            const bool diag = (left >= up) && (col[c] >= 1);
            const uint32_t word = __ballot_sync(0xffffffffu, diag && live[c]);
            if (lane == 0 && g < W) drow[g] = word;
            const float m = up > left ? up : left;  // max(prev_log1, prev_log2), alignment.py:59
            const float cur = live[c] ? __fadd_rn(xv, m) : NEG_INF;
            prev[c] = cur;
            if (lane == 31) bnd[par ^ 1][g] = cur;
        }
        issue(i + PF);  // the slot just consumed is refilled PF frames ahead
        __syncthreads();
        // flush the buffered direction words of frames (i−n+1 … i) with coalesced stores
        const int n_buf = ((i - 1) % MAS_DB) + 1;
        if (n_buf == MAS_DB || i == n_mel - 1) {
            const int first = i - n_buf + 1;
            for (int e = tid; e < n_buf * W; e += blockDim.x) d[(size_t)first * W + e] = sdir[e];
            __syncthreads();  // sdir is rewritten by the next frame
        }
    }
    cp_async_wait<0>();

    // ---- backtrack (alignment.py:62-73), warp 0, 32 frames per step ----
    if (warp == 0) {
        int j = n_text - 1;
        for (int top = n_mel - 1; top >= 1; top -= 32) {
            const int row = top - lane;
            const int wj = j >> 5;
            uint32_t hi = 0, lo = 0;
            if (row >= 1) {
                hi = d[(size_t)row * W + wj];
                if (wj > 0) lo = d[(size_t)row * W + wj - 1];
            }
            const int base = 32 * (wj - 1);  // bit k of (hi:lo) ↔ column base + k
            int myj = -1;
#pragma unroll 4
            for (int r = 0; r < 32; ++r) {
                if (top - r < 1) break;  // warp-uniform
                const uint32_t h = __shfl_sync(0xffffffffu, hi, r);
                const uint32_t l = __shfl_sync(0xffffffffu, lo, r);
                if (lane == r) myj = j;
                const int k = j - base;
                const uint32_t bit = (k >= 32) ? ((h >> (k - 32)) & 1u) : ((l >> k) & 1u);
                j -= (int)bit;
            }
            if (row >= 1) p[row] = myj;
        }
        if (lane == 0) p[0] = j;  // opt[0, j] = 1   (alignment.py:73)
    }
    __syncthreads();
    // durations = attn_hard.sum over frames (variance_adaptor.py:267-268)
    for (int f = tid; f < n_mel; f += blockDim.x) atomicAdd(&dur[p[f]], 1);
}

// FOUR CONSECUTIVE COLUMNS PER THREAD, SKEWED BLOCKS OF 8 FRAMES (the default kernel).
// The one-column-per-thread kernel above is issue bound, not latency bound: at T = 1000 its 32 warps share four schedulers and
// every (warp, frame) costs ≈ 40 instructions for ONE cell per lane — ≈ 350 ns per frame whatever the barrier scheme (measured:
// one barrier per frame 2.95 ms, one per 8 frames 2.88 ms, shared-memory flags 4.9 ms at 8000 × 1000).  Here a thread owns columns
// 4t … 4t+3 of the row: one 16-byte shared-memory load and ONE shuffle per four cells (the other three left neighbours are the
// thread's own registers), ≈ 25 instructions per (thread, frame), and T = 1000 needs 8 warps — two per scheduler.
//   * Warps run one block of 8 frames apart (warp w works on block s − w in step s): inside a block a warp needs nothing from
//     the other warps but the previous frames' last column of warp w − 1, which that warp left in shared memory one step earlier
//     (three rotating buffers, read with two 16-byte loads per block).  One __syncthreads per 8 frames.
//   * Each thread streams exactly the 16 bytes per frame it will read, PF frames ahead, with cp.async into its own ring slots
//     (completion per thread: cp.async.wait_group, no barrier); a consumed slot is refilled at once.
//   * Directions: bit 4k + c of the thread's word for the block = the backtrack predicate of cell (frame 1 + 8·blk + k, column
//     4t + c); one coalesced 4-byte store per thread and block.  dirs layout: [ceil((F−1)/8)][blockDim] words per utterance.
//   * Columns ≥ n_text are not masked: a cell only feeds cells to its right, and the backtrack starts at n_text − 1 and moves
//     left, so whatever they hold is never read.  Arithmetic and tie rule are those of mas_dp_kernel — bit-exact.
// The backtrack (warp 0, 32 frames per step) gathers, per lane, the 32-column window [j − 31, j] of its frame from ≤ 9 words,
// then runs the dependent chain on shuffles.
constexpr int MQ_R = 8;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}

template <int PF, bool VEC>
__global__ void __launch_bounds__(1024, 1)
mas_dp_quad_kernel(const float* __restrict__ attn, const int* __restrict__ in_lens, const int* __restrict__ out_lens, int F, int T,
                   uint32_t* __restrict__ dirs, size_t dirs_stride, int* __restrict__ path, int* __restrict__ durations) {
    pdl_wait();
    extern __shared__ float4 ring4[];  // [PF][blockDim] (a thread-major ring with immediate slot offsets was measured 1.7× slower: the
                                       // LDGSTS write path wants a warp's 512 bytes contiguous)
    __shared__ __align__(16) float bnd[3][32][MQ_R];
    const int b = blockIdx.x;
    const int n_text = min(in_lens[b], T);
    const int n_mel = min(out_lens[b], F);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5, TQ = blockDim.x;
    const float NEG_INF = -INFINITY;
    const float* x = attn + (size_t)b * F * T;
    uint32_t* d = dirs + (size_t)b * dirs_stride;
    int* p = path + (size_t)b * F;
    int* dur = durations + (size_t)b * T;

    for (int t = tid; t < T; t += blockDim.x) dur[t] = 0;
    for (int f = max(n_mel, 0) + tid; f < F; f += blockDim.x) p[f] = -1;
    if (n_mel <= 0 || n_text <= 0) {
        for (int f = tid; f < F; f += blockDim.x) p[f] = -1;
        return;
    }
    const int col0 = 4 * tid;
    const bool any_live = col0 < n_text;            // this thread copies (and its results can matter)
    const bool warp_live = warp * 128 < n_text;     // warps right of the text only produce cells nobody reads
    float p0 = NEG_INF, p1 = NEG_INF, p2 = NEG_INF, p3 = NEG_INF;
    if (tid == 0) p0 = x[0];  // row 0: log_p[0,0] = x[0,0]; log_p[0,1:] = -inf (alignment.py:53-54)
    // frame 0's boundary cell plays "last frame of block −1": where warp w + 1 looks for it in its first step
    if (lane == 31) bnd[(warp + 2) % 3][warp][MQ_R - 1] = p3;

    const int n_steps = n_mel - 1;  // DP frames e = 0 … n_steps − 1 ↔ mel frame e + 1
    const int n_blocks = (n_steps + MQ_R - 1) / MQ_R;
    const float* gp = x + (size_t)T + col0;  // this thread's 4 cells of frame 1
    int to_issue = any_live ? n_steps : 0;   // frames this thread still has to request
    auto issue = [&](float4* slot) {  // one commit group per frame, also when nothing is copied
        if (to_issue > 0) {
            if (VEC) {
                cp_async16(slot, gp);
            } else {
                float* s4 = reinterpret_cast<float*>(slot);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (col0 + c < T) cp_async4(s4 + c, gp + c);
            }
        }
        cp_async_commit();
        --to_issue;
        gp += T;
    };
    float4* const my_ring = ring4 + tid;
    const int rs = TQ;
    for (int r = 0; r < PF; ++r) issue(my_ring + r * rs);
    __syncthreads();

    uint32_t word = 0;
    float bo[MQ_R];
    float4* slot = nullptr;
    // one frame: max(prev_log1, prev_log2) + attn (alignment.py:59; fmaxf == the reference's max for NaN-free input) and the
    // backtrack predicate of alignment.py:68 (tie → diagonal; column 0 never moves) for this thread's four cells
    auto frame = [&](int k, float left_lane0) {
        cp_async_wait<PF - 1>();
        const float4 xv = slot[k * rs];
        float left = __shfl_up_sync(0xffffffffu, p3, 1);
        if (lane == 0) left = left_lane0;
        if (left >= p0 && tid != 0) word |= 1u << (4 * k);
        if (p0 >= p1) word |= 2u << (4 * k);
        if (p1 >= p2) word |= 4u << (4 * k);
        if (p2 >= p3) word |= 8u << (4 * k);
        const float m0 = fmaxf(p0, left), m1 = fmaxf(p1, p0), m2 = fmaxf(p2, p1), m3 = fmaxf(p3, p2);
        p0 = __fadd_rn(xv.x, m0);
        p1 = __fadd_rn(xv.y, m1);
        p2 = __fadd_rn(xv.z, m2);
        p3 = __fadd_rn(xv.w, m3);
        bo[k] = p3;
        issue(slot + k * rs);  // the slot just consumed is refilled PF frames ahead
    };
    int b0 = 0, b1 = 1, b2 = 2;  // s % 3, (s + 1) % 3, (s + 2) % 3
    for (int s = 0; s < n_blocks + nwarps - 1; ++s) {
        const int blk = s - warp;
        if (blk >= 0 && blk < n_blocks && warp_live) {
            const int e0 = blk * MQ_R;
            float bl[MQ_R];  // lane 0's left neighbours: the previous frame's last column of warp w − 1, for each frame of the block
#pragma unroll
            for (int k = 0; k < MQ_R; ++k) bl[k] = NEG_INF;
            if (warp > 0) {
                bl[0] = bnd[b1][warp - 1][MQ_R - 1];
                const float4 a = *reinterpret_cast<const float4*>(&bnd[b2][warp - 1][0]);
                const float4 c = *reinterpret_cast<const float4*>(&bnd[b2][warp - 1][4]);
                bl[1] = a.x; bl[2] = a.y; bl[3] = a.z; bl[4] = a.w; bl[5] = c.x; bl[6] = c.y; bl[7] = c.z;
            }
            word = 0;
            slot = my_ring + (e0 % PF) * rs;  // PF is a multiple of the block: no wrap inside it
            if (e0 + MQ_R <= n_steps) {
#pragma unroll
                for (int k = 0; k < MQ_R; ++k) frame(k, bl[k]);
            } else {  // the last, partial block
#pragma unroll
                for (int k = 0; k < MQ_R; ++k) {
                    bo[k] = NEG_INF;
                    if (e0 + k < n_steps) frame(k, bl[k]);
                }
            }
            d[(size_t)blk * TQ + tid] = word;
            if (lane == 31) {
                *reinterpret_cast<float4*>(&bnd[b0][warp][0]) = make_float4(bo[0], bo[1], bo[2], bo[3]);
                *reinterpret_cast<float4*>(&bnd[b0][warp][4]) = make_float4(bo[4], bo[5], bo[6], bo[7]);
            }
        }
        __syncthreads();
        const int t0 = b0;
        b0 = b1; b1 = b2; b2 = t0;
    }
    cp_async_wait<0>();
    __syncthreads();

    if (warp == 0) {  // backtrack (alignment.py:62-73), 32 frames per step
        int j = n_text - 1;
        for (int top = n_mel - 1; top >= 1; top -= 32) {
            const int row = top - lane;
            const int jlow = j - 31;
            uint32_t win = 0;  // bit q ↔ the predicate of cell (row, jlow + q)
            if (row >= 1) {
                const int e = row - 1, sh = 4 * (e & 7);
                const uint32_t* dr = d + (size_t)(e >> 3) * TQ;
                const int tlo = jlow >> 2;  // floor also for negative jlow
                unsigned long long acc = 0;
#pragma unroll
                for (int q = 0; q < 9; ++q) {
                    const int t = tlo + q;
                    if (t >= 0 && 4 * t <= j) acc |= (unsigned long long)((dr[t] >> sh) & 0xFu) << (4 * t - jlow + 3);
                }
                win = (uint32_t)(acc >> 3);
            }
            int myj = -1, jj = 31;  // jj = j − jlow
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
                if (top - r < 1) break;  // warp-uniform
                const uint32_t h = __shfl_sync(0xffffffffu, win, r);
                if (lane == r) myj = jlow + jj;
                jj -= (int)((h >> jj) & 1u);
            }
            j = jlow + jj;
            if (row >= 1) p[row] = myj;
        }
        if (lane == 0) p[0] = j;  // opt[0, j] = 1   (alignment.py:73)
    }
    __syncthreads();
    // durations = attn_hard.sum over frames (variance_adaptor.py:267-268)
    for (int f = tid; f < n_mel; f += blockDim.x) atomicAdd(&dur[p[f]], 1);
}

// log of the soft alignment (the `torch.log(attn.data)` of variance_adaptor.py:168) as a fully parallel
// HBM-bound pass, so the sequential DP loop carries no transcendental
__global__ void __launch_bounds__(256)
mas_log_kernel(const float* __restrict__ x, long N, float* __restrict__ y) {
    pdl_prologue();
    const long N4 = N >> 2;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < N4; i += (long)gridDim.x * blockDim.x) {
        const float4 v = ld_stream(reinterpret_cast<const float4*>(x) + i);
        reinterpret_cast<float4*>(y)[i] = make_float4(logf(v.x), logf(v.y), logf(v.z), logf(v.w));
    }
    if (blockIdx.x == 0)
        for (long i = (N4 << 2) + threadIdx.x; i < N; i += blockDim.x) y[i] = logf(x[i]);
}

// dense 0/1 map [B,1,F,T] from the per-frame column index: HBM-write bound (4·B·F·T bytes)
__global__ void __launch_bounds__(256)
mas_dense_kernel(const int* __restrict__ path, int B, int F, int T, float* __restrict__ hard) {
    pdl_prologue();
    const long rows = (long)B * F;
    const int lane = threadIdx.x & 31;
    for (long row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows;
         row += (long)gridDim.x * (blockDim.x >> 5)) {
        const int j = path[row];
        float* o = hard + row * T;
        if ((T & 3) == 0) {
            for (int t = lane * 4; t < T; t += 128) {
                float4 v = make_float4(t == j, t + 1 == j, t + 2 == j, t + 3 == j);
                st_stream(reinterpret_cast<float4*>(o + t), v);
            }
        } else {
            for (int t = lane; t < T; t += 32) o[t] = (t == j) ? 1.f : 0.f;
        }
    }
}

template <int CHUNKS, int PF, bool LOG>
static cudaError_t launch_mas(int B, int threads, cudaStream_t s, const float* attn, const int* in_lens, const int* out_lens,
                              int F, int T, int W, uint32_t* dirs, int* path, int* durations) {
    const size_t smem = (size_t)PF * CHUNKS * threads * sizeof(float) + (size_t)MAS_DB * W * sizeof(uint32_t);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(mas_dp_kernel<CHUNKS, PF, LOG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    fs2k_launch_serial(mas_dp_kernel<CHUNKS, PF, LOG>, dim3(B), dim3(threads), smem, s, attn, in_lens, out_lens, F, T, W, dirs, path, durations);
    return cudaGetLastError();
}

}  // namespace fs2k

using namespace fs2k;

static int g_mas_wavefront = 1;
// 1 (default): four columns per thread, skewed blocks of 8 frames; 0: one column per thread, one __syncthreads per frame
// (the round-1 kernel, kept for A/B measurements and as a second implementation in the tests)
extern "C" int fs2k_mas_set_wavefront(int enabled) {
    g_mas_wavefront = enabled ? 1 : 0;
    return FS2K_OK;
}

static int mas_quad_threads(int T) { return ((T + 3) / 4 + 31) / 32 * 32; }
static size_t mas_quad_stride(int F, int T) { return (size_t)((F + MQ_R - 1) / MQ_R) * mas_quad_threads(T); }  // words per utterance
static size_t mas_dirs_bytes(int B, int F, int T) {
    size_t words = (size_t)F * ((T + 31) / 32);
    if (mas_quad_stride(F, T) > words) words = mas_quad_stride(F, T);
    const size_t n = (size_t)B * words * sizeof(uint32_t);
    return (n + 255) / 256 * 256;
}
// direction words + (for take_log) the log-probabilities [B,F,T]
extern "C" size_t fs2k_mas_workspace_bytes(int B, int F, int T) {
    return mas_dirs_bytes(B, F, T) + (size_t)B * F * T * sizeof(float);
}

template <int PF, bool VEC>
static cudaError_t launch_mas_quad(int B, int threads, cudaStream_t s, const float* attn, const int* in_lens, const int* out_lens, int F,
                                   int T, uint32_t* dirs, int* path, int* durations) {
    const size_t smem = (size_t)PF * threads * sizeof(float4);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(mas_dp_quad_kernel<PF, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    fs2k_launch_serial(mas_dp_quad_kernel<PF, VEC>, dim3(B), dim3(threads), smem, s, attn, in_lens, out_lens, F, T, dirs,
                       mas_quad_stride(F, T), path, durations);
    return cudaGetLastError();
}

extern "C" int fs2k_mas_fwd(const float* attn, int take_log, const int* in_lens, const int* out_lens, int B,
                            int F, int T, int* path, int* durations, float* hard, void* workspace,
                            size_t workspace_bytes, fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && F >= 0 && T >= 0, FS2K_ERR_BAD_SHAPE);
    if (B == 0 || F == 0 || T == 0) return FS2K_OK;
    FS2K_REQUIRE(T <= 4096, FS2K_ERR_UNSUPPORTED);
    FS2K_REQUIRE(workspace_bytes >= fs2k_mas_workspace_bytes(B, F, T), FS2K_ERR_WORKSPACE);
    FS2K_REQUIRE(attn && in_lens && out_lens && path && durations && workspace, FS2K_ERR_NULL);
    const int W = (T + 31) / 32;
    int threads = ((T < 1024 ? T : 1024) + 31) / 32 * 32;
    const int chunks = (T + threads - 1) / threads;
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t* dirs = (uint32_t*)workspace;
    if (take_log) {
        float* logbuf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + mas_dirs_bytes(B, F, T));
        const long n = (long)B * F * T;
        long g = (n / 4 + 255) / 256;
        if (g > 148 * 16) g = 148 * 16;
        if (g < 1) g = 1;
        fs2k_launch(mas_log_kernel, dim3((int)g), dim3(256), 0, s, attn, n, logbuf);
        FS2K_CHECK_LAUNCH();
        attn = logbuf;
    }
    cudaError_t e;
#define MAS_ARGS B, threads, s, attn, in_lens, out_lens, F, T, W, dirs, path, durations
    if (g_mas_wavefront) {
        const int qt = mas_quad_threads(T);
        const bool vec = (T % 4 == 0) && ((reinterpret_cast<uintptr_t>(attn) & 15) == 0);
#define MQ_ARGS B, qt, s, attn, in_lens, out_lens, F, T, dirs, path, durations
        // ring depth: ≈ 1 µs of frames ahead of the consumer, within the shared memory of one SM
        if (qt <= 256) e = vec ? launch_mas_quad<32, true>(MQ_ARGS) : launch_mas_quad<32, false>(MQ_ARGS);
        else if (qt <= 512) e = vec ? launch_mas_quad<16, true>(MQ_ARGS) : launch_mas_quad<16, false>(MQ_ARGS);
        else e = vec ? launch_mas_quad<8, true>(MQ_ARGS) : launch_mas_quad<8, false>(MQ_ARGS);
#undef MQ_ARGS
    } else if (chunks == 1) e = launch_mas<1, 16, false>(MAS_ARGS);
    else if (chunks == 2) e = launch_mas<2, 16, false>(MAS_ARGS);
    else { threads = 1024; e = launch_mas<4, 8, false>(MAS_ARGS); }
#undef MAS_ARGS
    if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    if (hard) {
        const long rows = (long)B * F;
        int grid = (int)((rows + 7) / 8);
        if (grid > 148 * 16) grid = 148 * 16;
        fs2k_launch(mas_dense_kernel, dim3(grid), dim3(256), 0, s, path, B, F, T, hard);
        FS2K_CHECK_LAUNCH();
    }
    return FS2K_OK;
}
