// Learned-alignment scores: the tail of ConvAttention.forward, fs2/attn/attention.py:238-251.
//
//   dist[b,f,t]   = −0.0005 · Σ_c (q[b,f,c] − k[b,t,c])²                        (:239-241)
//   logprob       = log_softmax_t(dist) + log(prior + 1e-8)                     (:242-245)
//   soft          = softmax_t(logprob with −inf on padded keys)                 (:247-250)
// The reference materialises the [B,80,F,T] difference tensor; here the 80-channel reduction
// happens in registers (direct (q−k)² form — bit-compatible ordering is not required, the fp32
// tolerance is) and one warp then normalises each (b,f) row out of shared memory, so HBM sees
// only the prior (read) and the two outputs (write) plus one round trip of the raw scores.
#include "common.cuh"

namespace fs2k {

constexpr int DT = 64;  // tile edge (frames × phones)

template <int CCH>
__global__ void __launch_bounds__(256)
aligner_dist_kernel(const float* __restrict__ q,  // [B,F,C]
                    const float* __restrict__ k,  // [B,T,C]
                    int F, int T, int C, float coef, float* __restrict__ dist /* [B,F,T] */) {
    pdl_prologue();
    __shared__ __align__(16) float Qs[DT][CCH + 4];
    __shared__ __align__(16) float Ks[DT][CCH + 4];
    const int b = blockIdx.z, f0 = blockIdx.y * DT, t0 = blockIdx.x * DT;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int C4 = C >> 2;
    for (int i = tid; i < DT * C4; i += 256) {
        const int r = i / C4, c = i % C4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f), u = v;
        if (f0 + r < F) v = *reinterpret_cast<const float4*>(q + ((size_t)b * F + f0 + r) * C + c * 4);
        if (t0 + r < T) u = *reinterpret_cast<const float4*>(k + ((size_t)b * T + t0 + r) * C + c * 4);
        *reinterpret_cast<float4*>(&Qs[r][c * 4]) = v;
        *reinterpret_cast<float4*>(&Ks[r][c * 4]) = u;
    }
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int c = 0; c < C4; ++c) {
        float4 q4[4], k4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) q4[i] = *reinterpret_cast<const float4*>(&Qs[ty * 4 + i][c * 4]);
#pragma unroll
        for (int j = 0; j < 4; ++j) k4[j] = *reinterpret_cast<const float4*>(&Ks[tx + 16 * j][c * 4]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float d;
                d = q4[i].x - k4[j].x; acc[i][j] = fmaf(d, d, acc[i][j]);
                d = q4[i].y - k4[j].y; acc[i][j] = fmaf(d, d, acc[i][j]);
                d = q4[i].z - k4[j].z; acc[i][j] = fmaf(d, d, acc[i][j]);
                d = q4[i].w - k4[j].w; acc[i][j] = fmaf(d, d, acc[i][j]);
            }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int f = f0 + ty * 4 + i;
        if (f >= F) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int t = t0 + tx + 16 * j;
            if (t < T) dist[((size_t)b * F + f) * T + t] = coef * acc[i][j];
        }
    }
}

// one warp per (b,f) row; the row is cached in shared memory (T floats per warp)
__global__ void __launch_bounds__(128)
aligner_softmax_kernel(const float* dist,   // [B,F,T] raw scores (may alias `soft`)
                       const float* __restrict__ prior,  // [B,F,T] or null
                       const int* __restrict__ key_lens, // [B] or null (no key mask)
                       int B, int F, int T, float* __restrict__ logprob, float* soft) {
    pdl_prologue();
    extern __shared__ float rows[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* r = rows + (size_t)warp * T;
    const long n_rows = (long)B * F;
    for (long row = (long)blockIdx.x * 4 + warp; row < n_rows; row += (long)gridDim.x * 4) {
        const int b = (int)(row / F);
        const int klen = key_lens ? min(key_lens[b], T) : T;
        const float* d = dist + row * T;
        float mx = -INFINITY;
        for (int t = lane; t < T; t += 32) {
            const float v = d[t];
            r[t] = v;
            mx = fmaxf(mx, v);
        }
        mx = warp_max(mx);
        float mx2 = -INFINITY;
        if (prior) {
            float s = 0.f;
            for (int t = lane; t < T; t += 32) s += expf(r[t] - mx);
            const float lse = mx + logf(warp_sum(s));
            const float* p = prior + row * T;
            for (int t = lane; t < T; t += 32) {
                const float v = (r[t] - lse) + logf(p[t] + 1e-8f);
                r[t] = v;
                logprob[row * T + t] = v;
                if (t < klen) mx2 = fmaxf(mx2, v);
            }
        } else {
            for (int t = lane; t < T; t += 32) {
                logprob[row * T + t] = r[t];
                if (t < klen) mx2 = fmaxf(mx2, r[t]);
            }
        }
        mx2 = warp_max(mx2);
        float s2 = 0.f;
        for (int t = lane; t < klen; t += 32) s2 += expf(r[t] - mx2);
        const float inv = 1.0f / warp_sum(s2);
        for (int t = lane; t < T; t += 32) soft[row * T + t] = t < klen ? expf(r[t] - mx2) * inv : 0.f;
        __syncwarp();
    }
}

}  // namespace fs2k

using namespace fs2k;

extern "C" int fs2k_aligner_fwd(const float* q, const float* k, const float* prior, const int* key_lens, int B, int F,
                                int T, int C, float* logprob, float* soft, fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && F >= 0 && T >= 0 && C > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE((C & 3) == 0 && C <= 80 && T <= 8192, FS2K_ERR_UNSUPPORTED);  // n_att_channels = 80 (attention.py:105)
    if (B == 0 || F == 0 || T == 0) return FS2K_OK;
    FS2K_REQUIRE(q && k && logprob && soft, FS2K_ERR_NULL);
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid(cdiv(T, DT), cdiv(F, DT), B);
    // raw scores are staged in `soft` (overwritten by the normalisation pass, row by row, after it is read)
    fs2k_launch(aligner_dist_kernel<80>, dim3(grid), dim3(256), 0, s, q, k, F, T, C, -0.0005f, soft);
    FS2K_CHECK_LAUNCH();
    const int smem = 4 * T * (int)sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(aligner_softmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    }
    long g = ((long)B * F + 3) / 4;
    if (g > 148 * 16) g = 148 * 16;
    fs2k_launch(aligner_softmax_kernel, dim3((int)g), dim3(128), smem, s, soft, prior, key_lens, B, F, T, logprob, soft);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}
