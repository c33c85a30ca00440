// LengthRegulator: duration expansion as a cumsum scan plus a vectorised, coalesced gather.
//
// Replaces LengthRegulator.forward, fs2/variance_adaptor.py:65-81 (per-item repeat_interleave +
// pad_sequence + a mask built on the CPU, ≥ B+2 host syncs).  The implied index is
//     idx[b,f] = #{ t : cumsum(dur[b])[t] <= f }            (bit-exact integer contract)
//     out[b,f,:] = f < total[b] ? x[b, idx[b,f], :] : 0 ;   mask[b,f] = f < total[b]
// with `total[b] = Σ_t dur[b,t]` the un-truncated length (variance_adaptor.py:74-77).  The output
// width `min(max_b total, max_length)` is a host decision (the caller passes F_out).
//
// The decoder's positional term (`+ pos_emb · tgt_mask`, fs2/model.py:233-241) can be fused into
// the same pass (out_pos), since the row is in registers anyway.
#include "common.cuh"

namespace fs2k {

// arg − 2π·rint(arg / 2π) for 0 ≤ arg < 5e4, to ≈ 1e-7: 2π = C1 + C2 + C3 with C1 (9 significant bits) and C2 chosen so
// that k·C1 and k·C2 are exact in fp32 for k < 2^13
__device__ __forceinline__ float posenc_reduce(float arg) {
    const float k = rintf(arg * 0.15915494309189535f);
    float r = fmaf(k, -6.28125f, arg);
    r = fmaf(k, -1.9350051879882812e-3f, r);
    r = fmaf(k, -3.0199160505556543e-7f, r);
    return r;
}

// ---- scan: one warp per utterance, inclusive prefix sum of durations ----
__global__ void __launch_bounds__(128)
lr_scan_kernel(const int* __restrict__ dur, int B, int T, int* __restrict__ cum, int* __restrict__ total) {
    pdl_prologue();
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    int carry = 0;
    for (int t0 = 0; t0 < T; t0 += 32) {
        const int t = t0 + lane;
        int v = t < T ? max(dur[(size_t)b * T + t], 0) : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += n;
        }
        v += carry;
        if (t < T) cum[(size_t)b * T + t] = v;
        carry = __shfl_sync(0xffffffffu, v, 31);
    }
    if (lane == 0) total[b] = carry;
}

constexpr int kLrTile = 16;  // frames per CTA: 2 rows per warp (64 left one CTA per SM at B = 16, F = 500 — a serial chain of 8 dependent row copies per warp)
constexpr int kLrMaxSmemT = 2048;

// ---- gather: CTA = kLrTile consecutive frames of one utterance ----
template <bool POS>
__global__ void __launch_bounds__(256)
lr_gather_kernel(const float* __restrict__ x,    // [B,T,D]
                 const int* __restrict__ cum,    // [B,T] inclusive cumsum
                 const int* __restrict__ total,  // [B]
                 int T, int D, int F_out,
                 float* __restrict__ out,        // [B,F_out,D] or null
                 float* __restrict__ out_pos,    // [B,F_out,D] = out + posenc(f)·mask, or null
                 const float* __restrict__ inv_freq,  // [D/2]
                 uint8_t* __restrict__ mask,     // [B,F_out] or null
                 int* __restrict__ idx_out)      // [B,F_out] or null (−1 on padding)
{
    pdl_prologue();
    __shared__ int s_cum[kLrMaxSmemT];
    __shared__ int s_idx[kLrTile];
    const int b = blockIdx.y;
    const int f0 = blockIdx.x * kLrTile;
    const int tid = threadIdx.x;
    const int* c = cum + (size_t)b * T;
    const bool in_smem = T <= kLrMaxSmemT;
    if (in_smem)
        for (int t = tid; t < T; t += blockDim.x) s_cum[t] = c[t];
    __syncthreads();
    const int tot = total[b];
    if (tid < kLrTile) {
        const int f = f0 + tid;
        int idx = -1;
        if (f < F_out && f < tot) {
            // upper bound: first t with cum[t] > f  ==  #{t : cum[t] <= f}
            int lo = 0, hi = T;
            const int* cc = in_smem ? s_cum : c;
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (cc[mid] <= f) lo = mid + 1; else hi = mid;
            }
            idx = lo;
        }
        s_idx[tid] = idx;
        if (f < F_out) {
            if (mask) mask[(size_t)b * F_out + f] = f < tot;
            if (idx_out) idx_out[(size_t)b * F_out + f] = idx;
        }
    }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    const int D4 = D >> 2;
    for (int r = warp; r < kLrTile; r += 8) {
        const int f = f0 + r;
        if (f >= F_out) break;
        const int idx = s_idx[r];
        const float4* src = reinterpret_cast<const float4*>(x + ((size_t)b * T + max(idx, 0)) * D);
        float4* dst = out ? reinterpret_cast<float4*>(out + ((size_t)b * F_out + f) * D) : nullptr;
        float4* dstp = POS ? reinterpret_cast<float4*>(out_pos + ((size_t)b * F_out + f) * D) : nullptr;
        for (int q = lane; q < D4; q += 32) {
            float4 v = idx >= 0 ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (dst) st_stream(dst + q, v);
            if (POS) {
                if (idx >= 0) {
                    // PositionalEmbedding (fs2/layers.py:132-140): [sin(f·ω_i) | cos(f·ω_i)], i < D/2.  The reference rounds the
                    // product f·ω to fp32 first (a matmul of fp32 vectors) and takes sin / cos of THAT number — so does this:
                    // no angle-addition recurrence.  sin / cos of the rounded argument: three-term Cody-Waite reduction by 2π
                    // (exact for the ≤ 2^13 periods of positions up to 50 000) + the MUFU sine / cosine on [−π, π]
                    // (abs error 4e-7): ≈ 10 instructions per channel instead of libdevice's ≈ 40.
                    const int half = D >> 1;
                    float e[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int ch = q * 4 + k;
                        const float arg = __fmul_rn((float)f, inv_freq[ch < half ? ch : ch - half]);
                        const float r = posenc_reduce(arg);
                        e[k] = ch < half ? __sinf(r) : __cosf(r);
                    }
                    v.x += e[0]; v.y += e[1]; v.z += e[2]; v.w += e[3];
                }
                st_stream(dstp + q, v);
            }
        }
    }
}

}  // namespace fs2k

extern "C" int fs2k_lr_scan(const int* durations, int B, int T, int* cum, int* total, fs2k_stream_t stream) {
    using namespace fs2k;
    FS2K_REQUIRE(B >= 0 && T >= 0, FS2K_ERR_BAD_SHAPE);
    if (B == 0) return FS2K_OK;
    FS2K_REQUIRE(durations && cum && total, FS2K_ERR_NULL);
    fs2k_launch(lr_scan_kernel, dim3(cdiv(B, 4)), dim3(128), 0, (cudaStream_t)stream, durations, B, T, cum, total);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_lr_gather(const float* x, const int* cum, const int* total, int B, int T, int D, int F_out,
                              float* out, float* out_pos, const float* inv_freq, uint8_t* mask, int* idx_out,
                              fs2k_stream_t stream) {
    using namespace fs2k;
    FS2K_REQUIRE(B >= 0 && T >= 0 && D > 0 && F_out >= 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE((D & 3) == 0, FS2K_ERR_UNSUPPORTED);
    if (B == 0 || F_out == 0) return FS2K_OK;
    FS2K_REQUIRE(x && cum && total, FS2K_ERR_NULL);
    FS2K_REQUIRE(!out_pos || inv_freq, FS2K_ERR_NULL);
    dim3 grid(cdiv(F_out, kLrTile), B);
    if (out_pos)
        fs2k_launch(lr_gather_kernel<true>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, x, cum, total, T, D, F_out, out, out_pos, inv_freq, mask, idx_out);
    else
        fs2k_launch(lr_gather_kernel<false>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, x, cum, total, T, D, F_out, out, nullptr, nullptr, mask, idx_out);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}
