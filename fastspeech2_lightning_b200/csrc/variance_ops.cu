// Memory-bound pieces of the variance adaptor and the input embedding.
//   bucketize + embedding gather + add   fs2/variance_adaptor.py:183-205, :322, :343
//   average_variance                     fs2/variance_adaptor.py:207-222
//   inference duration rounding          fs2/variance_adaptor.py:359-366
//   token embedding + positional term    fs2/model.py:183-190, fs2/layers.py:132-140
//   broadcast add of per-utterance rows  fs2/model.py:196-213 (GST / speaker / language)
#include "common.cuh"

namespace fs2k {

// torch.bucketize(v, bins, right=False): id = first i with bins[i] >= v; NaN compares false → n_bins.
__device__ __forceinline__ int bucket_of(float v, const float* __restrict__ bins, int n) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (!(bins[mid] >= v)) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// one warp per position; lanes stride the D channels as float4
__global__ void __launch_bounds__(256)
bucketize_embed_add_kernel(const float* __restrict__ v_in,   // [N] value to bucketize
                           float scale,                      // inference: prediction *= control
                           float* __restrict__ v_scaled,     // [N] or null: v·scale written back
                           const float* __restrict__ bins, int n_bins,
                           const float* __restrict__ table,  // [n_bins+1, D]
                           const float* __restrict__ x,      // [N,D]
                           float* __restrict__ y,            // [N,D] = x + table[id]   (may alias x)
                           long long* __restrict__ ids,      // [N] int64 or null
                           long N, int D) {
    pdl_prologue();
    extern __shared__ float s_bins[];
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x) s_bins[i] = bins[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int D4 = D >> 2;
    for (long n = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); n < N;
         n += (long)gridDim.x * (blockDim.x >> 5)) {
        float v = v_in[n];
        if (scale != 1.0f) v = __fmul_rn(v, scale);
        const int id = bucket_of(v, s_bins, n_bins);
        if (lane == 0) {
            if (ids) ids[n] = id;
            if (v_scaled) v_scaled[n] = v;
        }
        const float4* e = reinterpret_cast<const float4*>(table + (size_t)id * D);
        const float4* xi = reinterpret_cast<const float4*>(x + (size_t)n * D);
        float4* yo = reinterpret_cast<float4*>(y + (size_t)n * D);
        for (int q = lane; q < D4; q += 32) {
            float4 a = xi[q], t = __ldg(e + q);
            a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
            yo[q] = a;
        }
    }
}

__global__ void bucketize_kernel(const float* __restrict__ v, const float* __restrict__ bins, int n_bins,
                                 long long* __restrict__ ids, long N) {
    pdl_prologue();
    long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N) ids[n] = bucket_of(v[n], bins, n_bins);
}

// average_variance: one thread per phone, direct fp32 sum over the phone's frame span (mean of the
// non-zero frames, 0 when there are none).  The reference takes differences of fp32 prefix sums,
// which is order dependent; the direct sum is the more accurate of the two and agrees
// to a few ulp of the running sum.
__global__ void average_variance_kernel(const float* __restrict__ var,  // [B,F]
                                        const int* __restrict__ cum,    // [B,T] inclusive cumsum of durations
                                        int B, int F, int T, float* __restrict__ out) {
    pdl_prologue();
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * T) return;
    const int b = (int)(i / T), t = (int)(i % T);
    const int end = min(cum[i], F);
    const int start = min(t > 0 ? cum[i - 1] : 0, F);
    float s = 0.f;
    int cnt = 0;
    for (int f = start; f < end; ++f) {
        const float v = var[(size_t)b * F + f];
        s += v;
        cnt += (v != 0.0f);
    }
    out[i] = cnt == 0 ? 0.f : s / (float)cnt;
}

// dur = int(clamp(rint(exp(logd) − 1) · control, min 0))   (round half to even, truncating cast)
__global__ void round_durations_kernel(const float* __restrict__ log_dur, const uint8_t* __restrict__ mask,
                                       float control, long N, int* __restrict__ dur) {
    pdl_prologue();
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float d = rintf(__fadd_rn(expf(log_dur[i]), -1.0f));
    d = fmaxf(__fmul_rn(d, control), 0.0f);
    // float → int32 like torch's .int() (C cast; saturate instead of UB for huge values)
    int r = d >= 2147483520.f ? 2147483647 : (int)d;
    if (d != d) r = 0;
    (void)mask;
    dur[i] = r;
}

// x0 = table[text] ; x = x0 + posenc(t)·[t < len]      one warp per (b,t)
__global__ void __launch_bounds__(256)
embed_posenc_kernel(const int* __restrict__ text, const float* __restrict__ table, int n_sym,
                    const float* __restrict__ inv_freq, const int* __restrict__ lens, int B, int T, int D,
                    float* __restrict__ emb, float* __restrict__ x, int* __restrict__ err) {
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    const int D4 = D >> 2, half = D >> 1;
    const long N = (long)B * T;
    for (long n = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); n < N;
         n += (long)gridDim.x * (blockDim.x >> 5)) {
        const int b = (int)(n / T), t = (int)(n % T);
        int id = text[n];
        if (id < 0 || id >= n_sym) {
            if (lane == 0 && err) atomicExch(err, 1);
            id = 0;
        }
        const bool valid = t < lens[b];
        const float4* e = reinterpret_cast<const float4*>(table + (size_t)id * D);
        for (int q = lane; q < D4; q += 32) {
            float4 v = __ldg(e + q);
            if (emb) reinterpret_cast<float4*>(emb + (size_t)n * D)[q] = v;
            if (valid) {
                float p[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int ch = q * 4 + k;
                    const float arg = __fmul_rn((float)t, inv_freq[ch < half ? ch : ch - half]);
                    p[k] = ch < half ? sinf(arg) : cosf(arg);
                }
                v.x += p[0]; v.y += p[1]; v.z += p[2]; v.w += p[3];
            }
            reinterpret_cast<float4*>(x + (size_t)n * D)[q] = v;
        }
    }
}

// x[b,l,:] (+)= pos(l)·mask — for inputs that are not an embedding lookup (phonological features)
__global__ void __launch_bounds__(256)
add_posenc_kernel(const float* __restrict__ x_in, const float* __restrict__ inv_freq, const int* __restrict__ lens,
                  int B, int L, int D, float* __restrict__ x) {
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    const int half = D >> 1;
    const long N = (long)B * L;
    for (long n = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); n < N;
         n += (long)gridDim.x * (blockDim.x >> 5)) {
        const int b = (int)(n / L), l = (int)(n % L);
        const bool valid = l < lens[b];
        for (int ch = lane; ch < D; ch += 32) {
            float v = x_in[(size_t)n * D + ch];
            if (valid) {
                const float arg = __fmul_rn((float)l, inv_freq[ch < half ? ch : ch - half]);
                v += ch < half ? sinf(arg) : cosf(arg);
            }
            x[(size_t)n * D + ch] = v;
        }
    }
}

// y[b,l,:] = x[b,l,:] + Σ_k rows_k[id_k[b], :]   (k ≤ 3: style / speaker / language rows)
__global__ void __launch_bounds__(256)
add_rows_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int L, int D,
                const float* r0, const int* i0, const float* r1, const int* i1, const float* r2, const int* i2) {
    pdl_prologue();
    const long N = (long)B * L * (D >> 2);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long)gridDim.x * blockDim.x) {
        const int q = (int)(i % (D >> 2));
        const int b = (int)(i / ((long)L * (D >> 2)));
        float4 v = reinterpret_cast<const float4*>(x)[i];
        const float* rs[3] = {r0, r1, r2};
        const int* is[3] = {i0, i1, i2};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (!rs[k]) continue;
            const int row = is[k] ? is[k][b] : b;
            const float4 a = __ldg(reinterpret_cast<const float4*>(rs[k] + (size_t)row * D) + q);
            v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
        }
        reinterpret_cast<float4*>(y)[i] = v;
    }
}

// lens[b] = #{mask[b,:] != 0} ; one warp per utterance (fs2/model.py:226-230)
__global__ void mask_lens_kernel(const uint8_t* __restrict__ mask, int B, int L, int* __restrict__ lens) {
    pdl_prologue();
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    int c = 0;
    for (int l = lane; l < L; l += 32) c += mask[(size_t)b * L + l] != 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) lens[b] = c;
}

}  // namespace fs2k

using namespace fs2k;

static inline int grid_for_warps(long n_rows, int warps_per_cta) {
    long g = (n_rows + warps_per_cta - 1) / warps_per_cta;
    const long cap = 148L * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

extern "C" int fs2k_bucketize_embed_add(const float* v, float scale, float* v_scaled, const float* bins, int n_bins,
                                        const float* table, const float* x, float* y, long long* ids, long N, int D,
                                        fs2k_stream_t stream) {
    FS2K_REQUIRE(N >= 0 && D > 0 && n_bins >= 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE((D & 3) == 0 && n_bins <= 8192, FS2K_ERR_UNSUPPORTED);
    if (N == 0) return FS2K_OK;
    FS2K_REQUIRE(v && bins && table && x && y, FS2K_ERR_NULL);
    fs2k_launch(bucketize_embed_add_kernel, dim3(grid_for_warps(N, 8)), dim3(256), n_bins * sizeof(float), (cudaStream_t)stream, v, scale, v_scaled, bins, n_bins, table, x, y, ids, N, D);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_bucketize(const float* v, const float* bins, int n_bins, long long* ids, long N,
                              fs2k_stream_t stream) {
    FS2K_REQUIRE(N >= 0 && n_bins >= 0, FS2K_ERR_BAD_SHAPE);
    if (N == 0) return FS2K_OK;
    FS2K_REQUIRE(v && bins && ids, FS2K_ERR_NULL);
    fs2k_launch(bucketize_kernel, dim3(cdiv(N, 256)), dim3(256), 0, (cudaStream_t)stream, v, bins, n_bins, ids, N);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_average_variance(const float* var, const int* cum, int B, int F, int T, float* out,
                                     fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && F >= 0 && T >= 0, FS2K_ERR_BAD_SHAPE);
    if ((long)B * T == 0) return FS2K_OK;
    FS2K_REQUIRE(var && cum && out, FS2K_ERR_NULL);
    fs2k_launch(average_variance_kernel, dim3(cdiv((long)B * T, 128)), dim3(128), 0, (cudaStream_t)stream, var, cum, B, F, T, out);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_round_durations(const float* log_dur, float control, long N, int* dur, fs2k_stream_t stream) {
    FS2K_REQUIRE(N >= 0, FS2K_ERR_BAD_SHAPE);
    if (N == 0) return FS2K_OK;
    FS2K_REQUIRE(log_dur && dur, FS2K_ERR_NULL);
    fs2k_launch(round_durations_kernel, dim3(cdiv(N, 256)), dim3(256), 0, (cudaStream_t)stream, log_dur, nullptr, control, N, dur);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_embed_posenc(const int* text, const float* table, int n_sym, const float* inv_freq,
                                 const int* lens, int B, int T, int D, float* emb, float* x, int* err_flag,
                                 fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && T >= 0 && D > 0 && n_sym > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE((D & 3) == 0, FS2K_ERR_UNSUPPORTED);
    if ((long)B * T == 0) return FS2K_OK;
    FS2K_REQUIRE(text && table && inv_freq && lens && x, FS2K_ERR_NULL);
    fs2k_launch(embed_posenc_kernel, dim3(grid_for_warps((long)B * T, 8)), dim3(256), 0, (cudaStream_t)stream, text, table, n_sym, inv_freq, lens, B, T, D, emb, x, err_flag);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_add_posenc(const float* x_in, const float* inv_freq, const int* lens, int B, int L, int D,
                               float* x, fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && L >= 0 && D > 0, FS2K_ERR_BAD_SHAPE);
    if ((long)B * L == 0) return FS2K_OK;
    FS2K_REQUIRE(x_in && inv_freq && lens && x, FS2K_ERR_NULL);
    fs2k_launch(add_posenc_kernel, dim3(grid_for_warps((long)B * L, 8)), dim3(256), 0, (cudaStream_t)stream, x_in, inv_freq, lens, B, L, D, x);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_add_rows(const float* x, float* y, int B, int L, int D, const float* rows0, const int* ids0,
                             const float* rows1, const int* ids1, const float* rows2, const int* ids2,
                             fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && L >= 0 && D > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE((D & 3) == 0, FS2K_ERR_UNSUPPORTED);
    const long N = (long)B * L * (D >> 2);
    if (N == 0) return FS2K_OK;
    FS2K_REQUIRE(x && y, FS2K_ERR_NULL);
    long g = (N + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    fs2k_launch(add_rows_kernel, dim3((int)g), dim3(256), 0, (cudaStream_t)stream, x, y, B, L, D, rows0, ids0, rows1, ids1, rows2, ids2);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_mask_lens(const uint8_t* mask, int B, int L, int* lens, fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && L >= 0, FS2K_ERR_BAD_SHAPE);
    if (B == 0) return FS2K_OK;
    FS2K_REQUIRE(mask && lens, FS2K_ERR_NULL);
    fs2k_launch(mask_lens_kernel, dim3(cdiv(B, 4)), dim3(128), 0, (cudaStream_t)stream, mask, B, L, lens);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

namespace fs2k {
// mask[b,l] = l < lens[b]   (fs2/utils/heavy.py:11-15)
__global__ void lens_mask_kernel(const int* __restrict__ lens, int B, int L, uint8_t* __restrict__ mask) {
    pdl_prologue();
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (long)B * L) mask[i] = (int)(i % L) < lens[i / L];
}
}  // namespace fs2k

extern "C" int fs2k_lens_mask(const int* lens, int B, int L, uint8_t* mask, fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && L >= 0, FS2K_ERR_BAD_SHAPE);
    if ((long)B * L == 0) return FS2K_OK;
    FS2K_REQUIRE(lens && mask, FS2K_ERR_NULL);
    fs2k_launch(fs2k::lens_mask_kernel, dim3(cdiv((long)B * L, 256)), dim3(256), 0, (cudaStream_t)stream, lens, B, L, mask);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}
