// Forward-sum (CTC) alignment loss and its gradient — reference fs2/attn/attention_loss.py:22-62.
//
// The reference prepends a blank column (log-prob −1) to attn_logprob [B,1,F,T], masks the keys past
// key_len with −1e15, applies log_softmax over the T+1 classes and calls nn.CTCLoss(blank=0, mean,
// zero_infinity) with the targets 1..key_len and input length query_len.  Everything is fused here:
//
//   lse[b,t]      = logsumexp(blank, logit[b,t,0..K))                       (masked keys contribute exp(−1e15−max) = 0)
//   lp[t, class]  = value − lse[b,t]
//   α_t(s)        = lp[t, l'_s] + logsumexp(α_{t−1}(s), α_{t−1}(s−1), α_{t−1}(s−2) [s odd ≥ 3])   over the
//                   2K+1 states  blank,1,blank,2,…,K,blank   (labels are distinct, so the skip is always allowed)
//   nll_b         = −logsumexp(α_{Q−1}(2K), α_{Q−1}(2K−1));   loss = mean_b( nll_b / max(K,1) ), inf → 0
//
// One CTA per utterance; the recursion is sequential over the Q ≤ F frames (one __syncthreads per frame,
// α ping-pongs in shared memory and is written to HBM for the backward).  The recursion runs in fp64: |α| grows
// by ≈ log(K) per frame, and at a few thousand the fp32 spacing (1e-4) would put ≈ 1e-3 of noise on every posterior
// (torch's fp32 kernel has exactly that noise); the work is latency-bound and tiny, so fp64 costs nothing visible.  The backward runs the mirrored
// β recursion and emits, frame by frame, the gradient with respect to the *raw* attn_logprob:
//   d logit[t,k] = gout/(B·max(K,1)) · ( softmax[t,k+1] − exp(α_t(2k+1) + β_t(2k+1) − lp[t,k+1] + nll) )
// which is torch's ctc_loss gradient pushed through log_softmax and the blank padding (blank column dropped).
#include "common.cuh"

namespace fs2k {

constexpr int CTC_THREADS = 512;
#define CTC_NEG_INF (-INFINITY)

__device__ __forceinline__ double lse3(double a, double b, double c) {
    const double m = fmax(a, fmax(b, c));
    if (m == (double)CTC_NEG_INF) return (double)CTC_NEG_INF;
    return m + log(exp(a - m) + exp(b - m) + exp(c - m));
}
__device__ __forceinline__ double lse2(double a, double b) {
    const double m = fmax(a, b);
    if (m == (double)CTC_NEG_INF) return (double)CTC_NEG_INF;
    return m + log(exp(a - m) + exp(b - m));
}

// dynamic smem: 2·S_max doubles (state ping-pong) + F doubles (lse of this utterance)
__global__ void __launch_bounds__(CTC_THREADS)
ctc_alpha_kernel(const float* __restrict__ logit, const int* __restrict__ key_lens, const int* __restrict__ query_lens,
                 int F, int T, float blank, double* __restrict__ lse_out, double* __restrict__ log_alpha,
                 double* __restrict__ nll_out) {
    pdl_wait();  // launched with fs2k_launch_serial (one CTA per SM matters here) and never triggers its dependents early
    extern __shared__ double sm[];
    const int b = blockIdx.x;
    const int K = min(max(key_lens[b], 0), T), Q = min(max(query_lens[b], 0), F);
    const int S = 2 * K + 1, S_max = 2 * T + 1;
    double* a0 = sm;
    double* a1 = sm + S_max;
    double* lse = sm + 2 * S_max;
    const float* x = logit + (size_t)b * F * T;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = CTC_THREADS / 32;

    // phase 1: row-wise logsumexp over {blank, K keys}
    for (int t = warp; t < Q; t += n_warps) {
        const float* row = x + (size_t)t * T;
        float m = blank;
        for (int k = lane; k < K; k += 32) m = fmaxf(m, row[k]);
        m = warp_max(m);
        double s = lane == 0 ? exp((double)blank - (double)m) : 0.0;
        for (int k = lane; k < K; k += 32) s += exp((double)row[k] - (double)m);
        s = warp_sum_d(s);
        if (lane == 0) {
            const double v = (double)m + log(s);
            lse[t] = v;
            lse_out[(size_t)b * F + t] = v;
        }
    }
    __syncthreads();
    if (Q == 0) {
        if (threadIdx.x == 0) nll_out[b] = (double)INFINITY;
        return;
    }

    // phase 2: α recursion
    double* la = log_alpha + (size_t)b * F * S_max;
    for (int s = threadIdx.x; s < S; s += CTC_THREADS) {
        double v = (double)CTC_NEG_INF;
        if (s == 0) v = (double)blank - lse[0];
        else if (s == 1) v = (double)x[0] - lse[0];
        a0[s] = v;
        la[s] = v;
    }
    __syncthreads();
    double* prev = a0;
    double* cur = a1;
    for (int t = 1; t < Q; ++t) {
        const float* row = x + (size_t)t * T;
        const double l = lse[t];
        for (int s = threadIdx.x; s < S; s += CTC_THREADS) {
            const double e = (double)((s & 1) ? row[s >> 1] : blank) - l;
            const double p0 = prev[s];
            const double p1 = s >= 1 ? prev[s - 1] : (double)CTC_NEG_INF;
            const double p2 = ((s & 1) && s >= 3) ? prev[s - 2] : (double)CTC_NEG_INF;
            const double v = lse3(p0, p1, p2) + e;
            cur[s] = v;
            la[(size_t)t * S_max + s] = v;
        }
        __syncthreads();
        double* tmp = prev; prev = cur; cur = tmp;
    }
    if (threadIdx.x == 0) {
        const double end = S >= 2 ? lse2(prev[S - 1], prev[S - 2]) : prev[0];
        nll_out[b] = -end;
    }
}

// loss = mean_b( isinf(nll) ? 0 : nll / max(K,1) ), summed in a fixed order
__global__ void ctc_finalize_kernel(const double* __restrict__ nll, const int* __restrict__ key_lens, int B, int T,
                                    float* __restrict__ loss) {
    pdl_prologue();
    double acc = 0.0;
    for (int b = 0; b < B; ++b) {
        const double v = nll[b];
        const int K = min(max(key_lens[b], 0), T);
        if (!isinf(v)) acc += v / (double)max(K, 1);
    }
    loss[0] = (float)(acc / (double)max(B, 1));
}

__global__ void __launch_bounds__(CTC_THREADS)
ctc_beta_grad_kernel(const float* __restrict__ logit, const double* __restrict__ lse_g, const double* __restrict__ log_alpha,
                     const double* __restrict__ nll_g, const int* __restrict__ key_lens, const int* __restrict__ query_lens,
                     const float* __restrict__ gout, int B, int F, int T, float blank, float* __restrict__ dlogit) {
    pdl_wait();  // launched with fs2k_launch_serial (one CTA per SM matters here) and never triggers its dependents early
    extern __shared__ double sm[];
    const int b = blockIdx.x;
    const int K = min(max(key_lens[b], 0), T), Q = min(max(query_lens[b], 0), F);
    const int S = 2 * K + 1, S_max = 2 * T + 1;
    double* b0 = sm;
    double* b1 = sm + S_max;
    const float* x = logit + (size_t)b * F * T;
    float* dx = dlogit + (size_t)b * F * T;
    const double nll = nll_g[b];
    const bool dead = isinf(nll) || Q == 0;  // zero_infinity: no gradient from an impossible alignment
    const double gr = (double)gout[0] / ((double)B * (double)max(K, 1));

    // rows past the utterance (and every row of a dead utterance) get zero gradient
    const int t_zero_from = dead ? 0 : Q;
    for (size_t i = (size_t)t_zero_from * T + threadIdx.x; i < (size_t)F * T; i += CTC_THREADS) dx[i] = 0.f;
    if (dead) return;

    const double* la = log_alpha + (size_t)b * F * S_max;
    const double* lse = lse_g + (size_t)b * F;
    double* next = b0;
    double* cur = b1;
    for (int t = Q - 1; t >= 0; --t) {
        const float* row = x + (size_t)t * T;
        const double l = lse[t];
        for (int s = threadIdx.x; s < S; s += CTC_THREADS) {
            const double e = (double)((s & 1) ? row[s >> 1] : blank) - l;
            double v;
            if (t == Q - 1) {
                v = (s == S - 1 || s == S - 2) ? e : (double)CTC_NEG_INF;
            } else {
                const double n0 = next[s];
                const double n1 = s + 1 < S ? next[s + 1] : (double)CTC_NEG_INF;
                const double n2 = ((s & 1) && s + 2 < S) ? next[s + 2] : (double)CTC_NEG_INF;
                v = lse3(n0, n1, n2) + e;
            }
            cur[s] = v;
            if (s & 1) {
                const double post = exp(la[(size_t)t * S_max + s] + v - e + nll);
                dx[(size_t)t * T + (s >> 1)] = (float)((exp(e) - post) * gr);
            }
        }
        // masked keys (k ≥ K): softmax = 0 and no state carries them
        for (int k = K + threadIdx.x; k < T; k += CTC_THREADS) dx[(size_t)t * T + k] = 0.f;
        __syncthreads();
        double* tmp = next; next = cur; cur = tmp;
    }
}

}  // namespace fs2k

using namespace fs2k;

extern "C" size_t fs2k_ctc_alpha_elems(int B, int F, int T) { return (size_t)B * F * (2 * (size_t)T + 1); }

extern "C" int fs2k_ctc_forward_sum_fwd(const float* attn_logprob, const int* key_lens, const int* query_lens, int B, int F,
                                        int T, float blank_logprob, double* lse, double* log_alpha, double* nll, float* loss,
                                        fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && F >= 0 && T >= 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(loss, FS2K_ERR_NULL);
    cudaStream_t s = (cudaStream_t)stream;
    if (B > 0) {
        FS2K_REQUIRE(attn_logprob && key_lens && query_lens && lse && log_alpha && nll, FS2K_ERR_NULL);
        const size_t smem = (2 * (2 * (size_t)T + 1) + (size_t)F) * sizeof(double);
        FS2K_REQUIRE(smem <= 200 * 1024, FS2K_ERR_UNSUPPORTED);
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(ctc_alpha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return fs2k_set_cuda_error(e);
        }
        fs2k_launch_serial(ctc_alpha_kernel, dim3(B), dim3(CTC_THREADS), smem, s, attn_logprob, key_lens, query_lens, F, T, blank_logprob, lse, log_alpha, nll);
        FS2K_CHECK_LAUNCH();
    }
    fs2k_launch(ctc_finalize_kernel, dim3(1), dim3(1), 0, s, nll, key_lens, B, T, loss);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_ctc_forward_sum_bwd(const float* attn_logprob, const double* lse, const double* log_alpha, const double* nll,
                                        const int* key_lens, const int* query_lens, const float* gout, int B, int F, int T,
                                        float blank_logprob, float* d_attn_logprob, fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && F >= 0 && T >= 0, FS2K_ERR_BAD_SHAPE);
    if (B == 0 || F == 0 || T == 0) return FS2K_OK;
    FS2K_REQUIRE(attn_logprob && lse && log_alpha && nll && key_lens && query_lens && gout && d_attn_logprob, FS2K_ERR_NULL);
    const size_t smem = 2 * (2 * (size_t)T + 1) * sizeof(double);
    FS2K_REQUIRE(smem <= 200 * 1024, FS2K_ERR_UNSUPPORTED);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(ctc_beta_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    }
    fs2k_launch_serial(ctc_beta_grad_kernel, dim3(B), dim3(CTC_THREADS), smem, (cudaStream_t)stream, attn_logprob, lse, log_alpha, nll, key_lens, query_lens,
                                                                         gout, B, F, T, blank_logprob, d_attn_logprob);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}
