// Global-style-token reference encoder — reference fs2/gst/model.py:103-257.
//
//   ReferenceEncoder: 6 × [Conv2d(3×3, stride 2, pad 1, no bias) → BatchNorm2d → ReLU] over the mel "image"
//                     [B,1,F,80], then a one-layer GRU(128) over the remaining time steps; last hidden state.
//   StyleTokenLayer:  4-head attention of that state over tanh(10 learned tokens).
//
// The convolutions are tiny (3 GFLOP for B=32, F=500) and shrink 4× per layer, so they are direct convolutions in
// channels-last layout: one thread per output value, output channel fastest (the 9·Cin input values are warp
// broadcasts, the re-packed weights [kh][kw][ci][co] are coalesced).  BatchNorm is folded (eval) into scale/shift.
// The GRU's input projection for all steps and the per-step hidden projection are fs2k_gemm_* calls; gru_gate is
// the elementwise cell update.  The training kernels (raw conv, dgrad, wgrad, GRU / attention backward) follow below.
#include "common.cuh"

namespace fs2k {

// x [B,H,W,Ci] → y = relu(conv·scale + shift): [B,Ho,Wo,Co] (cw_layout 0) or [B,Ho,Co,Wo] (cw_layout 1: the
// reference's `.transpose(1,2).view(B,T,C·W)` feature order for the GRU).  w [3][3][Ci][Co].
__global__ void __launch_bounds__(256)
conv2d_s2_bn_relu_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale,
                         const float* __restrict__ shift, int B, int H, int W, int Ci, int Co, int Ho, int Wo,
                         int cw_layout, int raw, float* __restrict__ y) {
    pdl_prologue();
    const long total = (long)B * Ho * Wo * Co;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int co = (int)(i % Co);
        const int wo = (int)((i / Co) % Wo);
        const int ho = (int)((i / ((long)Co * Wo)) % Ho);
        const int b = (int)(i / ((long)Co * Wo * Ho));
        float acc = 0.f;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int h = 2 * ho + kh - 1;
            if (h < 0 || h >= H) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int ww = 2 * wo + kw - 1;
                if (ww < 0 || ww >= W) continue;
                const float* xp = x + (((size_t)b * H + h) * W + ww) * Ci;
                const float* wp = w + ((size_t)(kh * 3 + kw) * Ci) * Co + co;
                for (int ci = 0; ci < Ci; ++ci) acc = fmaf(xp[ci], wp[(size_t)ci * Co], acc);
            }
        }
        const float v = raw ? acc : fmaxf(acc * scale[co] + shift[co], 0.f);  // raw: pre-BatchNorm output (training)
        const size_t o = cw_layout ? (((size_t)b * Ho + ho) * Co + co) * Wo + wo : (size_t)i;
        y[o] = v;
    }
}

// torch.nn.GRU cell (gate order r, z, n): xp = W_ih·x + b_ih, hp = W_hh·h + b_hh, both [B,3U]
//   r = σ(xp_r + hp_r), z = σ(xp_z + hp_z), n = tanh(xp_n + r·hp_n), h' = (1 − z)·n + z·h
__global__ void gru_gate_kernel(const float* __restrict__ xp, long xp_stride, const float* __restrict__ hp,
                                const float* __restrict__ h, int B, int U, float* __restrict__ h_out) {
    pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * U) return;
    const int b = i / U, u = i - b * U;
    const float* x = xp + (size_t)b * xp_stride;
    const float* p = hp + (size_t)b * 3 * U;
    const float r = 1.0f / (1.0f + expf(-(x[u] + p[u])));
    const float z = 1.0f / (1.0f + expf(-(x[U + u] + p[U + u])));
    const float n = tanhf(x[2 * U + u] + r * p[2 * U + u]);
    h_out[i] = (1.0f - z) * n + z * h[i];
}

// q [B,heads·dk], k/v [T,heads·dk] (T ≤ 32 tokens): out[b, hd·dk + d] = Σ_t softmax_t(q·k_t / √dk)·v[t, hd·dk + d].
// One warp per (b, head): lane t scores token t.
__global__ void __launch_bounds__(128)
gst_token_attention_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int B,
                           int T, int heads, int dk, float* __restrict__ out) {
    pdl_prologue();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B * heads) return;
    const int b = warp / heads, hd = warp - b * heads;
    const int D = heads * dk;
    const float* qp = q + (size_t)b * D + hd * dk;
    float s = -INFINITY;
    if (lane < T) {
        const float* kp = k + (size_t)lane * D + hd * dk;
        float a = 0.f;
        for (int d = 0; d < dk; ++d) a = fmaf(qp[d], kp[d], a);
        s = a / sqrtf((float)dk);
    }
    const float m = warp_max(s);
    const float e = lane < T ? expf(s - m) : 0.f;
    const float p = e / warp_sum(e);
    for (int d0 = 0; d0 < dk; d0 += 32) {
        const int d = d0 + lane;
        float acc = 0.f;
        for (int t = 0; t < T; ++t) {
            const float pt = __shfl_sync(0xffffffffu, p, t);
            if (d < dk) acc = fmaf(pt, v[(size_t)t * D + hd * dk + d], acc);
        }
        if (d < dk) out[(size_t)b * D + hd * dk + d] = acc;
    }
}

}  // namespace fs2k

using namespace fs2k;

extern "C" int fs2k_conv2d_s2_bn_relu(const float* x, const float* w_khwcico, const float* scale, const float* shift, int B,
                                      int H, int W, int Ci, int Co, int cw_layout, float* y, fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && H > 0 && W > 0 && Ci > 0 && Co > 0, FS2K_ERR_BAD_SHAPE);
    if (B == 0) return FS2K_OK;
    FS2K_REQUIRE(x && w_khwcico && y && (!scale == !shift), FS2K_ERR_NULL);
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;  // (n + 2·1 − 3)/2 + 1
    const long total = (long)B * Ho * Wo * Co;
    long g = (total + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    fs2k_launch(conv2d_s2_bn_relu_kernel, dim3((unsigned)g), dim3(256), 0, (cudaStream_t)stream, x, w_khwcico, scale, shift, B,
                H, W, Ci, Co, Ho, Wo, cw_layout, scale ? 0 : 1, y);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_gru_gate(const float* xproj, long xproj_row_stride, const float* hproj, const float* h, int B, int U,
                             float* h_out, fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && U > 0, FS2K_ERR_BAD_SHAPE);
    if (B == 0) return FS2K_OK;
    FS2K_REQUIRE(xproj && hproj && h && h_out, FS2K_ERR_NULL);
    fs2k_launch(gru_gate_kernel, dim3(cdiv((long)B * U, 128)), dim3(128), 0, (cudaStream_t)stream, xproj, xproj_row_stride, hproj,
                h, B, U, h_out);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_gst_token_attention(const float* q, const float* k, const float* v, int B, int T, int heads, int dk,
                                        float* out, fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && T > 0 && heads > 0 && dk > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(T <= 32, FS2K_ERR_UNSUPPORTED);
    if (B == 0) return FS2K_OK;
    FS2K_REQUIRE(q && k && v && out, FS2K_ERR_NULL);
    fs2k_launch(gst_token_attention_kernel, dim3(cdiv((long)B * heads * 32, 128)), dim3(128), 0, (cudaStream_t)stream, q, k, v,
                B, T, heads, dk, out);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Training through the reference encoder (same module, gradients): raw convolution (BatchNorm batch statistics and
// ReLU are the library's own BN kernels: fs2k_colstats / fs2k_bn_finalize / fs2k_affine_act / fs2k_bn_act_bwd on the
// [B·Ho·Wo, Co] view), its data / weight gradients, the GRU cell backward and the token-attention backward.
// ---------------------------------------------------------------------------------------------------------------
namespace fs2k {

// dx[b,h,w,ci] = Σ_{kh,kw} Σ_co gz[b,ho,wo,co]·w[kh][kw][ci][co]  over the taps with 2·ho + kh − 1 = h, 2·wo + kw − 1 = w
__global__ void __launch_bounds__(256)
conv2d_s2_dgrad_kernel(const float* __restrict__ gz, const float* __restrict__ w, int B, int H, int W, int Ci, int Co,
                       int Ho, int Wo, float* __restrict__ dx) {
    pdl_prologue();
    const long total = (long)B * H * W * Ci;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Ci);
        const int ww = (int)((i / Ci) % W);
        const int h = (int)((i / ((long)Ci * W)) % H);
        const int b = (int)(i / ((long)Ci * W * H));
        float acc = 0.f;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int th = h + 1 - kh;
            if (th < 0 || (th & 1) || (th >> 1) >= Ho) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int tw = ww + 1 - kw;
                if (tw < 0 || (tw & 1) || (tw >> 1) >= Wo) continue;
                const float* gp = gz + (((size_t)b * Ho + (th >> 1)) * Wo + (tw >> 1)) * Co;
                const float* wp = w + ((size_t)(kh * 3 + kw) * Ci + ci) * Co;
                for (int co = 0; co < Co; ++co) acc = fmaf(gp[co], wp[co], acc);
            }
        }
        dx[i] = acc;
    }
}

// dw[kh][kw][ci][co] += Σ_{positions of this CTA's chunk} x[b,2ho+kh−1,2wo+kw−1,ci]·gz[b,ho,wo,co]   (dw zeroed by the launcher)
// grid (9·Ci, chunks); thread = co (Co ≤ 128 per pass)
__global__ void __launch_bounds__(128)
conv2d_s2_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gz, int B, int H, int W, int Ci, int Co,
                       int Ho, int Wo, long pos_per_cta, float* __restrict__ dw) {
    pdl_prologue();
    const int tap = blockIdx.x / Ci, ci = blockIdx.x - tap * Ci;
    const int kh = tap / 3, kw = tap - kh * 3;
    const long P = (long)B * Ho * Wo;
    const long p0 = (long)blockIdx.y * pos_per_cta, p1 = min(P, p0 + pos_per_cta);
    for (int co = threadIdx.x; co < Co; co += blockDim.x) {
        float acc = 0.f;
        for (long p = p0; p < p1; ++p) {
            const int wo = (int)(p % Wo), ho = (int)((p / Wo) % Ho), b = (int)(p / ((long)Wo * Ho));
            const int h = 2 * ho + kh - 1, ww = 2 * wo + kw - 1;
            if (h < 0 || h >= H || ww < 0 || ww >= W) continue;
            acc = fmaf(x[(((size_t)b * H + h) * W + ww) * Ci + ci], gz[(size_t)p * Co + co], acc);
        }
        atomicAdd(&dw[((size_t)tap * Ci + ci) * Co + co], acc);
    }
}

// GRU cell backward (see gru_gate_kernel).  dh: gradient of h' [B,U]; outputs d_xp [B,3U] (row stride dxp_stride),
// d_hp [B,3U], dh_prev [B,U] = dh·z (the W_hh·d_hp term is added by the caller's GEMM).
__global__ void gru_gate_bwd_kernel(const float* __restrict__ xp, long xp_stride, const float* __restrict__ hp,
                                    const float* __restrict__ h, const float* __restrict__ dh, int B, int U,
                                    float* __restrict__ dxp, long dxp_stride, float* __restrict__ dhp,
                                    float* __restrict__ dh_prev) {
    pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * U) return;
    const int b = i / U, u = i - b * U;
    const float* x = xp + (size_t)b * xp_stride;
    const float* p = hp + (size_t)b * 3 * U;
    const float r = 1.0f / (1.0f + expf(-(x[u] + p[u])));
    const float z = 1.0f / (1.0f + expf(-(x[U + u] + p[U + u])));
    const float hn = p[2 * U + u];
    const float n = tanhf(x[2 * U + u] + r * hn);
    const float g = dh[i];
    const float dn = g * (1.0f - z);
    const float dz = g * (h[i] - n);
    const float dpre = dn * (1.0f - n * n);
    const float dar = dpre * hn * r * (1.0f - r);
    const float daz = dz * z * (1.0f - z);
    float* dx = dxp + (size_t)b * dxp_stride;
    float* dp = dhp + (size_t)b * 3 * U;
    dx[u] = dar; dx[U + u] = daz; dx[2 * U + u] = dpre;
    dp[u] = dar; dp[U + u] = daz; dp[2 * U + u] = dpre * r;
    dh_prev[i] = g * z;
}

// backward of gst_token_attention_kernel: one warp per (b, head); dk / dv are summed over the batch with atomics
// (zeroed by the launcher).
__global__ void __launch_bounds__(128)
gst_token_attention_bwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                               const float* __restrict__ dout, int B, int T, int heads, int dk, float* __restrict__ dq,
                               float* __restrict__ dkk, float* __restrict__ dv) {
    pdl_prologue();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B * heads) return;
    const int b = warp / heads, hd = warp - b * heads;
    const int D = heads * dk;
    const float inv = 1.0f / sqrtf((float)dk);
    const float* qp = q + (size_t)b * D + hd * dk;
    const float* gp = dout + (size_t)b * D + hd * dk;
    float s = -INFINITY, dp = 0.f;
    if (lane < T) {
        const float* kp = k + (size_t)lane * D + hd * dk;
        const float* vp = v + (size_t)lane * D + hd * dk;
        float a = 0.f;
        for (int d = 0; d < dk; ++d) { a = fmaf(qp[d], kp[d], a); dp = fmaf(gp[d], vp[d], dp); }
        s = a * inv;
    }
    const float m = warp_max(s);
    const float e = lane < T ? expf(s - m) : 0.f;
    const float p = e / warp_sum(e);
    const float ds = p * (dp - warp_sum(p * dp)) * inv;  // d(score·√dk⁻¹) folded: gradient w.r.t. q·k
    for (int d0 = 0; d0 < dk; d0 += 32) {
        const int d = d0 + lane;
        float accq = 0.f;
        for (int t = 0; t < T; ++t) {
            const float dst = __shfl_sync(0xffffffffu, ds, t), pt = __shfl_sync(0xffffffffu, p, t);
            if (d < dk) {
                accq = fmaf(dst, k[(size_t)t * D + hd * dk + d], accq);
                atomicAdd(&dkk[(size_t)t * D + hd * dk + d], dst * qp[d]);
                atomicAdd(&dv[(size_t)t * D + hd * dk + d], pt * gp[d]);
            }
        }
        if (d < dk) dq[(size_t)b * D + hd * dk + d] = accq;
    }
}

}  // namespace fs2k

extern "C" int fs2k_conv2d_s2_dgrad(const float* gz, const float* w_khwcico, int B, int H, int W, int Ci, int Co, float* dx,
                                    fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && H > 0 && W > 0 && Ci > 0 && Co > 0, FS2K_ERR_BAD_SHAPE);
    if (B == 0) return FS2K_OK;
    FS2K_REQUIRE(gz && w_khwcico && dx, FS2K_ERR_NULL);
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long total = (long)B * H * W * Ci;
    long g = (total + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    fs2k_launch(conv2d_s2_dgrad_kernel, dim3((unsigned)g), dim3(256), 0, (cudaStream_t)stream, gz, w_khwcico, B, H, W, Ci, Co, Ho,
                Wo, dx);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_conv2d_s2_wgrad(const float* x, const float* gz, int B, int H, int W, int Ci, int Co, float* dw_khwcico,
                                    fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && H > 0 && W > 0 && Ci > 0 && Co > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(dw_khwcico, FS2K_ERR_NULL);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(dw_khwcico, 0, sizeof(float) * 9 * (size_t)Ci * Co, s);
    if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    if (B == 0) return FS2K_OK;
    FS2K_REQUIRE(x && gz, FS2K_ERR_NULL);
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long P = (long)B * Ho * Wo;
    long chunks = (148L * 8 + 9 * Ci - 1) / (9 * Ci);
    if (chunks < 1) chunks = 1;
    if (chunks > 1024) chunks = 1024;
    long per = (P + chunks - 1) / chunks;
    if (per < 16) per = 16;
    chunks = (P + per - 1) / per;
    fs2k_launch(conv2d_s2_wgrad_kernel, dim3(9 * Ci, (unsigned)chunks), dim3(Co < 128 ? ((Co + 31) / 32) * 32 : 128), 0, s, x, gz,
                B, H, W, Ci, Co, Ho, Wo, per, dw_khwcico);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_gru_gate_bwd(const float* xproj, long xproj_row_stride, const float* hproj, const float* h,
                                 const float* dh, int B, int U, float* dxproj, long dxproj_row_stride, float* dhproj,
                                 float* dh_prev, fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && U > 0, FS2K_ERR_BAD_SHAPE);
    if (B == 0) return FS2K_OK;
    FS2K_REQUIRE(xproj && hproj && h && dh && dxproj && dhproj && dh_prev, FS2K_ERR_NULL);
    fs2k_launch(gru_gate_bwd_kernel, dim3(cdiv((long)B * U, 128)), dim3(128), 0, (cudaStream_t)stream, xproj, xproj_row_stride,
                hproj, h, dh, B, U, dxproj, dxproj_row_stride, dhproj, dh_prev);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_gst_token_attention_bwd(const float* q, const float* k, const float* v, const float* dout, int B, int T,
                                            int heads, int dk, float* dq, float* dk_out, float* dv_out, fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && T > 0 && heads > 0 && dk > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(T <= 32, FS2K_ERR_UNSUPPORTED);
    FS2K_REQUIRE(dk_out && dv_out, FS2K_ERR_NULL);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(dk_out, 0, sizeof(float) * (size_t)T * heads * dk, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(dv_out, 0, sizeof(float) * (size_t)T * heads * dk, s);
    if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    if (B == 0) return FS2K_OK;
    FS2K_REQUIRE(q && k && v && dout && dq, FS2K_ERR_NULL);
    fs2k_launch(gst_token_attention_bwd_kernel, dim3(cdiv((long)B * heads * 32, 128)), dim3(128), 0, s, q, k, v, dout, B, T, heads,
                dk, dq, dk_out, dv_out);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}
