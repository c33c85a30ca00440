// Global-style-token reference encoder, inference path — reference fs2/gst/model.py:103-257.
//
//   ReferenceEncoder: 6 × [Conv2d(3×3, stride 2, pad 1, no bias) → BatchNorm2d → ReLU] over the mel "image"
//                     [B,1,F,80], then a one-layer GRU(128) over the remaining time steps; last hidden state.
//   StyleTokenLayer:  4-head attention of that state over tanh(10 learned tokens).
//
// The convolutions are tiny (3 GFLOP for B=32, F=500) and shrink 4× per layer, so they are direct convolutions in
// channels-last layout: one thread per output value, output channel fastest (the 9·Cin input values are warp
// broadcasts, the re-packed weights [kh][kw][ci][co] are coalesced).  BatchNorm is folded (eval) into scale/shift.
// The GRU's input projection for all steps and the per-step hidden projection are fs2k_gemm_* calls; gru_gate is
// the elementwise cell update.  Training through this module stays on torch autograd (library kernels).
#include "common.cuh"

namespace fs2k {

// x [B,H,W,Ci] → y = relu(conv·scale + shift): [B,Ho,Wo,Co] (cw_layout 0) or [B,Ho,Co,Wo] (cw_layout 1: the
// reference's `.transpose(1,2).view(B,T,C·W)` feature order for the GRU).  w [3][3][Ci][Co].
__global__ void __launch_bounds__(256)
conv2d_s2_bn_relu_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale,
                         const float* __restrict__ shift, int B, int H, int W, int Ci, int Co, int Ho, int Wo,
                         int cw_layout, float* __restrict__ y) {
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
        const float v = fmaxf(acc * scale[co] + shift[co], 0.f);
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
    FS2K_REQUIRE(x && w_khwcico && scale && shift && y, FS2K_ERR_NULL);
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;  // (n + 2·1 − 3)/2 + 1
    const long total = (long)B * Ho * Wo * Co;
    long g = (total + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    fs2k_launch(conv2d_s2_bn_relu_kernel, dim3((unsigned)g), dim3(256), 0, (cudaStream_t)stream, x, w_khwcico, scale, shift, B,
                H, W, Ci, Co, Ho, Wo, cw_layout, y);
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
