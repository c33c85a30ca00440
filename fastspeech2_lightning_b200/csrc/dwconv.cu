// Depthwise 1-D convolutions on channels-last data (HBM-bound stencils).
//
//   Conformer conv module middle:  GLU → depthwise k=9 (no length mask) → BatchNorm1d → SiLU
//                                  torchaudio conformer.py:50-65
//   variance predictor:            depthwise k=3 (+bias), fs2/blocks.py:8-15
// One thread owns one channel and slides along time with the last K inputs in registers, so every
// input element is read once per tile (+K−1 halo) with fully coalesced 4-byte accesses across the
// channel dimension.  Padding frames are NOT masked (they are live in the reference).
//
// mode 0: y = conv + bias                         (training: pre-BatchNorm z; predictor stencil)
// mode 1: y = silu((conv + bias)·scale + shift)   (eval: BatchNorm folded into scale/shift)
#include "common.cuh"

namespace fs2k {

constexpr int kDwTile = 16;   // frames per CTA
constexpr int kDwUnroll = 8;  // frames whose inputs are in flight together

template <int K, bool GLU, int MODE>
__global__ void __launch_bounds__(256)
dwconv_kernel(const float* __restrict__ x,  // [B,L,ldx]  (GLU: value at c, gate at c + C)
              int ldx, int L, int C,
              const float* __restrict__ w,     // [C][K]  (PyTorch [C,1,K])
              const float* __restrict__ bias,  // [C] or null
              const float* __restrict__ scale, const float* __restrict__ shift,
              float* __restrict__ y)           // [B,L,C]
{
    pdl_prologue();
    constexpr int P = (K - 1) / 2;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.z;
    const int l0 = blockIdx.y * kDwTile;
    if (c >= C) return;
    float wk[K];
#pragma unroll
    for (int k = 0; k < K; ++k) wk[k] = w[(size_t)c * K + k];
    const float bs = bias ? bias[c] : 0.f;
    float sc = 1.f, sh = 0.f;
    if (MODE == 1) { sc = scale[c]; sh = shift[c]; }
    const float* xb = x + (size_t)b * L * ldx;
    auto fetch = [&](int l) -> float {
        if (l < 0 || l >= L) return 0.f;
        const float a = xb[(size_t)l * ldx + c];
        if (!GLU) return a;
        const float g = xb[(size_t)l * ldx + C + c];
        return a / (1.0f + expf(-g));  // a · sigmoid(g)
    };
    float win[K];
#pragma unroll
    for (int k = 0; k < K - 1; ++k) win[k + 1] = fetch(l0 - P + k);
    const int l_end = min(l0 + kDwTile, L);
    // kDwUnroll frames per trip: their inputs are fetched first (independent loads in flight), then consumed — a
    // one-load-per-step loop exposes the full L2 latency on every frame (26 µs for a 16×80×256 tile grid before)
    for (int l = l0; l < l_end; l += kDwUnroll) {
        float nx[kDwUnroll];
#pragma unroll
        for (int u = 0; u < kDwUnroll; ++u) nx[u] = (l + u < l_end) ? fetch(l + u + P) : 0.f;
#pragma unroll
        for (int u = 0; u < kDwUnroll; ++u) {
#pragma unroll
            for (int k = 0; k < K - 1; ++k) win[k] = win[k + 1];
            win[K - 1] = nx[u];
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) acc = fmaf(win[k], wk[k], acc);
            acc += bs;
            if (MODE == 1) acc = silu(acc * sc + sh);
            if (l + u < l_end) y[((size_t)b * L + l + u) * C + c] = acc;
        }
    }
}

}  // namespace fs2k

using namespace fs2k;

template <int K>
static int launch_dw(const float* x, int ldx, int B, int L, int C, const float* w, const float* bias, int glu,
                     const float* scale, const float* shift, float* y, cudaStream_t s) {
    dim3 grid(cdiv(C, 256), cdiv(L, kDwTile), B);
    const int threads = C < 256 ? ((C + 31) / 32) * 32 : 256;
    grid.x = cdiv(C, threads);
    if (glu) {
        if (scale) fs2k_launch(dwconv_kernel<K, true, 1>, dim3(grid), dim3(threads), 0, s, x, ldx, L, C, w, bias, scale, shift, y);
        else fs2k_launch(dwconv_kernel<K, true, 0>, dim3(grid), dim3(threads), 0, s, x, ldx, L, C, w, bias, scale, shift, y);
    } else {
        if (scale) fs2k_launch(dwconv_kernel<K, false, 1>, dim3(grid), dim3(threads), 0, s, x, ldx, L, C, w, bias, scale, shift, y);
        else fs2k_launch(dwconv_kernel<K, false, 0>, dim3(grid), dim3(threads), 0, s, x, ldx, L, C, w, bias, scale, shift, y);
    }
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_dwconv_fwd(const float* x, int ldx, int B, int L, int C, const float* w, int K, const float* bias,
                               int glu, const float* scale, const float* shift, float* y, fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && L >= 0 && C > 0 && K > 0 && (K & 1), FS2K_ERR_BAD_SHAPE);
    if (B == 0 || L == 0) return FS2K_OK;
    FS2K_REQUIRE(x && w && y, FS2K_ERR_NULL);
    FS2K_REQUIRE(!scale || shift, FS2K_ERR_NULL);
    FS2K_REQUIRE(B <= 65535, FS2K_ERR_UNSUPPORTED);
    cudaStream_t s = (cudaStream_t)stream;
    switch (K) {
        case 1: return launch_dw<1>(x, ldx, B, L, C, w, bias, glu, scale, shift, y, s);
        case 3: return launch_dw<3>(x, ldx, B, L, C, w, bias, glu, scale, shift, y, s);
        case 5: return launch_dw<5>(x, ldx, B, L, C, w, bias, glu, scale, shift, y, s);
        case 7: return launch_dw<7>(x, ldx, B, L, C, w, bias, glu, scale, shift, y, s);
        case 9: return launch_dw<9>(x, ldx, B, L, C, w, bias, glu, scale, shift, y, s);
        case 11: return launch_dw<11>(x, ldx, B, L, C, w, bias, glu, scale, shift, y, s);
        case 13: return launch_dw<13>(x, ldx, B, L, C, w, bias, glu, scale, shift, y, s);
        case 15: return launch_dw<15>(x, ldx, B, L, C, w, bias, glu, scale, shift, y, s);
        case 31: return launch_dw<31>(x, ldx, B, L, C, w, bias, glu, scale, shift, y, s);
        default: return FS2K_ERR_UNSUPPORTED;
    }
}
