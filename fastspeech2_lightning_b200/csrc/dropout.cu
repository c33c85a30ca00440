// Dropout (training only): y = x · keep / (1 − p), keep ~ Bernoulli(1 − p) from a counter-based hash of
// (seed, element index) — the backward pass regenerates the same mask from the seed instead of storing it.
// Reference call sites: torchaudio conformer.py:74,106,108,157 ; fs2/layers.py:43,208-209.
#include "common.cuh"

namespace fs2k {

__global__ void __launch_bounds__(256)
dropout_kernel(const float* __restrict__ x, const float* __restrict__ residual, float p, float inv_keep,
               unsigned long long seed, long N, float* __restrict__ y) {
    pdl_prologue();
    seed = seed_with_base(seed);
    const uint32_t thr16 = drop_thr16(p);
    const float keep_scale = drop_inv_keep(thr16);
    (void)inv_keep;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long)gridDim.x * blockDim.x) {
        const float v = drop_keep(seed, (unsigned long long)i, thr16) ? x[i] * keep_scale : 0.f;
        y[i] = residual ? v + residual[i] : v;
    }
}

}  // namespace fs2k

extern "C" int fs2k_dropout(const float* x, const float* residual, float p, long seed, long N, float* y,
                            fs2k_stream_t stream) {
    FS2K_REQUIRE(N >= 0 && p >= 0.f && p < 1.f, FS2K_ERR_BAD_SHAPE);
    if (N == 0) return FS2K_OK;
    FS2K_REQUIRE(x && y, FS2K_ERR_NULL);
    long g = (N + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    fs2k_launch(fs2k::dropout_kernel, dim3((int)g), dim3(256), 0, (cudaStream_t)stream, x, residual, p, 1.0f / (1.0f - p), (unsigned long long)seed, N, y);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

FS2K_DEFINE_SEED_BASE_SETTER(dropout)
