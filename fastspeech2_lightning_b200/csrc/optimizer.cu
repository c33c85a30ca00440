// Training-step tail over one flat parameter buffer: global gradient norm, norm clipping and AdamW in two
// launches (reference: torch.optim.AdamW at fs2/model.py:530-537, gradient_clip_val = 1.0 at fs2/cli/train.py:38,
// which Lightning applies as clip_grad_norm_).  The clip coefficient stays on the device — no host sync.
#include <cuda_bf16.h>

#include "common.cuh"

namespace fs2k {

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, long N, double* __restrict__ out) {
    pdl_prologue();
    double acc = 0.0;
    const long N4 = N >> 2;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < N4; i += (long)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(g)[i];
        acc += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (long i = N4 << 2; i < N; ++i) acc += (double)g[i] * g[i];
    acc = warp_sum_d(acc);
    __shared__ double s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) t += s[w];
        atomicAdd(out, t);
    }
}

// clip: g ← g · min(1, max_norm / (‖g‖ + 1e-6)) (skipped when sumsq is null), then decoupled-weight-decay Adam
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long N,
             float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt, float max_norm,
             float grad_scale, const double* __restrict__ sumsq, const float* __restrict__ step_state,
             __nv_bfloat16* __restrict__ p16) {
    pdl_prologue();
    if (step_state) {  // graph replay: the step's scalars come from device memory
        lr = step_state[0];
        bc1 = step_state[1];
        bc2_sqrt = step_state[2];
    }
    float clip = grad_scale;
    if (sumsq) {
        const float total = (float)sqrt(sumsq[0]) * grad_scale;
        const float c = max_norm / (total + 1e-6f);
        clip *= c < 1.f ? c : 1.f;
    }
    const float step_size = lr / bc1;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long)gridDim.x * blockDim.x) {
        const float gi = g[i] * clip;
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        const float pi = p[i] * (1.f - lr * wd) - step_size * (mi / denom);
        p[i] = pi;
        if (p16) p16[i] = __float2bfloat16_rn(pi);  // bf16 shadow of the weights: the bf16-mode GEMM operands, no cast pass
    }
}

}  // namespace fs2k

using namespace fs2k;

extern "C" int fs2k_sumsq(const float* g, long N, double* out, fs2k_stream_t stream) {
    FS2K_REQUIRE(N >= 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(g && out, FS2K_ERR_NULL);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double), s);
    if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    if (N == 0) return FS2K_OK;
    long grid = (N / 4 + 255) / 256;
    if (grid > 148 * 4) grid = 148 * 4;
    if (grid < 1) grid = 1;
    fs2k_launch(sumsq_kernel, dim3((int)grid), dim3(256), 0, s, g, N, out);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_adamw_step(float* p, const float* g, float* m, float* v, long N, float lr, float beta1, float beta2,
                               float eps, float weight_decay, long step, float max_norm, float grad_scale,
                               const double* sumsq, void* p_bf16, fs2k_stream_t stream) {
    FS2K_REQUIRE(N >= 0 && step >= 1, FS2K_ERR_BAD_SHAPE);
    if (N == 0) return FS2K_OK;
    FS2K_REQUIRE(p && g && m && v, FS2K_ERR_NULL);
    // double arithmetic on the float betas: the same numbers FusedAdamW.begin_graph_step publishes for a replayed step, so
    // replicas stay bit-identical when one rank runs this entry point and another replays its captured update
    const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    long grid = (N + 255) / 256;
    if (grid > 148 * 8) grid = 148 * 8;
    fs2k_launch(adamw_kernel, dim3((int)grid), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, N, lr, beta1, beta2, eps, weight_decay, bc1,
                                                             bc2_sqrt, max_norm, grad_scale, sumsq, nullptr, (__nv_bfloat16*)p_bf16);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_adamw_step_dev(float* p, const float* g, float* m, float* v, long N, const float* step_state,
                                   float beta1, float beta2, float eps, float weight_decay, float max_norm,
                                   float grad_scale, const double* sumsq, void* p_bf16, fs2k_stream_t stream) {
    FS2K_REQUIRE(N >= 0, FS2K_ERR_BAD_SHAPE);
    if (N == 0) return FS2K_OK;
    FS2K_REQUIRE(p && g && m && v && step_state, FS2K_ERR_NULL);
    long grid = (N + 255) / 256;
    if (grid > 148 * 8) grid = 148 * 8;
    fs2k_launch(adamw_kernel, dim3((int)grid), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, N, 0.f, beta1, beta2, eps, weight_decay, 1.f, 1.f,
                                                             max_norm, grad_scale, sumsq, step_state, (__nv_bfloat16*)p_bf16);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

namespace fs2k {
__global__ void set_step_state_kernel(float* __restrict__ st, unsigned long long* __restrict__ seed_base, float lr, float bc1,
                                      float bc2_sqrt, unsigned long long base) {
    pdl_prologue();
    if (st) { st[0] = lr; st[1] = bc1; st[2] = bc2_sqrt; st[3] = 0.f; }
    if (seed_base) seed_base[0] = base;
}
}  // namespace fs2k

extern "C" int fs2k_set_step_state(float* step_state, unsigned long long* seed_base, float lr, float bias_correction1,
                                   float bias_correction2_sqrt, long seed_base_value, fs2k_stream_t stream) {
    FS2K_REQUIRE(step_state || seed_base, FS2K_ERR_NULL);
    fs2k_launch(set_step_state_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, step_state, seed_base, lr, bias_correction1, bias_correction2_sqrt,
                                                            (unsigned long long)seed_base_value);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}
