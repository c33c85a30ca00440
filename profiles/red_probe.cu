// Stand-alone probe (not product code): split-K partial tiles folded into the weight gradient with vector reductions in L2
// (red.global.add.v4.f32) versus the workspace + reduce-kernel scheme, at the wgrad shapes of the training step.
// Each CTA owns one 128 x tile_k fp32 partial tile (as the wgrad kernels do after tcgen05.ld) and either
//   A) adds it into out[N][K] with red.v4 (row-strided 16-byte accesses, 8 warps = 4 row quarters x 2 column halves), or
//   B) stores it to ws[split][N][K] with float4 stores, followed by the reduce kernel (splits planes -> out).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/red_probe.bin profiles/red_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) tile_out_kernel(float* out, float* ws, int N, int K, int tile_k, int use_red) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, half = warp >> 2;
    const int n0 = blockIdx.x * 128, k0 = blockIdx.y * tile_k, split = blockIdx.z;
    const int n = n0 + q * 32 + lane;
    float* dst = use_red ? out + (size_t)n * K + k0 : ws + ((size_t)split * N + n) * K + k0;
    for (int c0 = half * 16; c0 < tile_k; c0 += 32) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 1.0f + 0.001f * (float)(c0 + j);
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            if (use_red)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
            else
                *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
    }
}

__global__ void __launch_bounds__(256) reduce_kernel(const float* __restrict__ ws, int splits, long total4, float* __restrict__ out) {
    __shared__ float4 s_part[8][32];
    const int lane = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const long i4 = (long)blockIdx.x * 32 + lane;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i4 < total4) {
        const float4* p = reinterpret_cast<const float4*>(ws) + i4;
        for (int sp = sl; sp < splits; sp += 8) {
            const float4 v = p[(size_t)sp * total4];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    s_part[sl][lane] = acc;
    __syncthreads();
    if (sl == 0 && i4 < total4) {
        for (int g = 1; g < 8; ++g) { const float4 v = s_part[g][lane]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        float4* o = reinterpret_cast<float4*>(out) + i4;
        const float4 c = *o;
        *o = make_float4(acc.x + c.x, acc.y + c.y, acc.z + c.z, acc.w + c.w);
    }
}

int main() {
    struct Shape { int N, K, splits; } shapes[] = {{256, 256, 74}, {1024, 256, 19}, {256, 1024, 19}, {768, 256, 25}, {512, 256, 37}};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (auto sh : shapes) {
        const int tile_k = 256;
        float *out, *ws;
        cudaMalloc(&out, (size_t)sh.N * sh.K * 4);
        cudaMalloc(&ws, (size_t)sh.splits * sh.N * sh.K * 4);
        cudaMemset(out, 0, (size_t)sh.N * sh.K * 4);
        dim3 grid(sh.N / 128, sh.K / tile_k, sh.splits);
        const long total4 = (long)sh.N * sh.K / 4;
        float ms[2];
        for (int mode = 0; mode < 2; ++mode) {
            for (int it = 0; it < 3; ++it) {
                tile_out_kernel<<<grid, 256>>>(out, ws, sh.N, sh.K, tile_k, mode);
                if (!mode) reduce_kernel<<<(int)((total4 + 31) / 32), 256>>>(ws, sh.splits, total4, out);
            }
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            for (int it = 0; it < 50; ++it) {
                tile_out_kernel<<<grid, 256>>>(out, ws, sh.N, sh.K, tile_k, mode);
                if (!mode) reduce_kernel<<<(int)((total4 + 31) / 32), 256>>>(ws, sh.splits, total4, out);
            }
            cudaEventRecord(e1);
            cudaDeviceSynchronize();
            cudaEventElapsedTime(&ms[mode], e0, e1);
        }
        printf("N=%4d K=%4d splits=%3d  (%5.1f MB of partial tiles)  workspace+reduce %7.2f us   red.v4 %7.2f us   err=%s\n", sh.N, sh.K, sh.splits,
               (double)sh.splits * sh.N * sh.K * 4 / 1e6, ms[0] * 1e3 / 50, ms[1] * 1e3 / 50, cudaGetErrorString(cudaGetLastError()));
        cudaFree(out); cudaFree(ws);
    }
    return 0;
}
