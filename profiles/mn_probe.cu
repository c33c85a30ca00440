// Stand-alone probe (not product code): which shared-memory descriptor fields make tcgen05.mma kind::f16 read
// MN-major bf16 tiles (64-column blocks × 64 contraction rows × 128 B, TMA-style 128-byte swizzle) correctly.
// One CTA, D[128×128] = A[128×64] · B[128×64]ᵀ with each operand given either K-major or MN-major; every
// (LBO, SBO, k-step) candidate is compared against a host reference.  Build: see profiles/README (nvcc -arch sm_100a).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../fastspeech2_lightning_b200/csrc/bf16_common.cuh"

int g_fs2k_pdl_enabled = 0;
int fs2k_set_cuda_error(cudaError_t) { return -5; }

using namespace fs2k;

// A_km: [128][64] row-major (K contiguous); A_mn: [64][128] (M contiguous) — same values transposed.  Same for B.
__global__ void __launch_bounds__(128) probe_kernel(const __nv_bfloat16* A_km, const __nv_bfloat16* A_mn, const __nv_bfloat16* B_km,
                                                    const __nv_bfloat16* B_mn, int a_mn, int b_mn, uint32_t lbo, uint32_t sbo,
                                                    uint32_t kstep, float* D) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_done;
    __shared__ uint32_t s_tmem;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;
    uint8_t* sb = smem + 16384;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // K-major tile: 128 rows × 64 cols, row = 128 B.  MN-major tile: 64 rows (k) × 128 cols (mn) = 2 blocks of 64 cols.
    for (int i = tid; i < 128 * 64; i += 128) {
        if (!a_mn) {
            const int r = i / 64, c = i % 64;
            *reinterpret_cast<__nv_bfloat16*>(sa + hb_tile_off(r, c & ~3, 16384) + (c & 3) * 2) = A_km[r * 64 + c];
        } else {
            const int r = i / 128, c = i % 128;  // r = k, c = m
            *reinterpret_cast<__nv_bfloat16*>(sa + hb_tile_off(r, c & ~3, 8192) + (c & 3) * 2) = A_mn[r * 128 + c];
        }
        if (!b_mn) {
            const int r = i / 64, c = i % 64;
            *reinterpret_cast<__nv_bfloat16*>(sb + hb_tile_off(r, c & ~3, 16384) + (c & 3) * 2) = B_km[r * 64 + c];
        } else {
            const int r = i / 128, c = i % 128;
            *reinterpret_cast<__nv_bfloat16*>(sb + hb_tile_off(r, c & ~3, 8192) + (c & 3) * 2) = B_mn[r * 128 + c];
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) {
        mbar_init(smem_u32(&s_done), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        const uint32_t idesc = hb_idesc(128, a_mn, b_mn);
        for (int k = 0; k < 4; ++k) {
            const uint64_t ad = a_mn ? hb_desc_mn(smem_u32(sa) + k * kstep, lbo, sbo) : hb_desc_k(smem_u32(sa) + k * 32);
            const uint64_t bd = b_mn ? hb_desc_mn(smem_u32(sb) + k * kstep, lbo, sbo) : hb_desc_k(smem_u32(sb) + k * 32);
            hb_mma(tmem, ad, bd, idesc, k ? 1u : 0u);
        }
        tc_commit(smem_u32(&s_done));
    }
    mbar_wait(smem_u32(&s_done), 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < 128; c0 += 16) {
        float v[16];
        tc_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        for (int j = 0; j < 16; ++j) D[row * 128 + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
    }
}

int main() {
    std::vector<float> a(128 * 64), b(128 * 64);
    srand(1);
    for (auto& v : a) v = (float)(rand() % 17 - 8) / 8.f;
    for (auto& v : b) v = (float)(rand() % 13 - 6) / 4.f;
    std::vector<__nv_bfloat16> a_km(128 * 64), a_mn(64 * 128), b_km(128 * 64), b_mn(64 * 128);
    for (int r = 0; r < 128; ++r)
        for (int k = 0; k < 64; ++k) {
            a_km[r * 64 + k] = __float2bfloat16(a[r * 64 + k]);
            a_mn[k * 128 + r] = __float2bfloat16(a[r * 64 + k]);
            b_km[r * 64 + k] = __float2bfloat16(b[r * 64 + k]);
            b_mn[k * 128 + r] = __float2bfloat16(b[r * 64 + k]);
        }
    std::vector<float> ref(128 * 128);
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 128; ++n) {
            float s = 0.f;
            for (int k = 0; k < 64; ++k) s += a[m * 64 + k] * b[n * 64 + k];  // values are exact in bf16
            ref[m * 128 + n] = s;
        }
    __nv_bfloat16 *dA, *dAm, *dB, *dBm;
    float* dD;
    cudaMalloc(&dA, 128 * 64 * 2); cudaMalloc(&dAm, 128 * 64 * 2); cudaMalloc(&dB, 128 * 64 * 2); cudaMalloc(&dBm, 128 * 64 * 2);
    cudaMalloc(&dD, 128 * 128 * 4);
    cudaMemcpy(dA, a_km.data(), 128 * 64 * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dAm, a_mn.data(), 128 * 64 * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, b_km.data(), 128 * 64 * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dBm, b_mn.data(), 128 * 64 * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
    const uint32_t cand[][3] = {{8192, 1024, 2048}, {1024, 8192, 2048}, {8192, 1024, 256}, {1024, 8192, 256}, {128, 1024, 2048}, {8192, 128, 2048}};
    std::vector<float> d(128 * 128);
    for (int mode = 0; mode < 4; ++mode) {
        const int a_mn_f = mode & 1, b_mn_f = mode >> 1;
        for (auto& c : cand) {
            cudaMemset(dD, 0, 128 * 128 * 4);
            probe_kernel<<<1, 128, 40 * 1024>>>(dA, dAm, dB, dBm, a_mn_f, b_mn_f, c[0], c[1], c[2], dD);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode a_mn=%d b_mn=%d lbo=%u sbo=%u kstep=%u: CUDA error %s\n", a_mn_f, b_mn_f, c[0], c[1], c[2], cudaGetErrorString(e)); return 1; }
            cudaMemcpy(d.data(), dD, 128 * 128 * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0;
            for (int i = 0; i < 128 * 128; ++i) maxerr = fmax(maxerr, fabs((double)d[i] - ref[i]));
            printf("mode a_mn=%d b_mn=%d lbo=%u sbo=%u kstep=%u: maxerr %.4g %s\n", a_mn_f, b_mn_f, c[0], c[1], c[2], maxerr, maxerr < 1e-3 ? "OK" : "");
            if (mode == 0) break;  // K-major both: candidates irrelevant
        }
    }
    return 0;
}
