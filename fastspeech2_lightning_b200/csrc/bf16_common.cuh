// Shared pieces of the bf16 tcgen05 kernels (gemm_bf16.cu, gemm_wgrad_bf16.cu, attention_tc.cu): shared-memory
// matrix descriptors for 128-byte-swizzled bf16 tiles (K-major and MN-major), the kind::f16 instruction descriptor,
// the MMA issue wrapper and fp32 → bf16 packing helpers.
#pragma once
#include <cuda_bf16.h>

#include "tc_common.cuh"

namespace fs2k {

constexpr int HB_BM = 128;       // rows per tile == TMEM lanes

// K-major operand, rows of 64 bf16 = one 128-byte swizzle atom row; 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t hb_desc_k(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                 // LBO unused for swizzled K-major
    d |= (uint64_t)(1024 >> 4) << 32;       // SBO: next group of 8 rows
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
// MN-major operand: atoms of 64 MN-elements (128 B) × 8 contraction rows; LBO = bytes between 64-wide MN blocks,
// SBO = 1024 B between groups of 8 contraction rows
__device__ __forceinline__ uint64_t hb_desc_mn(uint32_t addr, uint32_t lbo, uint32_t sbo = 1024) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor: D = f32, A = B = bf16, M = 128, N = n; major bits 15 (A) / 16 (B): 1 = MN-major
__host__ __device__ constexpr uint32_t hb_idesc(int n, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(HB_BM >> 4) << 24);
}
__device__ __forceinline__ void hb_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}

__device__ __forceinline__ float hb_act(float v, int act) {
    if (act == FS2K_ACT_RELU) return fmaxf(v, 0.f);
    if (act == FS2K_ACT_SILU) return silu(v);
    if (act == FS2K_ACT_TANH) return tanhf(v);
    return v;
}
__device__ __forceinline__ uint2 hb_pack4(const float4& v) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<const uint32_t*>(&lo);
    r.y = *reinterpret_cast<const uint32_t*>(&hi);
    return r;
}
// hi = bf16(v); lo = bf16(v − hi)
__device__ __forceinline__ void hb_split4(const float4& v, uint2& hi, uint2& lo) {
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
    const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
    hi.x = *reinterpret_cast<const uint32_t*>(&h0); hi.y = *reinterpret_cast<const uint32_t*>(&h1);
    lo.x = *reinterpret_cast<const uint32_t*>(&l0); lo.y = *reinterpret_cast<const uint32_t*>(&l1);
}


// byte offset of the 8-byte group holding columns [c, c+4) of row r inside a tile of 128-byte rows (64 bf16 per row,
// 16-byte chunks XOR-swizzled with the row index — what TMA SWIZZLE_128B writes and the descriptors above read);
// 64-column blocks are `block_bytes` apart
__device__ __forceinline__ uint32_t hb_tile_off(int r, int c, uint32_t block_bytes) {
    return (uint32_t)(c >> 6) * block_bytes + (uint32_t)r * 128u + (uint32_t)((((c & 63) >> 3) ^ (r & 7)) << 4) + (uint32_t)((c >> 2) & 1) * 8u;
}

}  // namespace fs2k
