// Masked multi-head self-attention, flash style (no L×L matrix in memory), exact fp32.
//
//   nn.MultiheadAttention(q=k=v=LN(x), key_padding_mask, need_weights=False)
//   torchaudio conformer.py:151-153,193-202:  softmax(Q Kᵀ/√hd + (−inf on padded keys)) V
// Inputs are the packed in_proj output qkv[B,L,3D] (q | k | v, heads contiguous inside each third);
// output o[B,L,D] feeds out_proj.  Padded *queries* are computed like valid ones (their outputs are
// live, SURVEY §8a note P); key tiles entirely beyond lens[b] are skipped (they contribute exp(−inf)=0).
//
// CTA = 64 queries of one (b, head), 256 threads as a 16×16 grid: thread (ty,tx) owns queries
// ty·4..+3; keys tx+16j of the current 32-key tile; output dims tx·4..+3 (+64). Online softmax
// statistics live in registers and are reduced across the 16 tx lanes with shuffles.
#include "common.cuh"

namespace fs2k {

constexpr int ABQ = 64, ABK = 32;

template <int HD>
__global__ void __launch_bounds__(256)
attention_simt_kernel(const float* __restrict__ qkv, const int* __restrict__ lens, int L, int H, float scale,
                      float drop_p, unsigned long long seed, const int* __restrict__ order, float* __restrict__ out,
                      float* __restrict__ lse_out) {
    pdl_prologue();
    seed = seed_with_base(seed);
    constexpr int QS = HD + 4, PS = ABK + 4, NJ = HD / 64;
    extern __shared__ __align__(16) float smem[];
    float* Qs = smem;                 // [ABQ][QS]
    float* Ks = Qs + ABQ * QS;        // [ABK][QS]
    float* Vs = Ks + ABK * QS;        // [ABK][HD]
    float* Ps = Vs + ABK * HD;        // [ABQ][PS]
    // CTAs are dispatched in block-index order: `order` lists the utterances longest first (work ∝ lens[b]), so the
    // short ones fill the tail of the last wave instead of leaving SMs idle behind a long one
    const int b = order ? order[blockIdx.z] : blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * ABQ;
    const int D = H * HD, ld = 3 * D;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int len = min(lens[b], L);
    const float inv_keep = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.f;
    const float* base = qkv + (size_t)b * L * ld + h * HD;

    // stage the query tile (rows past L are zero)
    for (int i = tid; i < ABQ * (HD / 4); i += 256) {
        const int r = i / (HD / 4), c = i % (HD / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q0 + r < L) v = *reinterpret_cast<const float4*>(base + (size_t)(q0 + r) * ld + c * 4);
        *reinterpret_cast<float4*>(Qs + r * QS + c * 4) = v;
    }
    float m_run[4], l_run[4];
    float4 o[4][NJ];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m_run[i] = -INFINITY;
        l_run[i] = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) o[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int n_tiles = (len + ABK - 1) / ABK;
    for (int kt = 0; kt < n_tiles; ++kt) {
        const int k0 = kt * ABK;
        __syncthreads();  // previous tile fully consumed (and Qs visible on the first pass)
        for (int i = tid; i < ABK * (HD / 4); i += 256) {
            const int r = i / (HD / 4), c = i % (HD / 4);
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (k0 + r < L) {
                const float* p = base + (size_t)(k0 + r) * ld + c * 4;
                kv = *reinterpret_cast<const float4*>(p + D);
                vv = *reinterpret_cast<const float4*>(p + 2 * D);
            }
            *reinterpret_cast<float4*>(Ks + r * QS + c * 4) = kv;
            *reinterpret_cast<float4*>(Vs + r * HD + c * 4) = vv;
        }
        __syncthreads();
        // S = Q Kᵀ (4 queries × 2 keys per thread)
        float s[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i][0] = s[i][1] = 0.f;
#pragma unroll 8
        for (int c = 0; c < HD / 4; ++c) {
            float4 q4[4], k4[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) q4[i] = *reinterpret_cast<const float4*>(Qs + (ty * 4 + i) * QS + c * 4);
#pragma unroll
            for (int j = 0; j < 2; ++j) k4[j] = *reinterpret_cast<const float4*>(Ks + (tx + 16 * j) * QS + c * 4);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    s[i][j] = fmaf(q4[i].x, k4[j].x, s[i][j]);
                    s[i][j] = fmaf(q4[i].y, k4[j].y, s[i][j]);
                    s[i][j] = fmaf(q4[i].z, k4[j].z, s[i][j]);
                    s[i][j] = fmaf(q4[i].w, k4[j].w, s[i][j]);
                }
        }
        // online softmax
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int key = k0 + tx + 16 * j;
                s[i][j] = key < len ? s[i][j] * scale : -INFINITY;
                mx = fmaxf(mx, s[i][j]);
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            const float m_new = fmaxf(m_run[i], mx);
            const float m_use = m_new == -INFINITY ? 0.f : m_new;
            const float corr = expf(m_run[i] - m_use);
            float rs = 0.f;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float p = expf(s[i][j] - m_use);
                rs += p;  // the softmax normaliser uses the un-dropped probabilities
                Ps[(ty * 4 + i) * PS + tx + 16 * j] =
                    p * attn_keep(seed, drop_p, inv_keep, b, h, q0 + ty * 4 + i, k0 + tx + 16 * j, H, L);
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
            l_run[i] = l_run[i] * corr + rs;
            m_run[i] = m_new;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                o[i][j].x *= corr; o[i][j].y *= corr; o[i][j].z *= corr; o[i][j].w *= corr;
            }
        }
        __syncthreads();
        // O += P V
#pragma unroll 2
        for (int kk = 0; kk < ABK; kk += 4) {
            float4 p4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) p4[i] = *reinterpret_cast<const float4*>(Ps + (ty * 4 + i) * PS + kk);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const float4 v = *reinterpret_cast<const float4*>(Vs + (kk + u) * HD + tx * 4 + 64 * j);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float p = u == 0 ? p4[i].x : (u == 1 ? p4[i].y : (u == 2 ? p4[i].z : p4[i].w));
                        o[i][j].x = fmaf(p, v.x, o[i][j].x);
                        o[i][j].y = fmaf(p, v.y, o[i][j].y);
                        o[i][j].z = fmaf(p, v.z, o[i][j].z);
                        o[i][j].w = fmaf(p, v.w, o[i][j].w);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = q0 + ty * 4 + i;
        if (q >= L) continue;
        const float inv = 1.0f / l_run[i];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            float4 v = o[i][j];
            v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
            *reinterpret_cast<float4*>(out + ((size_t)b * L + q) * D + h * HD + tx * 4 + 64 * j) = v;
        }
        if (lse_out && tx == 0) lse_out[((size_t)b * H + h) * L + q] = m_run[i] + logf(l_run[i]);
    }
}

// order[rank] = utterance index, ranks by descending length (ties: lower index first).  One thread per utterance, O(B²).
__global__ void attn_order_kernel(const int* __restrict__ lens, int B, int* __restrict__ order) {
    pdl_prologue();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        const int li = lens[i];
        int rank = 0;
        for (int j = 0; j < B; ++j) {
            const int lj = lens[j];
            rank += (lj > li) || (lj == li && j < i);
        }
        order[rank] = i;
    }
}

}  // namespace fs2k

using namespace fs2k;

// fills order_ws[B] (longest utterance first) for the attention kernels; B ≤ 4096
extern "C" int fs2k_attention_order(const int* lens, int B, int* order_ws, fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && B <= 4096, FS2K_ERR_BAD_SHAPE);
    if (B == 0) return FS2K_OK;
    FS2K_REQUIRE(lens && order_ws, FS2K_ERR_NULL);
    fs2k_launch(attn_order_kernel, dim3(cdiv(B, 128)), dim3(128), 0, (cudaStream_t)stream, lens, B, order_ws);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_attention_f32(const float* qkv, const int* lens, int B, int L, int H, int head_dim, float dropout_p,
                                  long seed, float* out, float* lse_out, const int* order, fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && L >= 0 && H > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(head_dim == 64 || head_dim == 128, FS2K_ERR_UNSUPPORTED);
    FS2K_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, FS2K_ERR_BAD_SHAPE);
    if (B == 0 || L == 0) return FS2K_OK;
    FS2K_REQUIRE(qkv && lens && out, FS2K_ERR_NULL);
    const float scale = 1.0f / sqrtf((float)head_dim);
    dim3 grid(cdiv(L, ABQ), H, B);
    cudaStream_t s = (cudaStream_t)stream;
    if (head_dim == 128) {
        const int smem = (ABQ * 132 + ABK * 132 + ABK * 128 + ABQ * (ABK + 4)) * 4;
        static bool set = false;
        if (!set) {
            cudaError_t e = cudaFuncSetAttribute(attention_simt_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return fs2k_set_cuda_error(e);
            set = true;
        }
        fs2k_launch(attention_simt_kernel<128>, dim3(grid), dim3(256), smem, s, qkv, lens, L, H, scale, dropout_p, (unsigned long long)seed, order, out, lse_out);
    } else {
        const int smem = (ABQ * 68 + ABK * 68 + ABK * 64 + ABQ * (ABK + 4)) * 4;
        fs2k_launch(attention_simt_kernel<64>, dim3(grid), dim3(256), smem, s, qkv, lens, L, H, scale, dropout_p, (unsigned long long)seed, order, out, lse_out);
    }
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

FS2K_DEFINE_SEED_BASE_SETTER(attention)
