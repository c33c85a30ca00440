// Backward kernels of the memory-bound ops: depthwise conv (+GLU), Linear(D→1), length regulator
// (segment sum), embedding scatter-adds, broadcast-row sums, and the aligner scores.
#include "common.cuh"

namespace fs2k {

constexpr int kDwbTile = 32;  // 16 was measured slower (0.91 vs 0.79 ms per step: twice the dw/db atomics and halo loads)
constexpr int kDwbUnroll = 4;

// gz [B,L,C] → dx (GLU: [B,L,2C] = (d value, d gate)), dw [C][K] (+=), dbias [C] (+=)
template <int K, bool GLU>
__global__ void __launch_bounds__(256)
dwconv_bwd_kernel(const float* __restrict__ gz, const float* __restrict__ x, int ldx, int L, int C,
                  const float* __restrict__ w, float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ dbias) {
    pdl_prologue();
    constexpr int P = (K - 1) / 2;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.z;
    const int l0 = blockIdx.y * kDwbTile;
    if (c >= C) return;
    float wk[K], dwk[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { wk[k] = w[(size_t)c * K + k]; dwk[k] = 0.f; }
    float db = 0.f;
    const float* xb = x + (size_t)b * L * ldx;
    const float* gb = gz + (size_t)b * L * C;
    auto fetch_a = [&](int l) -> float {
        if (l < 0 || l >= L) return 0.f;
        const float a = xb[(size_t)l * ldx + c];
        if (!GLU) return a;
        return a / (1.0f + expf(-xb[(size_t)l * ldx + C + c]));
    };
    auto fetch_g = [&](int l) -> float { return (l < 0 || l >= L) ? 0.f : gb[(size_t)l * C + c]; };
    float gwin[K], awin[K];  // values at positions l−P … l+P
#pragma unroll
    for (int k = 0; k < K - 1; ++k) { gwin[k + 1] = fetch_g(l0 - P + k); awin[k + 1] = fetch_a(l0 - P + k); }
    const int l_end = min(l0 + kDwbTile, L);
    // kDwbUnroll frames per trip: all their loads are issued before any is consumed (a one-load-per-step loop exposes
    // the L2 latency on every frame)
    for (int l = l0; l < l_end; l += kDwbUnroll) {
        float ng[kDwbUnroll], na[kDwbUnroll], vv[kDwbUnroll], gt[kDwbUnroll];
#pragma unroll
        for (int u = 0; u < kDwbUnroll; ++u) {
            const bool ok = l + u < l_end;
            ng[u] = ok ? fetch_g(l + u + P) : 0.f;
            na[u] = ok ? fetch_a(l + u + P) : 0.f;
            vv[u] = gt[u] = 0.f;
            if (GLU && ok) {
                vv[u] = xb[(size_t)(l + u) * ldx + c];
                gt[u] = xb[(size_t)(l + u) * ldx + C + c];
            }
        }
#pragma unroll
        for (int u = 0; u < kDwbUnroll; ++u) {
            if (l + u >= l_end) break;  // CTA-uniform
#pragma unroll
            for (int k = 0; k < K - 1; ++k) { gwin[k] = gwin[k + 1]; awin[k] = awin[k + 1]; }
            gwin[K - 1] = ng[u];
            awin[K - 1] = na[u];
            float da = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                da = fmaf(gwin[k], wk[K - 1 - k], da);        // dL/da[l] = Σ_k gz[l+P−k]·w[k]
                dwk[k] = fmaf(gwin[P], awin[k], dwk[k]);      // dL/dw[k] += gz[l]·a[l+k−P]
            }
            db += gwin[P];
            const size_t o = ((size_t)b * L + l + u) * ldx;
            if (GLU) {
                const float sg = 1.0f / (1.0f + expf(-gt[u]));
                dx[o + c] = da * sg;
                dx[o + C + c] = da * vv[u] * sg * (1.0f - sg);
            } else {
                dx[o + c] = da;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) atomicAdd(&dw[(size_t)c * K + k], dwk[k]);
    if (dbias) atomicAdd(&dbias[c], db);
}

// y[m] = (x[m]·w + b)·mask[m]  ⇒  dx[m,:] = g[m]·mask·w ; dw += Σ g·mask·x[m,:] ; db += Σ g·mask
__global__ void __launch_bounds__(256)
rowdot_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ w,
                  const uint8_t* __restrict__ mask, long M, int D, float* __restrict__ dx, float* __restrict__ dw,
                  float* __restrict__ db) {
    pdl_prologue();
    __shared__ float s_red[8][1024];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D4 = D >> 2;
    float4 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    float accb = 0.f;
    for (long m = (long)blockIdx.x * 8 + warp; m < M; m += (long)gridDim.x * 8) {
        const float gm = g[m] * (mask ? (mask[m] ? 1.f : 0.f) : 1.f);
        accb += gm;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int q = lane + 32 * i;
            if (q < D4) {
                const float4 xv = reinterpret_cast<const float4*>(x + (size_t)m * D)[q];
                const float4 wv = __ldg(reinterpret_cast<const float4*>(w) + q);
                reinterpret_cast<float4*>(dx + (size_t)m * D)[q] = make_float4(gm * wv.x, gm * wv.y, gm * wv.z, gm * wv.w);
                acc[i].x += gm * xv.x; acc[i].y += gm * xv.y; acc[i].z += gm * xv.z; acc[i].w += gm * xv.w;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int q = lane + 32 * i;
        if (q < D4) { float* p = &s_red[warp][q * 4]; p[0] = acc[i].x; p[1] = acc[i].y; p[2] = acc[i].z; p[3] = acc[i].w; }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) s += s_red[ww][c];
        atomicAdd(&dw[c], s);
    }
    if (lane == 0 && db) atomicAdd(db, accb);
}

// LengthRegulator backward: dx[b,t,:] = Σ_{f ∈ [cum[t−1], min(cum[t],F))} (g1[b,f,:] + g2[b,f,:])   one warp per (b,t)
__global__ void __launch_bounds__(256)
lr_bwd_kernel(const float* __restrict__ g1, const float* __restrict__ g2, const int* __restrict__ cum, int B, int T,
              int D, int F, float* __restrict__ dx) {
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    const int D4 = D >> 2;
    const long n = (long)B * T;
    for (long i = (long)blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += (long)gridDim.x * 8) {
        const int b = (int)(i / T), t = (int)(i % T);
        const int end = min(cum[i], F), start = min(t > 0 ? cum[i - 1] : 0, F);
        for (int q = lane; q < D4; q += 32) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int f = start; f < end; ++f) {
                if (g1) { const float4 a = reinterpret_cast<const float4*>(g1 + ((size_t)b * F + f) * D)[q]; s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w; }
                if (g2) { const float4 a = reinterpret_cast<const float4*>(g2 + ((size_t)b * F + f) * D)[q]; s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w; }
            }
            reinterpret_cast<float4*>(dx + (size_t)i * D)[q] = s;
        }
    }
}

// dtable[ids[n],:] += g1[n,:] (+ g2[n,:]) ; rows with id == skip_id receive nothing (nn.Embedding padding_idx)
template <typename IdT>
__global__ void __launch_bounds__(256)
scatter_add_rows_kernel(const float* __restrict__ g1, const float* __restrict__ g2, const IdT* __restrict__ ids, long N,
                        int D, long skip_id, float* __restrict__ dtable) {
    pdl_prologue();
    const long total = N * D;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long n = i / D;
        const long id = (long)ids[n];
        if (id == skip_id) continue;
        float v = g1[i];
        if (g2) v += g2[i];
        atomicAdd(&dtable[(size_t)id * D + (i % D)], v);
    }
}

// drows[ids ? ids[b] : b, :] += Σ_l g[b,l,:]
__global__ void __launch_bounds__(256)
rows_sum_scatter_kernel(const float* __restrict__ g, const int* __restrict__ ids, int B, int L, int D,
                        float* __restrict__ drows) {
    pdl_prologue();
    const int b = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= D) return;
    float s = 0.f;
    for (int l = 0; l < L; ++l) s += g[((size_t)b * L + l) * D + c];
    atomicAdd(&drows[(size_t)(ids ? ids[b] : b) * D + c], s);
}

// ---- aligner backward, stage 1: per (b,f) row, gradient w.r.t. the raw scores d[b,f,:] ----
__global__ void __launch_bounds__(128)
aligner_bwd_rows_kernel(const float* __restrict__ g_soft, const float* __restrict__ g_logprob,
                        const float* __restrict__ soft, const float* __restrict__ logprob,
                        const float* __restrict__ prior, const int* __restrict__ key_lens, int B, int F, int T,
                        float* __restrict__ dd) {
    pdl_prologue();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long n_rows = (long)B * F;
    for (long row = (long)blockIdx.x * 4 + warp; row < n_rows; row += (long)gridDim.x * 4) {
        const int b = (int)(row / F);
        const int klen = key_lens ? min(key_lens[b], T) : T;
        const size_t o = (size_t)row * T;
        float dot = 0.f;
        if (g_soft)
            for (int t = lane; t < klen; t += 32) dot += g_soft[o + t] * soft[o + t];
        dot = warp_sum(dot);
        float tot = 0.f;
        for (int t = lane; t < T; t += 32) {
            float glp = g_logprob ? g_logprob[o + t] : 0.f;
            if (g_soft && t < klen) glp += soft[o + t] * (g_soft[o + t] - dot);
            dd[o + t] = glp;
            tot += glp;
        }
        tot = warp_sum(tot);
        if (prior) {
            __syncwarp();
            for (int t = lane; t < T; t += 32) {
                const float pd = expf(logprob[o + t] - logf(prior[o + t] + 1e-8f));  // softmax of the raw scores
                dd[o + t] -= pd * tot;
            }
        }
    }
}

// out[b,i,:] = coef·(Σ_j W(i,j)·V[b,j,:] − U[b,i,:]·Σ_j W(i,j)),   W(i,j) = TRANS ? Wm[b,j,i] : Wm[b,i,j]
// C ≤ 80 channels; CTA = 64 rows i, 256 threads as 16×16 (4 rows × 5 channels each)
template <bool TRANS>
__global__ void __launch_bounds__(256)
aligner_bwd_proj_kernel(const float* __restrict__ Wm, const float* __restrict__ V, const float* __restrict__ U, int I,
                        int J, int C, float coef, float* __restrict__ out) {
    pdl_prologue();
    __shared__ float Ws[64][65];
    __shared__ float Vs[64][81];
    const int b = blockIdx.y, i0 = blockIdx.x * 64;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[4][5], rs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 5; ++c) acc[a][c] = 0.f;
    const float* Wb = Wm + (size_t)b * I * J;
    for (int j0 = 0; j0 < J; j0 += 64) {
        __syncthreads();
        for (int e = tid; e < 64 * 64; e += 256) {
            float v = 0.f;
            if (!TRANS) {
                const int ii = e >> 6, jj = e & 63;
                if (i0 + ii < I && j0 + jj < J) v = Wb[(size_t)(i0 + ii) * J + j0 + jj];
                Ws[ii][jj] = v;
            } else {
                const int jj = e >> 6, ii = e & 63;
                if (i0 + ii < I && j0 + jj < J) v = Wb[(size_t)(j0 + jj) * I + i0 + ii];
                Ws[ii][jj] = v;
            }
        }
        for (int e = tid; e < 64 * C; e += 256) {
            const int jj = e / C, c = e % C;
            Vs[jj][c] = (j0 + jj < J) ? V[((size_t)b * J + j0 + jj) * C + c] : 0.f;
        }
        __syncthreads();
        for (int jj = 0; jj < 64; ++jj) {
            float wv[4], vv[5];
#pragma unroll
            for (int a = 0; a < 4; ++a) { wv[a] = Ws[ty * 4 + a][jj]; rs[a] += wv[a]; }
#pragma unroll
            for (int c = 0; c < 5; ++c) vv[c] = (tx + 16 * c < C) ? Vs[jj][tx + 16 * c] : 0.f;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 5; ++c) acc[a][c] = fmaf(wv[a], vv[c], acc[a][c]);
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i = i0 + ty * 4 + a;
        if (i >= I) continue;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const int ch = tx + 16 * c;
            if (ch < C) {
                const size_t o = ((size_t)b * I + i) * C + ch;
                out[o] = coef * (acc[a][c] - U[o] * rs[a]);
            }
        }
    }
}

}  // namespace fs2k

using namespace fs2k;

static inline int ew_grid3(long n, int per = 256) {
    long g = (n + per - 1) / per;
    if (g > 148 * 8) g = 148 * 8;
    return (int)(g < 1 ? 1 : g);
}

template <int K>
static int launch_dwb(const float* gz, const float* x, int ldx, int B, int L, int C, const float* w, int glu, float* dx,
                      float* dw, float* dbias, cudaStream_t s) {
    const int threads = C < 256 ? ((C + 31) / 32) * 32 : 256;
    dim3 grid(cdiv(C, threads), cdiv(L, kDwbTile), B);
    if (glu) fs2k_launch(dwconv_bwd_kernel<K, true>, dim3(grid), dim3(threads), 0, s, gz, x, ldx, L, C, w, dx, dw, dbias);
    else fs2k_launch(dwconv_bwd_kernel<K, false>, dim3(grid), dim3(threads), 0, s, gz, x, ldx, L, C, w, dx, dw, dbias);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_dwconv_bwd(const float* gz, const float* x, int ldx, int B, int L, int C, const float* w, int K,
                               int glu, float* dx, float* dw, float* dbias, int accumulate, fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && L >= 0 && C > 0 && K > 0 && (K & 1), FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(gz && x && w && dx && dw, FS2K_ERR_NULL);
    FS2K_REQUIRE(B <= 65535, FS2K_ERR_UNSUPPORTED);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
    if (!accumulate) {  // otherwise the atomics add on top of what dw / dbias already hold
        e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)C * K, s);
        if (e == cudaSuccess && dbias) e = cudaMemsetAsync(dbias, 0, sizeof(float) * C, s);
    }
    if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    if (B == 0 || L == 0) return FS2K_OK;
    switch (K) {
        case 3: return launch_dwb<3>(gz, x, ldx, B, L, C, w, glu, dx, dw, dbias, s);
        case 5: return launch_dwb<5>(gz, x, ldx, B, L, C, w, glu, dx, dw, dbias, s);
        case 7: return launch_dwb<7>(gz, x, ldx, B, L, C, w, glu, dx, dw, dbias, s);
        case 9: return launch_dwb<9>(gz, x, ldx, B, L, C, w, glu, dx, dw, dbias, s);
        case 11: return launch_dwb<11>(gz, x, ldx, B, L, C, w, glu, dx, dw, dbias, s);
        case 15: return launch_dwb<15>(gz, x, ldx, B, L, C, w, glu, dx, dw, dbias, s);
        default: return FS2K_ERR_UNSUPPORTED;
    }
}

extern "C" int fs2k_rowdot_bwd(const float* g, const float* x, const float* w, const uint8_t* mask, long M, int D,
                               float* dx, float* dw, float* db, fs2k_stream_t stream) {
    FS2K_REQUIRE(M >= 0 && D > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE((D & 3) == 0 && D <= 1024, FS2K_ERR_UNSUPPORTED);
    FS2K_REQUIRE(g && x && w && dx && dw, FS2K_ERR_NULL);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * D, s);
    if (e == cudaSuccess && db) e = cudaMemsetAsync(db, 0, sizeof(float), s);
    if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    if (M == 0) return FS2K_OK;
    long grid = (M + 63) / 64;
    if (grid > 148 * 2) grid = 148 * 2;
    fs2k_launch(rowdot_bwd_kernel, dim3((int)grid), dim3(256), 0, s, g, x, w, mask, M, D, dx, dw, db);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_lr_bwd(const float* g_out, const float* g_out_pos, const int* cum, int B, int T, int D, int F,
                           float* dx, fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && T >= 0 && D > 0 && F >= 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE((D & 3) == 0, FS2K_ERR_UNSUPPORTED);
    if ((long)B * T == 0) return FS2K_OK;
    FS2K_REQUIRE(cum && dx, FS2K_ERR_NULL);
    fs2k_launch(lr_bwd_kernel, dim3(ew_grid3((long)B * T, 8)), dim3(256), 0, (cudaStream_t)stream, g_out, g_out_pos, cum, B, T, D, F, dx);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_scatter_add_rows(const float* g1, const float* g2, const void* ids, int ids_are_int64, long N, int D,
                                     long skip_id, float* dtable, fs2k_stream_t stream) {
    FS2K_REQUIRE(N >= 0 && D > 0, FS2K_ERR_BAD_SHAPE);
    if (N == 0) return FS2K_OK;
    FS2K_REQUIRE(g1 && ids && dtable, FS2K_ERR_NULL);
    cudaStream_t s = (cudaStream_t)stream;
    if (ids_are_int64) fs2k_launch(scatter_add_rows_kernel<long long>, dim3(ew_grid3(N * D)), dim3(256), 0, s, g1, g2, (const long long*)ids, N, D, skip_id, dtable);
    else fs2k_launch(scatter_add_rows_kernel<int>, dim3(ew_grid3(N * D)), dim3(256), 0, s, g1, g2, (const int*)ids, N, D, skip_id, dtable);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_rows_sum_scatter(const float* g, const int* ids, int B, int L, int D, float* drows,
                                     fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && L >= 0 && D > 0, FS2K_ERR_BAD_SHAPE);
    if (B == 0) return FS2K_OK;
    FS2K_REQUIRE(g && drows, FS2K_ERR_NULL);
    FS2K_REQUIRE(B <= 65535, FS2K_ERR_UNSUPPORTED);
    fs2k_launch(rows_sum_scatter_kernel, dim3(dim3(cdiv(D, 128), B)), dim3(128), 0, (cudaStream_t)stream, g, ids, B, L, D, drows);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_aligner_bwd(const float* g_soft, const float* g_logprob, const float* soft, const float* logprob,
                                const float* prior, const int* key_lens, const float* q, const float* k, int B, int F,
                                int T, int C, float* dd /* [B,F,T] scratch */, float* dq, float* dk,
                                fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && F >= 0 && T >= 0 && C > 0, FS2K_ERR_BAD_SHAPE);
    FS2K_REQUIRE(C <= 80, FS2K_ERR_UNSUPPORTED);
    if (B == 0 || F == 0 || T == 0) return FS2K_OK;
    FS2K_REQUIRE(soft && logprob && q && k && dd && dq && dk && (g_soft || g_logprob), FS2K_ERR_NULL);
    FS2K_REQUIRE(B <= 65535, FS2K_ERR_UNSUPPORTED);
    cudaStream_t s = (cudaStream_t)stream;
    long g = ((long)B * F + 3) / 4;
    if (g > 148 * 16) g = 148 * 16;
    fs2k_launch(aligner_bwd_rows_kernel, dim3((int)g), dim3(128), 0, s, g_soft, g_logprob, soft, logprob, prior, key_lens, B, F, T, dd);
    FS2K_CHECK_LAUNCH();
    fs2k_launch(aligner_bwd_proj_kernel<false>, dim3(dim3(cdiv(F, 64), B)), dim3(256), 0, s, dd, k, q, F, T, C, 0.001f, dq);
    FS2K_CHECK_LAUNCH();
    fs2k_launch(aligner_bwd_proj_kernel<true>, dim3(dim3(cdiv(T, 64), B)), dim3(256), 0, s, dd, q, k, T, F, C, 0.001f, dk);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}
