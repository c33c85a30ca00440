// Batch builder and prediction trimming on the device.
//   unpack_ragged  — the zero-padding half of FastSpeech2DataModule.collate_method (fs2/dataset.py:270-287): the host
//                    packs only the VALID values of every item back to back into one pinned staging buffer (one
//                    H2D copy per batch); this kernel scatters them into the padded [B, Rmax, Cmax] tensor and
//                    writes the zeros (pad_sequence, and the two-sided padding of the [F,T] attention prior).
//   trim_transpose — prediction_writing_callback.py:255-262: the valid frames of every utterance, transposed to
//                    [n_mels, T_b] and packed back to back, ready for one D2H copy.
#include "common.cuh"

namespace fs2k {

// 4-byte words.  item b: rows[b] × cols[b] words at packed + off[b]; out [B][Rmax][Cmax] words.
__global__ void __launch_bounds__(256)
unpack_ragged_kernel(const uint32_t* __restrict__ packed, const long long* __restrict__ off, const int* __restrict__ rows,
                     const int* __restrict__ cols, int B, int Rmax, int Cmax, uint32_t* __restrict__ out) {
    pdl_prologue();
    const long total = (long)B * Rmax * Cmax;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cmax);
        const int r = (int)((i / Cmax) % Rmax);
        const int b = (int)(i / ((long)Cmax * Rmax));
        const int cb = cols[b];
        out[i] = (r < rows[b] && c < cb) ? packed[off[b] + (long long)r * cb + c] : 0u;
    }
}

// mel [B,F,C] → out[out_off[b] + c·len_b + t] = mel[b,t,c], t < len_b = min(lens[b], F).  32×32 smem transpose tiles.
__global__ void __launch_bounds__(256)
trim_transpose_kernel(const float* __restrict__ mel, const int* __restrict__ lens, const long long* __restrict__ out_off,
                      int F, int C, float* __restrict__ out) {
    pdl_prologue();
    __shared__ float tile[32][33];
    const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int len = min(lens[b], F);
    if (t0 >= len) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 × 8
    for (int r = ty; r < 32; r += 8) {
        const int t = t0 + r, c = c0 + tx;
        tile[r][tx] = (t < len && c < C) ? mel[((size_t)b * F + t) * C + c] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, t = t0 + tx;
        if (c < C && t < len) out[out_off[b] + (long long)c * len + t] = tile[tx][r];
    }
}

}  // namespace fs2k

using namespace fs2k;

extern "C" int fs2k_unpack_ragged(const void* packed_words, const long long* word_offsets, const int* rows, const int* cols,
                                  int B, int Rmax, int Cmax, void* out_words, fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && Rmax >= 0 && Cmax >= 0, FS2K_ERR_BAD_SHAPE);
    const long total = (long)B * Rmax * Cmax;
    if (total == 0) return FS2K_OK;
    FS2K_REQUIRE(packed_words && word_offsets && rows && cols && out_words, FS2K_ERR_NULL);
    long g = (total + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    fs2k_launch(unpack_ragged_kernel, dim3((unsigned)g), dim3(256), 0, (cudaStream_t)stream, (const uint32_t*)packed_words,
                word_offsets, rows, cols, B, Rmax, Cmax, (uint32_t*)out_words);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}

extern "C" int fs2k_trim_transpose(const float* mel, const int* lens, const long long* out_offsets, int B, int F, int C,
                                   float* out, fs2k_stream_t stream) {
    FS2K_REQUIRE(B >= 0 && F >= 0 && C > 0, FS2K_ERR_BAD_SHAPE);
    if (B == 0 || F == 0) return FS2K_OK;
    FS2K_REQUIRE(mel && lens && out_offsets && out, FS2K_ERR_NULL);
    FS2K_REQUIRE(B <= 65535, FS2K_ERR_UNSUPPORTED);
    fs2k_launch(trim_transpose_kernel, dim3(cdiv(F, 32), cdiv(C, 32), B), dim3(256), 0, (cudaStream_t)stream, mel, lens,
                out_offsets, F, C, out);
    FS2K_CHECK_LAUNCH();
    return FS2K_OK;
}
