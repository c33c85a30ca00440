// libfs2k: error reporting and device capability probe.
#include "common.cuh"

static thread_local cudaError_t g_last_cuda_error = cudaSuccess;

int fs2k_set_cuda_error(cudaError_t e) {
    g_last_cuda_error = e;
    return FS2K_ERR_CUDA;
}

extern "C" const char* fs2k_strerror(int code) {
    switch (code) {
        case FS2K_OK: return "ok";
        case FS2K_ERR_BAD_SHAPE: return "bad shape (negative or inconsistent dimension)";
        case FS2K_ERR_UNSUPPORTED: return "unsupported dimension / option for this kernel";
        case FS2K_ERR_WORKSPACE: return "workspace too small";
        case FS2K_ERR_NULL: return "required pointer is null";
        case FS2K_ERR_CUDA: return cudaGetErrorString(g_last_cuda_error);
        case FS2K_ERR_ARCH: return "device is not compute capability 10.x (sm_100a build)";
        default: return "unknown fs2k error";
    }
}

int g_fs2k_pdl_enabled = 1;
extern "C" int fs2k_set_pdl(int enabled) {
    g_fs2k_pdl_enabled = enabled ? 1 : 0;
    return FS2K_OK;
}

extern "C" int fs2k_version(void) { return 100; }

// 0 when the current device can run this sm_100a-only build
extern "C" int fs2k_check_device(void) {
    int dev = 0, major = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return fs2k_set_cuda_error(e);
    return major == 10 ? FS2K_OK : FS2K_ERR_ARCH;
}

// every translation unit with dropout kernels keeps its own copy of the seed-base pointer (common.cuh)
extern "C" int fs2k_seed_base_set_attention_bwd(const void*);
extern "C" int fs2k_seed_base_set_attention(const void*);
extern "C" int fs2k_seed_base_set_dropout(const void*);
extern "C" int fs2k_seed_base_set_gemm_bwd(const void*);
extern "C" int fs2k_seed_base_set_norms(const void*);
extern "C" int fs2k_seed_base_set_norms_bwd(const void*);
extern "C" int fs2k_seed_base_set_gemm_tc(const void*);
extern "C" int fs2k_seed_base_set_gemm_bf16(const void*);
extern "C" int fs2k_seed_base_set_attention_tc(const void*);
extern "C" int fs2k_seed_base_set_gemm_bf16_panel(const void*);

extern "C" int fs2k_set_dropout_seed_base(const unsigned long long* device_counter) {
    int (*setters[])(const void*) = {fs2k_seed_base_set_attention_bwd, fs2k_seed_base_set_attention, fs2k_seed_base_set_dropout,
                                     fs2k_seed_base_set_gemm_bwd,      fs2k_seed_base_set_norms,     fs2k_seed_base_set_norms_bwd,
                                     fs2k_seed_base_set_gemm_tc,       fs2k_seed_base_set_gemm_bf16,
                                     fs2k_seed_base_set_attention_tc,  fs2k_seed_base_set_gemm_bf16_panel};
    for (auto f : setters) {
        const int r = f(device_counter);
        if (r != FS2K_OK) return r;
    }
    return FS2K_OK;
}
