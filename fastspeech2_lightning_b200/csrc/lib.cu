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
