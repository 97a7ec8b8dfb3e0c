// Host-side plumbing shared by the launchers: error slot, device checks, launch counter.
#pragma once
#include "../../include/gd_b200.h"
#include <cuda.h>
#include <cuda_runtime.h>

namespace gd {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int set_error(int code, const char* fmt, ...);  // returns code
PFN_encodeTiled get_encode_tiled();             // resolved through cudaGetDriverEntryPoint (no -lcuda link)
constexpr int GD_MAX_DEVICES = 64;
int current_device();                           // cudaGetDevice(), clamped to [0, GD_MAX_DEVICES)
int sm_count();                                 // SMs of the current device (148 on B200), cached per device
int check_device();                             // GD_OK iff current device is compute capability 10.x
void count_launch();
int validate_ddpm(const gd_ddpm_desc* u);
// 2-D row-major tensor map: box = box_rows x box_cols elements, innermost dimension = columns (gemm_tcgen05.cu)
int make_tmap_2d(CUtensorMap* m, CUtensorMapDataType dt, int elt_bytes, const void* base, uint64_t rows, uint64_t cols,
                 uint64_t ld_elems, uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle swz);

// Launch configuration with the library-wide attributes: optional cluster width and programmatic stream serialization
// (GD_PDL=0 in the environment turns the latter off).
struct LaunchCfg {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attrs[2];
};
void fill_launch(LaunchCfg& L, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x = 1);
// Names the kernel family of the launches issued while it is alive ("gemm", "ddpm", "ln", "attn", "scatter", "step",
// "pack", "speech"); GD_PDL_OFF=scatter,step (comma list) launches those families WITHOUT programmatic stream
// serialization - the bisect switch of profiles/determinism_bisect.py.
struct KindScope {
    explicit KindScope(const char* kind);
    ~KindScope();
    const char* prev;
};

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                            Args&&... args) {
    LaunchCfg L;
    fill_launch(L, grid, block, smem, stream, cluster_x);
    return cudaLaunchKernelEx(&L.cfg, kern, static_cast<Args&&>(args)...);
}

#define GD_CUDA_CHECK(expr)                                                                          \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return ::gd::set_error(GD_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                   __FILE__, __LINE__);                                              \
    } while (0)

}  // namespace gd
