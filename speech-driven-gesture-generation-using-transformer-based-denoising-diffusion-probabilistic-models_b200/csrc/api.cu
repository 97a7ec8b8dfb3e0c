// C-ABI plumbing: error slot, device checks, launch counter, descriptor validation.
#include "host_util.h"
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace gd {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

PFN_encodeTiled get_encode_tiled() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    return reinterpret_cast<PFN_encodeTiled>(fn);
}

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    return dev < GD_MAX_DEVICES ? dev : GD_MAX_DEVICES - 1;
}

int sm_count() {
    static int n_dev[GD_MAX_DEVICES] = {};
    int& n = n_dev[current_device()];
    if (!n) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, current_device());
        if (n <= 0) n = 148;
    }
    return n;
}

int check_device() {
    static int ok_dev[GD_MAX_DEVICES];
    static bool init = false;
    if (!init) {
        for (int i = 0; i < GD_MAX_DEVICES; ++i) ok_dev[i] = -1;
        init = true;
    }
    int& ok = ok_dev[current_device()];
    if (ok < 0) {
        int dev = 0, major = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
            return set_error(GD_ERR_CUDA, "no CUDA device available");
        ok = (major == 10) ? 1 : 0;
    }
    if (!ok) return set_error(GD_ERR_ARCH, "these kernels are built for sm_100a (B200) only");
    return GD_OK;
}

static thread_local const char* g_kind = "";
KindScope::KindScope(const char* kind) : prev(g_kind) { g_kind = kind; }
KindScope::~KindScope() { g_kind = prev; }

static bool pdl_enabled() {
    const char* e = getenv("GD_PDL");
    if (e && e[0] == '0') return false;
    const char* off = getenv("GD_PDL_OFF");  // comma-separated kernel families launched without early start
    if (off && g_kind[0]) {
        const size_t n = strlen(g_kind);
        for (const char* q = off; *q;) {
            const char* c = strchr(q, ',');
            const size_t len = c ? (size_t)(c - q) : strlen(q);
            if (len == n && strncmp(q, g_kind, n) == 0) return false;
            q += len + (c ? 1 : 0);
        }
    }
    return true;
}

void fill_launch(LaunchCfg& L, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x) {
    L.cfg = cudaLaunchConfig_t{};
    L.cfg.gridDim = grid, L.cfg.blockDim = block, L.cfg.dynamicSmemBytes = smem, L.cfg.stream = stream;
    int n = 0;
    if (cluster_x > 1) {
        L.attrs[n].id = cudaLaunchAttributeClusterDimension;
        L.attrs[n].val.clusterDim.x = cluster_x, L.attrs[n].val.clusterDim.y = 1, L.attrs[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl_enabled()) {
        L.attrs[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        L.attrs[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    L.cfg.attrs = L.attrs, L.cfg.numAttrs = n;
}

int validate_ddpm(const gd_ddpm_desc* u) {
    if (!u) return set_error(GD_ERR_INVALID, "ddpm: null descriptor");
    if (!u->x || !u->coef_A || !u->coef_B || !u->coef_C1 || !u->coef_C2 || !u->sigma || !u->step_ptr)
        return set_error(GD_ERR_INVALID, "ddpm: x / coefficient tables / step_ptr must be non-null");
    if (u->n_clips <= 0 || u->C <= 0 || u->T <= 0) return set_error(GD_ERR_INVALID, "ddpm: bad shape");
    if (u->inpaint_seed && (!u->inpaint_mask || !u->inpaint_factor))
        return set_error(GD_ERR_INVALID, "ddpm: inpaint_seed needs inpaint_mask and inpaint_factor");
    return GD_OK;
}

}  // namespace gd

#ifdef GD_TRACE
namespace gd {
void set_trace_gemm(unsigned long long* buf);
void set_trace_elementwise(unsigned long long* buf);
void set_trace_attention(unsigned long long* buf);
}  // namespace gd
// profiling builds only: buf[0] = slot counter, buf[1] = capacity, slots of 10 words from buf[16]
extern "C" int gd_debug_set_trace(void* buf) {
    gd::set_trace_gemm(static_cast<unsigned long long*>(buf));
    gd::set_trace_elementwise(static_cast<unsigned long long*>(buf));
    gd::set_trace_attention(static_cast<unsigned long long*>(buf));
    return cudaDeviceSynchronize() == cudaSuccess ? GD_OK : GD_ERR_CUDA;
}
#endif

extern "C" int gd_abi_version(void) { return GD_ABI_VERSION; }
extern "C" const char* gd_last_error(void) { return gd::g_err; }
extern "C" uint64_t gd_launch_count(void) { return gd::g_launches.load(std::memory_order_relaxed); }
