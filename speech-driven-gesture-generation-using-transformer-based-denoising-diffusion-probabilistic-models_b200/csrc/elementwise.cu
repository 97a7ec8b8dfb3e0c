// HBM-bound row kernels of the sampling chain: LayerNorm (fp32 -> bf16 GEMM operand), the
// standalone DDPM update, the per-step timestep-row scatter and the layout/cast helpers.
// All of them are one-pass, 16-byte vectorised and coalesced along the contiguous axis.
#include "common.cuh"
#include "ddpm_math.cuh"
#include "host_util.h"

namespace gd {

// ------------------------------------------------------------------------------- LayerNorm
// One warp per row, the whole row lives in registers (D/32 floats per lane): one HBM read,
// one bf16 write, two shuffle reductions (mean, then centred second moment like ATen's
// two-pass/Welford result — not E[x^2]-E[x]^2, which loses bits when |mean| >> std).
template <int D, int R>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const float* __restrict__ x, int ldx,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta,
                                                             __nv_bfloat16* __restrict__ out, int ldo, int M,
                                                             float eps, const float* __restrict__ gamma2,
                                                             const float* __restrict__ beta2, int split_row) {
    const int trace_slot = threadIdx.x == 0 ? GD_TRACE_OPEN(200) : -1;
    GD_TRACE_MARK(trace_slot, 0);
    pdl_launch_dependents();
    pdl_wait();
    GD_TRACE_MARK(trace_slot, 2);
    constexpr int V = D / 128;  // float4 chunks per lane and row
    // a warp owns R consecutive rows and issues all their loads before the first reduction (more bytes in flight per warp)
    const int row0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * R;
    if (row0 >= M) return;
    const int lane = threadIdx.x & 31;
    float4 v[R][V];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int row = min(row0 + r, M - 1);
        const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * ldx);
#pragma unroll
        for (int i = 0; i < V; ++i) v[r][i] = __ldcg(xr + i * 32 + lane);  // L2-coherent: see common.cuh (PDL and L1)
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float mean, rstd;
        ln_row_stats<V>(v[r], eps, mean, rstd);
        if (row0 + r < M) {
            // rows >= split_row take the second parameter set (tedexp: pose rows -> norm_ff, memory rows -> norm_ff_mem)
            const bool second = gamma2 != nullptr && row0 + r >= split_row;
            const float4* g4 = reinterpret_cast<const float4*>(second ? gamma2 : gamma);
            const float4* b4 = reinterpret_cast<const float4*>(second ? beta2 : beta);
            uint2* orow = reinterpret_cast<uint2*>(out + (size_t)(row0 + r) * ldo);
#pragma unroll
            for (int i = 0; i < V; ++i)
                orow[i * 32 + lane] = ln_apply_pack(v[r][i], mean, rstd, __ldg(g4 + i * 32 + lane), __ldg(b4 + i * 32 + lane));
        }
    }
    GD_TRACE_MARK(trace_slot, 8);
}

// ------------------------------------------------------------------------------- DDPM update
// Standalone form (eps already in HBM, (N,C,T)).  One thread per (clip, channel<ld_xa, frame);
// frame is the fastest index so x/noise/eps accesses are coalesced.
__global__ void __launch_bounds__(256) ddpm_update_kernel(const gd_ddpm_desc u, const float* __restrict__ eps,
                                                          int c_span) {
    pdl_launch_dependents();
    pdl_wait();
    const int t = load_step(u.step_ptr);
    const DdpmStepCoefs cf = ddpm_load_coefs(u, t);
    const size_t total = (size_t)u.n_clips * c_span * u.T;
    const size_t tape_base = (size_t)t * u.n_clips * u.C * u.T;
    const bool inpaint = u.inpaint_seed != nullptr;
    bool aux = true;
    if (u.aux_step_ptr) {
        const int aux_t = load_step(u.aux_step_ptr);
        aux = aux_t < 0 || aux_t == t;
    }
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int frame = (int)(i % u.T);
        const int c = (int)((i / u.T) % c_span);
        const int clip = (int)(i / ((size_t)u.T * c_span));
        float xnext = 0.f, xadd = 0.f;
        if (c < u.C) {
            const size_t idx = ((size_t)clip * u.C + c) * u.T + frame;
            if (u.xa_add) xadd = __ldg(u.xa_add + idx);
            float m = 0.f, f = 0.f, seed = 0.f;
            if (inpaint) {
                m = __ldg(u.inpaint_mask + (size_t)clip * u.T + frame);
                f = __ldg(u.inpaint_factor + frame);
                seed = __ldg(u.inpaint_seed + ((size_t)clip * u.T + frame) * u.C + c);
            }
            const float z = u.noise_tape ? __ldg(u.noise_tape + tape_base + idx) : 0.f;
            const float e = __ldcg(eps + idx);
            float x0, mean, raw;
            xnext = ddpm_update_elem(cf, __ldcg(u.x + idx), e, z, inpaint, seed, m, f, u.clip_x0, &x0, &mean, &raw);
            u.x[idx] = xnext;
            if (aux) {
                if (u.eps_out) u.eps_out[idx] = e;
                if (u.x0_out) u.x0_out[idx] = x0;
                if (u.mean_out) u.mean_out[idx] = mean;
                if (u.raw_x0_out) u.raw_x0_out[idx] = raw;
            }
        }
        if (u.xa_bf16)
            reinterpret_cast<__nv_bfloat16*>(u.xa_bf16)[((size_t)clip * u.T + frame) * u.ld_xa + c] =
                __float2bfloat16_rn(xnext + xadd);
    }
}

// ------------------------------------------------------------------------------- step-row scatter
__global__ void __launch_bounds__(256) scatter_row_f32_kernel(float* __restrict__ dst, const float* __restrict__ init,
                                                              const float* __restrict__ table,
                                                              const int* __restrict__ step_ptr, int n_clips,
                                                              int rows_per_clip, int row_index, int width, int ld) {
    pdl_launch_dependents();
    pdl_wait();
    const int t = load_step(step_ptr);
    const int w4 = width >> 2;
    const float4* trow = reinterpret_cast<const float4*>(table + (size_t)t * width);
    if (init) {
        const size_t total = (size_t)n_clips * rows_per_clip * w4;
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
            const int j = (int)(i % w4);
            const size_t row = i / w4;
            const bool is_t = (int)(row % rows_per_clip) == row_index;
            const float4 val = is_t ? __ldg(trow + j) : __ldcg(reinterpret_cast<const float4*>(init + row * ld) + j);
            reinterpret_cast<float4*>(dst + row * ld)[j] = val;
        }
    } else {
        const size_t total = (size_t)n_clips * w4;
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
            const int j = (int)(i % w4);
            const size_t row = (i / w4) * rows_per_clip + row_index;
            reinterpret_cast<float4*>(dst + row * ld)[j] = __ldg(trow + j);
        }
    }
}

__global__ void __launch_bounds__(256) scatter_row_bf16_kernel(__nv_bfloat16* __restrict__ dst,
                                                               const __nv_bfloat16* __restrict__ table,
                                                               const int* __restrict__ step_ptr, int n_clips,
                                                               int rows_per_clip, int row_index, int width, int ld) {
    const int trace_slot = threadIdx.x == 0 ? GD_TRACE_OPEN(300) : -1;
    GD_TRACE_MARK(trace_slot, 0);
    pdl_launch_dependents();
    pdl_wait();
    GD_TRACE_MARK(trace_slot, 2);
    const int t = load_step(step_ptr);
    const int w8 = width >> 3;
    const uint4* trow = reinterpret_cast<const uint4*>(table + (size_t)t * width);
    const size_t total = (size_t)n_clips * w8;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(i % w8);
        const size_t row = (i / w8) * rows_per_clip + row_index;
        reinterpret_cast<uint4*>(dst + row * ld)[j] = __ldg(trow + j);
    }
}

// ------------------------------------------------------------------------------- layout helpers
__global__ void __launch_bounds__(256) pack_pose_rows_kernel(const float* __restrict__ x, const float* __restrict__ add,
                                                             __nv_bfloat16* __restrict__ xa, int n_clips, int C,
                                                             int T, int ld) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t total = (size_t)n_clips * T * ld;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % ld);
        const size_t row = i / ld;
        const int frame = (int)(row % T);
        const size_t clip = row / T;
        float v = 0.f;
        if (c < C) {
            const size_t idx = (clip * C + c) * T + frame;
            v = add ? __ldcg(x + idx) + __ldg(add + idx) : __ldcg(x + idx);
        }
        xa[i] = __float2bfloat16_rn(v);
    }
}

__global__ void __launch_bounds__(256) cast_rows_bf16_kernel(const float* __restrict__ src, int lds,
                                                             __nv_bfloat16* __restrict__ dst, int ldd, int rows,
                                                             int cols, int cols_padded) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t total = (size_t)rows * cols_padded;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cols_padded);
        const size_t r = i / cols_padded;
        dst[r * ldd + c] = __float2bfloat16_rn(c < cols ? __ldcg(src + r * lds + c) : 0.f);
    }
}

__global__ void step_add_kernel(int* step_ptr, int delta) {
    const int trace_slot = GD_TRACE_OPEN(400);
    GD_TRACE_MARK(trace_slot, 0);
    pdl_launch_dependents();
    pdl_wait();
    GD_TRACE_MARK(trace_slot, 2);
    *step_ptr = load_step(step_ptr) + delta;
    GD_TRACE_MARK(trace_slot, 8);
}

static inline int grid_for(size_t total, int block) {
    size_t g = (total + block - 1) / block;
    const size_t cap = (size_t)sm_count() * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

#ifdef GD_TRACE
void set_trace_elementwise(unsigned long long* buf) { cudaMemcpyToSymbol(t_trace_buf, &buf, sizeof(buf)); }
#endif

}  // namespace gd

using namespace gd;

extern "C" int gd_layernorm_split(const float* x, int32_t ldx, const float* gamma, const float* beta, const float* gamma2,
                                  const float* beta2, int32_t split_row, void* out_bf16, int32_t ldo, int32_t M, int32_t D,
                                  float eps, void* stream);

extern "C" int gd_layernorm(const float* x, int32_t ldx, const float* gamma, const float* beta, void* out_bf16,
                            int32_t ldo, int32_t M, int32_t D, float eps, void* stream) {
    return gd_layernorm_split(x, ldx, gamma, beta, nullptr, nullptr, 0, out_bf16, ldo, M, D, eps, stream);
}

extern "C" int gd_layernorm_split(const float* x, int32_t ldx, const float* gamma, const float* beta, const float* gamma2,
                                  const float* beta2, int32_t split_row, void* out_bf16, int32_t ldo, int32_t M, int32_t D,
                                  float eps, void* stream) {
    gd::KindScope kind_scope("ln");
    if (!x || !gamma || !beta || !out_bf16) return set_error(GD_ERR_INVALID, "gd_layernorm: null pointer");
    if ((gamma2 == nullptr) != (beta2 == nullptr)) return set_error(GD_ERR_INVALID, "gd_layernorm_split: gamma2 / beta2 go together");
    if (M <= 0) return set_error(GD_ERR_INVALID, "gd_layernorm: M <= 0");
    if (ldx % 4 || ldo % 4 || ldx < D || ldo < D) return set_error(GD_ERR_INVALID, "gd_layernorm: bad row stride");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    constexpr int R = 2;  // rows per warp
    const int wpb = 8, grid = (M + wpb * R - 1) / (wpb * R);
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out_bf16);
    if (D == 256) {
        GD_CUDA_CHECK(launch_k(layernorm_rows_kernel<256, R>, grid, wpb * 32, 0, s, 1, x, ldx, gamma, beta, o, ldo, M, eps, gamma2, beta2, split_row));
    } else if (D == 512) {
        GD_CUDA_CHECK(launch_k(layernorm_rows_kernel<512, R>, grid, wpb * 32, 0, s, 1, x, ldx, gamma, beta, o, ldo, M, eps, gamma2, beta2, split_row));
    } else if (D == 128) {
        GD_CUDA_CHECK(launch_k(layernorm_rows_kernel<128, R>, grid, wpb * 32, 0, s, 1, x, ldx, gamma, beta, o, ldo, M, eps, gamma2, beta2, split_row));
    } else if (D == 1024) {
        GD_CUDA_CHECK(launch_k(layernorm_rows_kernel<1024, R>, grid, wpb * 32, 0, s, 1, x, ldx, gamma, beta, o, ldo, M, eps, gamma2, beta2, split_row));
    } else {
        return set_error(GD_ERR_INVALID, "gd_layernorm: D=%d unsupported (128/256/512/1024)", D);
    }
    count_launch();
    GD_CUDA_CHECK(cudaGetLastError());
    return GD_OK;
}

extern "C" int gd_ddpm_update(const gd_ddpm_desc* u, const float* eps, void* stream) {
    gd::KindScope kind_scope("ddpm");
    int rc = validate_ddpm(u);
    if (rc) return rc;
    if (!eps) return set_error(GD_ERR_INVALID, "gd_ddpm_update: eps is null");
    int c_span = u->C;
    if (u->xa_bf16) {
        if (u->ld_xa < u->C) return set_error(GD_ERR_INVALID, "gd_ddpm_update: ld_xa < C");
        c_span = u->ld_xa;
    }
    const size_t total = (size_t)u->n_clips * c_span * u->T;
    GD_CUDA_CHECK(launch_k(ddpm_update_kernel, grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream), 1, *u, eps, c_span));
    count_launch();
    GD_CUDA_CHECK(cudaGetLastError());
    return GD_OK;
}

extern "C" int gd_scatter_step_row_f32(float* dst, const float* init, const float* table, const int32_t* step_ptr,
                                       int32_t n_clips, int32_t rows_per_clip, int32_t row_index, int32_t width,
                                       int32_t ld, void* stream) {
    gd::KindScope kind_scope("scatter");
    if (!dst || !table || !step_ptr) return set_error(GD_ERR_INVALID, "gd_scatter_step_row_f32: null pointer");
    if (width % 4 || ld % 4 || ld < width || row_index < 0 || row_index >= rows_per_clip || n_clips <= 0)
        return set_error(GD_ERR_INVALID, "gd_scatter_step_row_f32: bad shape");
    const size_t total = init ? (size_t)n_clips * rows_per_clip * (width / 4) : (size_t)n_clips * (width / 4);
    GD_CUDA_CHECK(launch_k(scatter_row_f32_kernel, grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream), 1, dst,
                           init, table, step_ptr, n_clips, rows_per_clip, row_index, width, ld));
    count_launch();
    GD_CUDA_CHECK(cudaGetLastError());
    return GD_OK;
}

extern "C" int gd_scatter_step_row_bf16(void* dst, const void* table, const int32_t* step_ptr, int32_t n_clips,
                                        int32_t rows_per_clip, int32_t row_index, int32_t width, int32_t ld,
                                        void* stream) {
    gd::KindScope kind_scope("scatter");
    if (!dst || !table || !step_ptr) return set_error(GD_ERR_INVALID, "gd_scatter_step_row_bf16: null pointer");
    if (width % 8 || ld % 8 || ld < width || row_index < 0 || row_index >= rows_per_clip || n_clips <= 0)
        return set_error(GD_ERR_INVALID, "gd_scatter_step_row_bf16: bad shape");
    const size_t total = (size_t)n_clips * (width / 8);
    GD_CUDA_CHECK(launch_k(scatter_row_bf16_kernel, grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream), 1,
                           reinterpret_cast<__nv_bfloat16*>(dst), reinterpret_cast<const __nv_bfloat16*>(table), step_ptr,
                           n_clips, rows_per_clip, row_index, width, ld));
    count_launch();
    GD_CUDA_CHECK(cudaGetLastError());
    return GD_OK;
}

extern "C" int gd_pack_pose_rows_add(const float* x, const float* add, void* xa_bf16, int32_t n_clips, int32_t C, int32_t T,
                                     int32_t ld, void* stream) {
    gd::KindScope kind_scope("pack");
    if (!x || !xa_bf16 || n_clips <= 0 || C <= 0 || T <= 0 || ld < C)
        return set_error(GD_ERR_INVALID, "gd_pack_pose_rows: bad argument");
    const size_t total = (size_t)n_clips * T * ld;
    GD_CUDA_CHECK(launch_k(pack_pose_rows_kernel, grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream), 1, x,
                           add, reinterpret_cast<__nv_bfloat16*>(xa_bf16), n_clips, C, T, ld));
    count_launch();
    GD_CUDA_CHECK(cudaGetLastError());
    return GD_OK;
}

extern "C" int gd_pack_pose_rows(const float* x, void* xa_bf16, int32_t n_clips, int32_t C, int32_t T, int32_t ld,
                                 void* stream) {
    return gd_pack_pose_rows_add(x, nullptr, xa_bf16, n_clips, C, T, ld, stream);
}

extern "C" int gd_cast_rows_bf16(const float* src, int32_t lds, void* dst, int32_t ldd, int32_t rows, int32_t cols,
                                 int32_t cols_padded, void* stream) {
    gd::KindScope kind_scope("pack");
    if (!src || !dst || rows <= 0 || cols <= 0 || cols_padded < cols || ldd < cols_padded || lds < cols)
        return set_error(GD_ERR_INVALID, "gd_cast_rows_bf16: bad argument");
    const size_t total = (size_t)rows * cols_padded;
    GD_CUDA_CHECK(launch_k(cast_rows_bf16_kernel, grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream), 1, src,
                           lds, reinterpret_cast<__nv_bfloat16*>(dst), ldd, rows, cols, cols_padded));
    count_launch();
    GD_CUDA_CHECK(cudaGetLastError());
    return GD_OK;
}

extern "C" int gd_step_add(int32_t* step_ptr, int32_t delta, void* stream) {
    gd::KindScope kind_scope("step");
    if (!step_ptr) return set_error(GD_ERR_INVALID, "gd_step_add: null pointer");
    GD_CUDA_CHECK(launch_k(step_add_kernel, 1, 1, 0, reinterpret_cast<cudaStream_t>(stream), 1, step_ptr, delta));
    count_launch();
    GD_CUDA_CHECK(cudaGetLastError());
    return GD_OK;
}
