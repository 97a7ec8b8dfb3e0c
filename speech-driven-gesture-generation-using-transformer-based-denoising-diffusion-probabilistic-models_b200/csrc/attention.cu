// MultiDConvHeadAttention core (reference transformer.py:19-44,88-126) as one fused kernel:
// depth-wise conv3 over tokens of the projected Q/K/V rows (prologue, straight from HBM with the
// +-1 token halo), scores, softmax over keys, P·V — one CTA per (clip, head), the whole head in
// shared memory, single pass (every K/V of a clip fits: <= 160 keys), warp-level softmax.
// v1 computes on the CUDA cores in fp32; inputs are bf16 (default) or fp32 (fp32-activation path).
#include "common.cuh"
#include "host_util.h"

namespace gd {

constexpr int ATT_THREADS = 256;
constexpr int ATT_WARPS = ATT_THREADS / 32;

struct AttnParams {
    const void* q[2];
    const void* k[2];
    const void* v[2];
    __nv_bfloat16* out[2];
    int q_rows[2], q_ld[2], kv_rows[2], kv_ld[2], out_ld[2];
    const float *wq, *bq, *wk, *bk, *wv, *bv;
    int heads, d_k, Lq, Lk, Lk_pad;
    float scale;
};

template <typename T>
__device__ __forceinline__ float ld_elem(const T* p);
template <>
__device__ __forceinline__ float ld_elem<float>(const float* p) {
    return *p;
}
template <>
__device__ __forceinline__ float ld_elem<__nv_bfloat16>(const __nv_bfloat16* p) {
    return __bfloat162float(*p);
}

// raw (pre-conv) element of token `pos` of the concatenated per-clip sequence; 0 outside [0, L).
template <typename T>
__device__ __forceinline__ float raw_token(const void* const (&seg)[2], const int (&rows)[2], const int (&ld)[2],
                                           int clip, int pos, int L, int col) {
    if (pos < 0 || pos >= L) return 0.f;
    if (pos < rows[0]) return ld_elem<T>(reinterpret_cast<const T*>(seg[0]) + ((size_t)clip * rows[0] + pos) * ld[0] + col);
    return ld_elem<T>(reinterpret_cast<const T*>(seg[1]) + ((size_t)clip * rows[1] + (pos - rows[0])) * ld[1] + col);
}

template <typename T>
__device__ __forceinline__ void conv_into_smem(float* dst, int stride, const void* const (&seg)[2],
                                               const int (&rows)[2], const int (&ld)[2], int clip, int head, int L,
                                               int d_k, const float* w, const float* b) {
    for (int i = threadIdx.x; i < L * d_k; i += ATT_THREADS) {
        const int c = i % d_k, pos = i / d_k;
        const int col = head * d_k + c;
        const float a0 = raw_token<T>(seg, rows, ld, clip, pos - 1, L, col);
        const float a1 = raw_token<T>(seg, rows, ld, clip, pos, L, col);
        const float a2 = raw_token<T>(seg, rows, ld, clip, pos + 1, L, col);
        dst[pos * stride + c] = __ldg(w + c * 3 + 0) * a0 + __ldg(w + c * 3 + 1) * a1 + __ldg(w + c * 3 + 2) * a2 + __ldg(b + c);
    }
}

template <typename T, int DK>
__global__ void __launch_bounds__(ATT_THREADS) dconv_attention_kernel(const AttnParams p) {
    extern __shared__ float sm[];
    constexpr int STR = DK + 1;  // +1 float: conflict-free row-strided reads
    float* sq = sm;
    float* sk = sq + p.Lq * STR;
    float* sv = sk + p.Lk * STR;
    float* sp = sv + p.Lk * STR;  // [ATT_WARPS][Lk_pad] probabilities
    const int clip = blockIdx.x / p.heads, head = blockIdx.x % p.heads;
    conv_into_smem<T>(sq, STR, p.q, p.q_rows, p.q_ld, clip, head, p.Lq, DK, p.wq, p.bq);
    conv_into_smem<T>(sk, STR, p.k, p.kv_rows, p.kv_ld, clip, head, p.Lk, DK, p.wk, p.bk);
    conv_into_smem<T>(sv, STR, p.v, p.kv_rows, p.kv_ld, clip, head, p.Lk, DK, p.wv, p.bv);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* prow = sp + warp * p.Lk_pad;
    constexpr int MAXJ = 5;  // <= 160 keys
    for (int i = warp; i < p.Lq; i += ATT_WARPS) {
        const float* qi = sq + i * STR;
        float s[MAXJ];
        float mx = -INFINITY;
#pragma unroll
        for (int jj = 0; jj < MAXJ; ++jj) {
            const int j = jj * 32 + lane;
            s[jj] = -INFINITY;
            if (j < p.Lk) {
                const float* kj = sk + j * STR;
                float acc = 0.f;
#pragma unroll 8
                for (int c = 0; c < DK; ++c) acc = fmaf(qi[c], kj[c], acc);
                s[jj] = acc * p.scale;
            }
            mx = fmaxf(mx, s[jj]);
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int jj = 0; jj < MAXJ; ++jj) {
            const int j = jj * 32 + lane;
            const float e = (j < p.Lk) ? __expf(s[jj] - mx) : 0.f;
            s[jj] = e;
            sum += e;
        }
        const float inv = 1.0f / warp_sum(sum);
#pragma unroll
        for (int jj = 0; jj < MAXJ; ++jj) {
            const int j = jj * 32 + lane;
            if (j < p.Lk) prow[j] = s[jj] * inv;
        }
        __syncwarp();
        float o[DK / 32];
#pragma unroll
        for (int r = 0; r < DK / 32; ++r) o[r] = 0.f;
        for (int j = 0; j < p.Lk; ++j) {
            const float pj = prow[j];
#pragma unroll
            for (int r = 0; r < DK / 32; ++r) o[r] = fmaf(pj, sv[j * STR + r * 32 + lane], o[r]);
        }
        __syncwarp();
        __nv_bfloat16* orow = (i < p.q_rows[0])
                                  ? p.out[0] + ((size_t)clip * p.q_rows[0] + i) * p.out_ld[0]
                                  : p.out[1] + ((size_t)clip * p.q_rows[1] + (i - p.q_rows[0])) * p.out_ld[1];
#pragma unroll
        for (int r = 0; r < DK / 32; ++r) orow[head * DK + r * 32 + lane] = __float2bfloat16_rn(o[r]);
    }
}

template <typename T, int DK>
static int launch_attention(const AttnParams& p, int n_clips, cudaStream_t s) {
    const size_t smem = ((size_t)(p.Lq + 2 * p.Lk) * (DK + 1) + (size_t)ATT_WARPS * p.Lk_pad) * sizeof(float);
    if (smem > 227 * 1024) return set_error(GD_ERR_INVALID, "gd_dconv_attention: sequence too long for shared memory");
    static size_t configured = 0;
    if (smem > configured) {
        GD_CUDA_CHECK(cudaFuncSetAttribute(dconv_attention_kernel<T, DK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
        configured = smem;
    }
    dconv_attention_kernel<T, DK><<<n_clips * p.heads, ATT_THREADS, smem, s>>>(p);
    count_launch();
    GD_CUDA_CHECK(cudaGetLastError());
    return GD_OK;
}

static int run_attention(const gd_attn_desc* d, bool fp32_in, void* stream) {
    if (!d) return set_error(GD_ERR_INVALID, "gd_dconv_attention: null descriptor");
    if (!d->q[0] || !d->k[0] || !d->v[0] || !d->out[0]) return set_error(GD_ERR_INVALID, "gd_dconv_attention: segment 0 missing");
    if (!d->conv_wq || !d->conv_bq || !d->conv_wk || !d->conv_bk || !d->conv_wv || !d->conv_bv)
        return set_error(GD_ERR_INVALID, "gd_dconv_attention: conv taps missing");
    if (d->n_clips <= 0 || d->heads <= 0) return set_error(GD_ERR_INVALID, "gd_dconv_attention: bad clip/head count");
    AttnParams p{};
    for (int s = 0; s < 2; ++s) {
        p.q[s] = d->q[s], p.k[s] = d->k[s], p.v[s] = d->v[s];
        p.out[s] = reinterpret_cast<__nv_bfloat16*>(d->out[s]);
        p.q_rows[s] = d->q[s] ? d->q_rows[s] : 0;
        p.kv_rows[s] = d->k[s] ? d->kv_rows[s] : 0;
        p.q_ld[s] = d->q_ld[s], p.kv_ld[s] = d->kv_ld[s], p.out_ld[s] = d->out_ld[s];
    }
    if ((d->q[1] && !d->out[1]) || (d->k[1] && !d->v[1])) return set_error(GD_ERR_INVALID, "gd_dconv_attention: segment 1 incomplete");
    p.wq = d->conv_wq, p.bq = d->conv_bq, p.wk = d->conv_wk, p.bk = d->conv_bk, p.wv = d->conv_wv, p.bv = d->conv_bv;
    p.heads = d->heads, p.d_k = d->d_k, p.scale = d->scale;
    p.Lq = p.q_rows[0] + p.q_rows[1];
    p.Lk = p.kv_rows[0] + p.kv_rows[1];
    if (p.Lq <= 0 || p.Lk <= 0 || p.Lk > 160) return set_error(GD_ERR_INVALID, "gd_dconv_attention: need 0 < keys <= 160 (got %d)", p.Lk);
    p.Lk_pad = (p.Lk + 31) & ~31;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (d->d_k == 32) return fp32_in ? launch_attention<float, 32>(p, d->n_clips, s) : launch_attention<__nv_bfloat16, 32>(p, d->n_clips, s);
    if (d->d_k == 64) return fp32_in ? launch_attention<float, 64>(p, d->n_clips, s) : launch_attention<__nv_bfloat16, 64>(p, d->n_clips, s);
    return set_error(GD_ERR_INVALID, "gd_dconv_attention: d_k=%d unsupported (32/64)", d->d_k);
}

}  // namespace gd

extern "C" int gd_dconv_attention(const gd_attn_desc* d, void* stream) { return gd::run_attention(d, false, stream); }
extern "C" int gd_dconv_attention_f32in(const gd_attn_desc* d, void* stream) { return gd::run_attention(d, true, stream); }
