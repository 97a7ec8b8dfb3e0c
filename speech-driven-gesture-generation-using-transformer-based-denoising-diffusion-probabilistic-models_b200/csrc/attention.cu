// MultiDConvHeadAttention core (reference transformer.py:19-44,88-126) as one fused kernel.
//
// One CTA per (clip, head).  Prologue: the depth-wise conv3 over tokens is applied while the projected Q/K/V rows
// are staged HBM -> shared memory (16-byte loads, each raw row read ~1.25x thanks to 8-token segments with a
// +-1 halo held in registers), results rounded to bf16.  Main loop: every warp owns 16-query tiles; S = Q·Kᵀ on
// the tensor cores (mma.sync m16n8k16 bf16, fragments via ldmatrix), softmax over keys entirely in registers
// (quad shuffles, exp2 with the scale folded in), then O = P·V with P re-used straight from the accumulator
// registers as the A fragment (no shared-memory round trip) and V read through ldmatrix.trans.
// All K/V of a head fit in shared memory (<= 160 keys), so this is single-pass: no running-max rescale.
// The token sequence of a clip is the concatenation of up to two row segments; the conv runs across the seam
// (tedexp joint attention over [x ; memory], nn.py:105-113) and zero-pads only at the sequence ends.
#include "common.cuh"
#include "host_util.h"

namespace gd {

constexpr int ATT_MAX_WARPS = 5;  // 4 or 5 warps per CTA (chosen per launch), several CTAs per SM

struct AttnParams {
    const void* q[2];
    const void* k[2];
    const void* v[2];
    __nv_bfloat16* out[2];
    int q_rows[2], q_ld[2], kv_rows[2], kv_ld[2], out_ld[2];
    const float *wq, *bq, *wk, *bk, *wv, *bv;
    int heads, Lq, Lk;
    float scale_log2;  // d_k^-1/2 * log2(e)
};

// Raw (pre-conv) storage of 8 consecutive elements of a projected row: one 16-B load for bf16, two for fp32.
// Loads go through the non-coherent path (ld.global.nc) so the compiler may batch them ahead of the smem stores.
template <typename T>
struct Raw8;
template <>
struct Raw8<__nv_bfloat16> {
    uint4 u;
    __device__ __forceinline__ void load(const __nv_bfloat16* p) { u = __ldg(reinterpret_cast<const uint4*>(p)); }
    __device__ __forceinline__ void zero() { u = make_uint4(0, 0, 0, 0); }
    __device__ __forceinline__ void unpack(float (&f)[8]) const {
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f[2 * i] = __uint_as_float(w[i] << 16);
            f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
};
template <>
struct Raw8<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float* p) {
        a = __ldg(reinterpret_cast<const float4*>(p));
        b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    }
    __device__ __forceinline__ void zero() { a = b = make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ void unpack(float (&f)[8]) const {
        f[0] = a.x, f[1] = a.y, f[2] = a.z, f[3] = a.w, f[4] = b.x, f[5] = b.y, f[6] = b.z, f[7] = b.w;
    }
};

// Branch-free fetch of 8 elements of token `pos` of the concatenated per-clip sequence (zeros outside [0, L)):
// the address is formed from a clamped position with 32-bit offsets and selects, the result is masked.
template <typename T>
__device__ __forceinline__ void raw_row8(Raw8<T>& r, const T* base0, const T* base1, int rows0, int ld0, int ld1, int pos,
                                         int L) {
    const bool valid = static_cast<unsigned>(pos) < static_cast<unsigned>(L);
    const int pc = min(max(pos, 0), L - 1);
    const bool in0 = pc < rows0;
    const T* p = in0 ? base0 + pc * ld0 : base1 + (pc - rows0) * ld1;
    r.load(p);
    if (!valid) r.zero();
}

// dst[pos][c] = bf16( w[c][0]*raw[pos-1][c] + w[c][1]*raw[pos][c] + w[c][2]*raw[pos+1][c] + b[c] ), rows >= L zeroed.
// One work item = SEGT consecutive tokens x 8 columns: all SEGT+2 raw rows are fetched up front (SEGT+2 independent
// 16-B loads in flight per thread), then the 3-tap FIR slides over them in registers (3 FFMA per element).
template <typename T, int DK>
__device__ __forceinline__ void conv_stage(__nv_bfloat16* dst, int L, int L_pad, const void* const (&seg)[2],
                                           const int (&rows)[2], const int (&ld)[2], int clip, int head,
                                           const float* s_taps /* [DK*3] */, const float* s_bias /* [DK] */) {
    constexpr int STR = DK + 8;
    constexpr int CH = DK / 8;
    constexpr int SEGT = sizeof(T) == 2 ? 8 : 4;
    const int n_seg = (L_pad + SEGT - 1) / SEGT;
    const int rows0 = rows[0], ld0 = ld[0], ld1 = ld[1];
    const T* clip0 = reinterpret_cast<const T*>(seg[0]) + (size_t)clip * rows0 * ld0 + head * DK;
    const T* clip1 = seg[1] ? reinterpret_cast<const T*>(seg[1]) + (size_t)clip * rows[1] * ld1 + head * DK : clip0;
    for (int item = threadIdx.x; item < n_seg * CH; item += blockDim.x) {
        const int ch = item % CH, sg = item / CH;
        const int c0 = ch * 8, p0 = sg * SEGT;
        Raw8<T> raw[SEGT + 2];
#pragma unroll
        for (int s = 0; s < SEGT + 2; ++s) raw_row8<T>(raw[s], clip0 + c0, clip1 + c0, rows0, ld0, ld1, p0 - 1 + s, L);
        float w0[8], w1[8], w2[8], bb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            w0[i] = s_taps[(c0 + i) * 3 + 0];
            w1[i] = s_taps[(c0 + i) * 3 + 1];
            w2[i] = s_taps[(c0 + i) * 3 + 2];
            bb[i] = s_bias[c0 + i];
        }
        float prev[8], cur[8], nxt[8];
        raw[0].unpack(prev);
        raw[1].unpack(cur);
#pragma unroll
        for (int s = 0; s < SEGT; ++s) {
            const int pos = p0 + s;
            raw[s + 2].unpack(nxt);
            float r[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = fmaf(w0[i], prev[i], fmaf(w1[i], cur[i], fmaf(w2[i], nxt[i], bb[i])));
            uint4 o;
            o.x = pack_bf16x2(r[0], r[1]);
            o.y = pack_bf16x2(r[2], r[3]);
            o.z = pack_bf16x2(r[4], r[5]);
            o.w = pack_bf16x2(r[6], r[7]);
            if (pos >= L) o = make_uint4(0, 0, 0, 0);
            if (pos < L_pad) *reinterpret_cast<uint4*>(dst + pos * STR + c0) = o;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                prev[i] = cur[i];
                cur[i] = nxt[i];
            }
        }
    }
}

__device__ __forceinline__ float ex2_approx(float x) {  // MUFU.EX2; -inf -> 0, which is what masked keys need
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// KB = number of 16-key blocks (keys padded to 16*KB), DK = head width
template <typename T, int DK, int KB>
__global__ void __launch_bounds__(ATT_MAX_WARPS * 32, 3) dconv_attention_kernel(const AttnParams p) {
    constexpr int STR = DK + 8;  // bf16 elements per smem row: +16 B keeps ldmatrix rows on distinct banks
    constexpr int LK_PAD = KB * 16;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int Lq_pad = (p.Lq + 15) & ~15;
    __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __nv_bfloat16* sk = sq + Lq_pad * STR;
    __nv_bfloat16* sv = sk + LK_PAD * STR;
    float* s_taps = reinterpret_cast<float*>(sv + LK_PAD * STR);  // [3][DK*3 + DK]
    const int clip = blockIdx.x / p.heads, head = blockIdx.x % p.heads;

    for (int i = threadIdx.x; i < 3 * DK * 4; i += blockDim.x) {
        const int which = i / (DK * 4), j = i % (DK * 4);
        const float* w = which == 0 ? p.wq : (which == 1 ? p.wk : p.wv);
        const float* b = which == 0 ? p.bq : (which == 1 ? p.bk : p.bv);
        s_taps[i] = j < DK * 3 ? __ldg(w + j) : __ldg(b + j - DK * 3);
    }
    __syncthreads();
    conv_stage<T, DK>(sq, p.Lq, Lq_pad, p.q, p.q_rows, p.q_ld, clip, head, s_taps, s_taps + DK * 3);
    conv_stage<T, DK>(sk, p.Lk, LK_PAD, p.k, p.kv_rows, p.kv_ld, clip, head, s_taps + DK * 4, s_taps + DK * 7);
    conv_stage<T, DK>(sv, p.Lk, LK_PAD, p.v, p.kv_rows, p.kv_ld, clip, head, s_taps + DK * 8, s_taps + DK * 11);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const uint32_t sq_u = smem_u32(sq), sk_u = smem_u32(sk), sv_u = smem_u32(sv);

    const int n_warps = blockDim.x >> 5;
    for (int qt = warp; qt * 16 < p.Lq; qt += n_warps) {
        // ---- S = Q Kᵀ : 16 x LK_PAD, fp32 accumulators in registers
        float s[KB * 2][4];
#pragma unroll
        for (int n = 0; n < KB * 2; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < DK / 16; ++kk) {
            uint32_t a0, a1, a2, a3;
            ldsm_x4(sq_u + ((qt * 16 + (lane & 15)) * STR + kk * 16 + (lane >> 4) * 8) * 2, a0, a1, a2, a3);
#pragma unroll
            for (int nb = 0; nb < KB; ++nb) {
                uint32_t b0, b1, b2, b3;
                ldsm_x4(sk_u + ((nb * 16 + (lane & 7) + (lane >> 4) * 8) * STR + kk * 16 + ((lane >> 3) & 1) * 8) * 2, b0, b1,
                        b2, b3);
                mma_bf16_16816(s[2 * nb], a0, a1, a2, a3, b0, b1);
                mma_bf16_16816(s[2 * nb + 1], a0, a1, a2, a3, b2, b3);
            }
        }
        // ---- softmax over keys (rows g and g+8 of the tile); only the last key tiles can hold padded keys
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int n = 0; n < KB * 2; ++n) {
            if (n >= KB * 2 - 2 && n * 8 + 8 > p.Lk) {  // warp-uniform; at most the two tiles of the last 16-key block
                const int j = n * 8 + 2 * t;
                if (j >= p.Lk) s[n][0] = s[n][2] = -INFINITY;
                if (j + 1 >= p.Lk) s[n][1] = s[n][3] = -INFINITY;
            }
            m0 = fmaxf(m0, fmaxf(s[n][0], s[n][1]));
            m1 = fmaxf(m1, fmaxf(s[n][2], s[n][3]));
        }
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        const float o0 = m0 * p.scale_log2, o1 = m1 * p.scale_log2;
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int n = 0; n < KB * 2; ++n) {
            s[n][0] = ex2_approx(fmaf(s[n][0], p.scale_log2, -o0));
            s[n][1] = ex2_approx(fmaf(s[n][1], p.scale_log2, -o0));
            s[n][2] = ex2_approx(fmaf(s[n][2], p.scale_log2, -o1));
            s[n][3] = ex2_approx(fmaf(s[n][3], p.scale_log2, -o1));
            sum0 += s[n][0] + s[n][1];
            sum1 += s[n][2] + s[n][3];
        }
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
        sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
        sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
        // ---- O = P V : P comes straight from the accumulators (C layout of two n-tiles == A layout of one k-block)
        float o[DK / 8][4];
#pragma unroll
        for (int n = 0; n < DK / 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
            const uint32_t a0 = pack_bf16x2(s[2 * kb][0], s[2 * kb][1]), a1 = pack_bf16x2(s[2 * kb][2], s[2 * kb][3]);
            const uint32_t a2 = pack_bf16x2(s[2 * kb + 1][0], s[2 * kb + 1][1]), a3 = pack_bf16x2(s[2 * kb + 1][2], s[2 * kb + 1][3]);
#pragma unroll
            for (int nb = 0; nb < DK / 16; ++nb) {
                uint32_t b0, b1, b2, b3;
                ldsm_x4_trans(sv_u + ((kb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * STR + nb * 16 + (lane >> 4) * 8) * 2, b0,
                              b1, b2, b3);
                mma_bf16_16816(o[2 * nb], a0, a1, a2, a3, b0, b1);
                mma_bf16_16816(o[2 * nb + 1], a0, a1, a2, a3, b2, b3);
            }
        }
        const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int i = qt * 16 + g + half * 8;
            if (i < p.Lq) {
                __nv_bfloat16* orow = (i < p.q_rows[0])
                                          ? p.out[0] + ((size_t)clip * p.q_rows[0] + i) * p.out_ld[0]
                                          : p.out[1] + ((size_t)clip * p.q_rows[1] + (i - p.q_rows[0])) * p.out_ld[1];
                const float inv = half ? inv1 : inv0;
#pragma unroll
                for (int n = 0; n < DK / 8; ++n)
                    *reinterpret_cast<uint32_t*>(orow + head * DK + n * 8 + 2 * t) =
                        pack_bf16x2(o[n][2 * half] * inv, o[n][2 * half + 1] * inv);
            }
        }
    }
}

template <typename T, int DK, int KB>
static int launch_attention(const AttnParams& p, int n_clips, cudaStream_t s) {
    const int Lq_pad = (p.Lq + 15) & ~15;
    const size_t smem = (size_t)(Lq_pad + 2 * KB * 16) * (DK + 8) * 2 + 3 * DK * 4 * sizeof(float);
    static size_t configured = 0;
    if (smem > configured) {
        const size_t want = smem > 48 * 1024 ? smem : 48 * 1024;
        GD_CUDA_CHECK(cudaFuncSetAttribute(dconv_attention_kernel<T, DK, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)want));
        // without this the driver sizes the L1/smem split for ONE block and occupancy collapses to 1 CTA/SM
        GD_CUDA_CHECK(cudaFuncSetAttribute(dconv_attention_kernel<T, DK, KB>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                           cudaSharedmemCarveoutMaxShared));
        configured = want;
    }
    // 4-5 warps per CTA and several CTAs per SM hide latency better than one fat CTA (measured); pick the count that
    // leaves the fewest idle warp-rounds over the 16-query tiles
    const int tiles = Lq_pad / 16;
    int warps = 4;
    if (tiles > 4) {
        const int r4 = (tiles + 3) / 4, r5 = (tiles + 4) / 5;
        warps = (r5 * 5 - tiles < r4 * 4 - tiles) ? 5 : 4;
    }
    dconv_attention_kernel<T, DK, KB><<<n_clips * p.heads, warps * 32, smem, s>>>(p);
    count_launch();
    GD_CUDA_CHECK(cudaGetLastError());
    return GD_OK;
}

template <typename T, int DK>
static int dispatch_kb(const AttnParams& p, int n_clips, cudaStream_t s) {
    switch ((p.Lk + 15) / 16) {
        case 1: return launch_attention<T, DK, 1>(p, n_clips, s);
        case 2: return launch_attention<T, DK, 2>(p, n_clips, s);
        case 3: return launch_attention<T, DK, 3>(p, n_clips, s);
        case 4: return launch_attention<T, DK, 4>(p, n_clips, s);
        case 5: return launch_attention<T, DK, 5>(p, n_clips, s);
        case 6: return launch_attention<T, DK, 6>(p, n_clips, s);
        case 7: return launch_attention<T, DK, 7>(p, n_clips, s);
        case 8: return launch_attention<T, DK, 8>(p, n_clips, s);
        case 9: return launch_attention<T, DK, 9>(p, n_clips, s);
        case 10: return launch_attention<T, DK, 10>(p, n_clips, s);
    }
    return set_error(GD_ERR_INVALID, "gd_dconv_attention: need 0 < keys <= 160 (got %d)", p.Lk);
}

static int run_attention(const gd_attn_desc* d, bool fp32_in, void* stream) {
    if (!d) return set_error(GD_ERR_INVALID, "gd_dconv_attention: null descriptor");
    if (!d->q[0] || !d->k[0] || !d->v[0] || !d->out[0]) return set_error(GD_ERR_INVALID, "gd_dconv_attention: segment 0 missing");
    if (!d->conv_wq || !d->conv_bq || !d->conv_wk || !d->conv_bk || !d->conv_wv || !d->conv_bv)
        return set_error(GD_ERR_INVALID, "gd_dconv_attention: conv taps missing");
    if (d->n_clips <= 0 || d->heads <= 0) return set_error(GD_ERR_INVALID, "gd_dconv_attention: bad clip/head count");
    if ((d->q[1] && !d->out[1]) || (d->k[1] && !d->v[1])) return set_error(GD_ERR_INVALID, "gd_dconv_attention: segment 1 incomplete");
    AttnParams p{};
    const int align = fp32_in ? 4 : 8;  // 16-byte row loads
    for (int s = 0; s < 2; ++s) {
        p.q[s] = d->q[s], p.k[s] = d->k[s], p.v[s] = d->v[s];
        p.out[s] = reinterpret_cast<__nv_bfloat16*>(d->out[s]);
        p.q_rows[s] = d->q[s] ? d->q_rows[s] : 0;
        p.kv_rows[s] = d->k[s] ? d->kv_rows[s] : 0;
        p.q_ld[s] = d->q_ld[s], p.kv_ld[s] = d->kv_ld[s], p.out_ld[s] = d->out_ld[s];
        if ((d->q[s] && (d->q_ld[s] % align || (reinterpret_cast<uintptr_t>(d->q[s]) & 15) || d->out_ld[s] % 2)) ||
            (d->k[s] && (d->kv_ld[s] % align || ((reinterpret_cast<uintptr_t>(d->k[s]) | reinterpret_cast<uintptr_t>(d->v[s])) & 15))))
            return set_error(GD_ERR_INVALID, "gd_dconv_attention: q/k/v rows must be 16-byte aligned");
    }
    p.wq = d->conv_wq, p.bq = d->conv_bq, p.wk = d->conv_wk, p.bk = d->conv_bk, p.wv = d->conv_wv, p.bv = d->conv_bv;
    p.heads = d->heads;
    p.scale_log2 = d->scale * 1.4426950408889634f;
    p.Lq = p.q_rows[0] + p.q_rows[1];
    p.Lk = p.kv_rows[0] + p.kv_rows[1];
    if (p.Lq <= 0 || p.Lk <= 0 || p.Lk > 160 || p.Lq > 1024)
        return set_error(GD_ERR_INVALID, "gd_dconv_attention: need 0 < keys <= 160 (got %d), queries <= 1024", p.Lk);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (d->d_k == 32) return fp32_in ? dispatch_kb<float, 32>(p, d->n_clips, s) : dispatch_kb<__nv_bfloat16, 32>(p, d->n_clips, s);
    if (d->d_k == 64) return fp32_in ? dispatch_kb<float, 64>(p, d->n_clips, s) : dispatch_kb<__nv_bfloat16, 64>(p, d->n_clips, s);
    return set_error(GD_ERR_INVALID, "gd_dconv_attention: d_k=%d unsupported (32/64)", d->d_k);
}

}  // namespace gd

extern "C" int gd_dconv_attention(const gd_attn_desc* d, void* stream) { return gd::run_attention(d, false, stream); }
extern "C" int gd_dconv_attention_f32in(const gd_attn_desc* d, void* stream) { return gd::run_attention(d, true, stream); }
