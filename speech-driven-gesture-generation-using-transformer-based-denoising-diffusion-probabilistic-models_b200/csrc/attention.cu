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
#include "attention_common.cuh"
#include <cstdlib>

namespace gd {

constexpr int ATT_MAX_WARPS = 5;  // 4 or 5 warps per CTA (chosen per launch), several CTAs per SM


// Raw (pre-conv) storage of 8 consecutive elements of a projected row: one 16-B load for bf16, two for fp32.
// Loads are L2-coherent (ld.global.cg): Q/K/V rows are written by the previous kernel of the chain (common.cuh, PDL and L1).
template <typename T>
struct Raw8;
template <>
struct Raw8<__nv_bfloat16> {
    uint4 u;
    __device__ __forceinline__ void load(const __nv_bfloat16* p) { u = __ldcg(reinterpret_cast<const uint4*>(p)); }
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
        a = __ldcg(reinterpret_cast<const float4*>(p));
        b = __ldcg(reinterpret_cast<const float4*>(p) + 1);
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
                                           const int (&rows)[2], const int (&stride)[2], const int (&ld)[2], int clip, int head,
                                           const float* s_taps /* [DK*3] */, const float* s_bias /* [DK] */) {
    constexpr int STR = DK + 8;
    constexpr int CH = DK / 8;
    constexpr int SEGT = sizeof(T) == 2 ? 8 : 4;
    const int n_seg = (L_pad + SEGT - 1) / SEGT;
    const int rows0 = rows[0], ld0 = ld[0], ld1 = ld[1];
    const T* clip0 = reinterpret_cast<const T*>(seg[0]) + (size_t)clip * stride[0] * ld0 + head * DK;
    const T* clip1 = seg[1] ? reinterpret_cast<const T*>(seg[1]) + (size_t)clip * stride[1] * ld1 + head * DK : clip0;
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

__device__ __forceinline__ void unpack_bf16x8(const uint4& u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
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
    pdl_launch_dependents();
    pdl_wait();  // the conv taps are weights; everything below reads the previous kernel's output
    __syncthreads();
    conv_stage<T, DK>(sq, p.Lq, Lq_pad, p.q, p.q_rows, p.q_stride, p.q_ld, clip, head, s_taps, s_taps + DK * 3);
    conv_stage<T, DK>(sk, p.Lk, LK_PAD, p.k, p.kv_rows, p.kv_rows, p.kv_ld, clip, head, s_taps + DK * 4, s_taps + DK * 7);
    conv_stage<T, DK>(sv, p.Lk, LK_PAD, p.v, p.kv_rows, p.kv_rows, p.kv_ld, clip, head, s_taps + DK * 8, s_taps + DK * 11);
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
            if (i < p.Lq && (i < p.q_rows[0] || p.out[1])) {  // a segment without an output is a conv halo only
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
    static size_t configured_dev[GD_MAX_DEVICES] = {};  // function attributes are per device
    size_t& configured = configured_dev[current_device()];
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
    GD_CUDA_CHECK(launch_k(dconv_attention_kernel<T, DK, KB>, n_clips * p.heads, warps * 32, smem, s, 1, p));
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


// =====================================================================================================================
// v2 (bf16 inputs): persistent, TMA-fed.  A work item is (clip, 64-column head group) = one head for d_k = 64, two for
// d_k = 32.  Thread 0 asks the TMA engine for the item's raw Q/K/V row blocks (dense 128-B rows, one 2-D box per row
// segment) and the whole CTA meanwhile works on the previous item, so HBM latency never sits in front of the math:
//     wait(raw full) -> conv3 raw -> cv (bf16, XOR-swizzled 128-B rows) -> sync -> TMA for the next item into raw
//                    -> QK^T / softmax / PV from cv (mma.sync + ldmatrix) -> sync
// Two CTAs per SM interleave their conv and MMA phases.  Shared memory per CTA: raw (Lq+2Lk)*128 B + cv
// (Lq_pad+2*Lk_pad)*128 B + taps, 111 KB for the 138-token joint attention.
// =====================================================================================================================
constexpr int ATT2_ROW_BYTES = 128;  // 64 bf16 columns per item row

__device__ __forceinline__ uint32_t cv_off(int row, int chunk) {  // byte offset of 16-B chunk `chunk` of row `row`
    return static_cast<uint32_t>(row * ATT2_ROW_BYTES + ((chunk ^ (row & 7)) << 4));
}

struct Attn2Geom {
    int n_items, groups;  // groups = heads / (64 / d_k)
    // byte offsets in dynamic smem.  Every raw tensor block is [zero row | L token rows | zero rows up to L_pad+1]:
    // the zero rows are the conv's "same" padding and are never written by the TMA engine
    // raw_stages (1 or 2) such blocks: small items keep two loads in flight per CTA
    int raw_k_off, raw_v_off, raw_stage_bytes, raw_stages, raw_bytes, cv_q_off, cv_k_off, cv_v_off, taps_off, bar_off;
    uint32_t tx_bytes;
};

// conv3 of 16 tokens x 4 columns: 18 raw rows (8-byte loads) -> 16 swizzled bf16 rows
__device__ __forceinline__ void conv_unit16(const uint8_t* src /* raw row p0 (= token p0-1), column group */,
                                            uint8_t* dst /* cv row p0, 8-byte half selected */, int chunk,
                                            const float* tp /* w0[8] w1[8] w2[8] b[8] of the chunk, +4 for odd halves */) {
    const float4 w0 = *reinterpret_cast<const float4*>(tp), w1 = *reinterpret_cast<const float4*>(tp + 8);
    const float4 w2 = *reinterpret_cast<const float4*>(tp + 16), bb = *reinterpret_cast<const float4*>(tp + 24);
    uint2 raw[18];
#pragma unroll
    for (int s = 0; s < 18; ++s) raw[s] = *reinterpret_cast<const uint2*>(src + s * ATT2_ROW_BYTES);
    float4 prev = unpack_bf16x4(raw[0]), cur = unpack_bf16x4(raw[1]);
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        const float4 nxt = unpack_bf16x4(raw[s + 2]);
        uint2 o;
        o.x = pack_bf16x2(fmaf(w0.x, prev.x, fmaf(w1.x, cur.x, fmaf(w2.x, nxt.x, bb.x))),
                          fmaf(w0.y, prev.y, fmaf(w1.y, cur.y, fmaf(w2.y, nxt.y, bb.y))));
        o.y = pack_bf16x2(fmaf(w0.z, prev.z, fmaf(w1.z, cur.z, fmaf(w2.z, nxt.z, bb.z))),
                          fmaf(w0.w, prev.w, fmaf(w1.w, cur.w, fmaf(w2.w, nxt.w, bb.w))));
        // p0 is a multiple of 16, so the row's swizzle term is the compile-time (s & 7)
        *reinterpret_cast<uint2*>(dst + s * ATT2_ROW_BYTES + ((chunk ^ (s & 7)) << 4)) = o;
        prev = cur;
        cur = nxt;
    }
}

// Same unit with packed bf16x2 arithmetic (three HFMA2 per column pair, no unpack / repack): a third of the
// instructions, at the price of rounding the taps and the two partial sums to bf16.
__device__ __forceinline__ void conv_unit16_bf16(const uint8_t* src, uint8_t* dst, int chunk, const float* tp) {
    const float4 w0 = *reinterpret_cast<const float4*>(tp), w1 = *reinterpret_cast<const float4*>(tp + 8);
    const float4 w2 = *reinterpret_cast<const float4*>(tp + 16), bb = *reinterpret_cast<const float4*>(tp + 24);
    const __nv_bfloat162 w0a = __floats2bfloat162_rn(w0.x, w0.y), w0b = __floats2bfloat162_rn(w0.z, w0.w);
    const __nv_bfloat162 w1a = __floats2bfloat162_rn(w1.x, w1.y), w1b = __floats2bfloat162_rn(w1.z, w1.w);
    const __nv_bfloat162 w2a = __floats2bfloat162_rn(w2.x, w2.y), w2b = __floats2bfloat162_rn(w2.z, w2.w);
    const __nv_bfloat162 bba = __floats2bfloat162_rn(bb.x, bb.y), bbb = __floats2bfloat162_rn(bb.z, bb.w);
    uint2 raw[18];
#pragma unroll
    for (int s = 0; s < 18; ++s) raw[s] = *reinterpret_cast<const uint2*>(src + s * ATT2_ROW_BYTES);
    auto b2 = [](uint32_t u) { return *reinterpret_cast<const __nv_bfloat162*>(&u); };
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        const __nv_bfloat162 ra = __hfma2(w0a, b2(raw[s].x), __hfma2(w1a, b2(raw[s + 1].x), __hfma2(w2a, b2(raw[s + 2].x), bba)));
        const __nv_bfloat162 rb = __hfma2(w0b, b2(raw[s].y), __hfma2(w1b, b2(raw[s + 1].y), __hfma2(w2b, b2(raw[s + 2].y), bbb)));
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&ra);
        o.y = *reinterpret_cast<const uint32_t*>(&rb);
        *reinterpret_cast<uint2*>(dst + s * ATT2_ROW_BYTES + ((chunk ^ (s & 7)) << 4)) = o;
    }
}

template <int DK, int KB, int MAXT, int MAXREG, bool CONV_BF16, bool TAIL_SPLIT>
__global__ void __launch_bounds__(MAXT) __maxnreg__(MAXREG)
dconv_attention_tma_kernel(const __grid_constant__ CUtensorMap tm_q0, const __grid_constant__ CUtensorMap tm_q1,
                           const __grid_constant__ CUtensorMap tm_k0, const __grid_constant__ CUtensorMap tm_k1,
                           const __grid_constant__ CUtensorMap tm_v0, const __grid_constant__ CUtensorMap tm_v1,
                           const AttnParams p, const Attn2Geom geo) {
    constexpr int HG = 64 / DK;       // heads per item
    constexpr int CPH = DK / 8;       // 16-B chunks per head row
    constexpr int LK_PAD = KB * 16;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    const int Lq = p.Lq, Lk = p.Lk;
    const int Lq_pad = (Lq + 15) & ~15;
    uint8_t* cv_q = smem + geo.cv_q_off;
    uint8_t* cv_k = smem + geo.cv_k_off;
    uint8_t* cv_v = smem + geo.cv_v_off;
    float* s_taps = reinterpret_cast<float*>(smem + geo.taps_off);  // [3][CPH][w0 8 | w1 8 | w2 8 | b 8]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + geo.bar_off);
    float* tail_stats = reinterpret_cast<float*>(smem + geo.bar_off + 16);  // [4 partial tasks][16 rows][max, sum]
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        prefetch_tensormap(&tm_q0), prefetch_tensormap(&tm_k0), prefetch_tensormap(&tm_v0);
        if (p.q_rows[1]) prefetch_tensormap(&tm_q1);
        if (p.kv_rows[1]) prefetch_tensormap(&tm_k1), prefetch_tensormap(&tm_v1);
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        fence_barrier_init();
    }
    for (int i = tid; i < 3 * DK * 4; i += nthr) {
        const int which = i / (DK * 4), r = i % (DK * 4);
        const int j = r / 32, part = (r % 32) / 8, e = r % 8;  // chunk j, part: tap 0..2 or bias
        const float* w = which == 0 ? p.wq : (which == 1 ? p.wk : p.wv);
        const float* b = which == 0 ? p.bq : (which == 1 ? p.bk : p.bv);
        const int c = j * 8 + e;
        s_taps[i] = part < 3 ? __ldg(w + c * 3 + part) : __ldg(b + c);
    }
    for (int i = tid; i < geo.raw_bytes / 16; i += nthr) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();  // the zero fill is ordered before the TMA writes that follow the barrier
    const int trace_slot = tid == 0 ? GD_TRACE_OPEN(500) : -1;
    GD_TRACE_MARK(trace_slot, 1);  // prologue (zero fill, taps) done
    pdl_launch_dependents();
    pdl_wait();  // set-up above touched only weights and shared memory; Q/K/V come from the previous kernel
    GD_TRACE_MARK(trace_slot, 2);
    __syncthreads();

    auto issue_load = [&](int item, int stage) {  // one thread; token rows start at raw row 1 of the stage's blocks
        const int clip = item / geo.groups, col0 = (item % geo.groups) * 64;
        uint8_t* raw_q = smem + stage * geo.raw_stage_bytes;
        uint8_t* raw_k = raw_q + geo.raw_k_off;
        uint8_t* raw_v = raw_q + geo.raw_v_off;
        uint64_t* bar = &full_bar[stage];
        mbar_arrive_expect_tx(bar, geo.tx_bytes);
        tma_load_2d(raw_q + ATT2_ROW_BYTES, &tm_q0, bar, col0, clip * p.q_stride[0]);
        if (p.q_rows[1]) tma_load_2d(raw_q + (1 + p.q_rows[0]) * ATT2_ROW_BYTES, &tm_q1, bar, col0, clip * p.q_stride[1]);
        tma_load_2d(raw_k + ATT2_ROW_BYTES, &tm_k0, bar, col0, clip * p.kv_rows[0]);
        tma_load_2d(raw_v + ATT2_ROW_BYTES, &tm_v0, bar, col0, clip * p.kv_rows[0]);
        if (p.kv_rows[1]) {
            tma_load_2d(raw_k + (1 + p.kv_rows[0]) * ATT2_ROW_BYTES, &tm_k1, bar, col0, clip * p.kv_rows[1]);
            tma_load_2d(raw_v + (1 + p.kv_rows[0]) * ATT2_ROW_BYTES, &tm_v1, bar, col0, clip * p.kv_rows[1]);
        }
    };

    int item = blockIdx.x;
    const int n_stages = geo.raw_stages;
    if (tid == 0)
        for (int st = 0; st < n_stages; ++st)
            if (item + st * (int)gridDim.x < geo.n_items) issue_load(item + st * gridDim.x, st);
    int k_iter = 0;
    const int g = lane >> 2, t = lane & 3;
    const int sw = lane & 7;  // every ldmatrix row index below is (multiple of 8) + (lane & 7)
    const int nsq = Lq_pad / 16, nsk = LK_PAD / 16;
    const int conv_items = (nsq + 2 * nsk) * 16;
    const int q_tiles = Lq_pad / 16;
    // per-lane ldmatrix bases (row part only; the swizzled column part is added per k-step)
    const uint32_t qrow_u = smem_u32(cv_q) + (lane & 15) * ATT2_ROW_BYTES;
    const uint32_t krow_u = smem_u32(cv_k) + ((lane & 7) + (lane >> 4) * 8) * ATT2_ROW_BYTES;
    const uint32_t vrow_u = smem_u32(cv_v) + ((lane & 7) + ((lane >> 3) & 1) * 8) * ATT2_ROW_BYTES;
    const float scale_log2 = p.scale_log2;

    for (; item < geo.n_items; item += gridDim.x, ++k_iter) {
        const int clip = item / geo.groups, grp = item % geo.groups;
        const int stage = n_stages == 2 ? (k_iter & 1) : 0;
        const uint32_t phase = (n_stages == 2 ? (k_iter >> 1) : k_iter) & 1;
        const uint8_t* raw_q = smem + stage * geo.raw_stage_bytes;
        const uint8_t* raw_k = raw_q + geo.raw_k_off;
        const uint8_t* raw_v = raw_q + geo.raw_v_off;
        mbar_wait(&full_bar[stage], phase);
        if (k_iter == 0) GD_TRACE_MARK(trace_slot, 3);  // first item's rows landed
        // ---- depth-wise conv3 over tokens: raw -> cv, one unit = 16 tokens x 4 columns, no boundary predicates
        for (int it = tid; it < conv_items; it += nthr) {
            const int hc4 = it & 15, sg = it >> 4;  // 8-byte column group within the 128-B row, 16-row segment
            const int which = sg < nsq ? 0 : (sg < nsq + nsk ? 1 : 2);
            const int p0 = (sg - (which == 0 ? 0 : (which == 1 ? nsq : nsq + nsk))) * 16;
            const uint8_t* src = (which == 0 ? raw_q : (which == 1 ? raw_k : raw_v)) + p0 * ATT2_ROW_BYTES + hc4 * 8;
            uint8_t* dst = (which == 0 ? cv_q : (which == 1 ? cv_k : cv_v)) + p0 * ATT2_ROW_BYTES + (hc4 & 1) * 8;
            const float* tp = s_taps + which * DK * 4 + ((hc4 >> 1) % CPH) * 32 + (hc4 & 1) * 4;
            if (CONV_BF16)
                conv_unit16_bf16(src, dst, hc4 >> 1, tp);
            else
                conv_unit16(src, dst, hc4 >> 1, tp);
        }
        __syncthreads();  // cv complete, raw free
        if (tid == 0 && item + n_stages * (int)gridDim.x < geo.n_items) issue_load(item + n_stages * gridDim.x, stage);

        // ---- attention core: one task = (head of the group, 16-query tile)
        auto run_tile = [&](const int hh, const int qt) {
            const int hc = hh * CPH;
            float s[KB * 2][4];
#pragma unroll
            for (int n = 0; n < KB * 2; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
            const uint32_t qaddr = qrow_u + qt * 16 * ATT2_ROW_BYTES;
#pragma unroll
            for (int kk = 0; kk < DK / 16; ++kk) {
                uint32_t a0, a1, a2, a3;
                ldsm_x4(qaddr + (((hc + kk * 2 + (lane >> 4)) ^ sw) << 4), a0, a1, a2, a3);
                const uint32_t kaddr = krow_u + (((hc + kk * 2 + ((lane >> 3) & 1)) ^ sw) << 4);
#pragma unroll
                for (int nb = 0; nb < KB; ++nb) {
                    uint32_t b0, b1, b2, b3;
                    ldsm_x4(kaddr + nb * 16 * ATT2_ROW_BYTES, b0, b1, b2, b3);
                    mma_bf16_16816(s[2 * nb], a0, a1, a2, a3, b0, b1);
                    mma_bf16_16816(s[2 * nb + 1], a0, a1, a2, a3, b2, b3);
                }
            }
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int n = 0; n < KB * 2; ++n) {
                if (n >= KB * 2 - 2 && n * 8 + 8 > Lk) {
                    const int j = n * 8 + 2 * t;
                    if (j >= Lk) s[n][0] = s[n][2] = -INFINITY;
                    if (j + 1 >= Lk) s[n][1] = s[n][3] = -INFINITY;
                }
                m0 = fmaxf(m0, fmaxf(s[n][0], s[n][1]));
                m1 = fmaxf(m1, fmaxf(s[n][2], s[n][3]));
            }
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
            m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
            m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
            const float o0 = m0 * scale_log2, o1 = m1 * scale_log2;
            float sum0 = 0.f, sum1 = 0.f;
            uint32_t pk[KB * 2][2];  // P rounded to bf16 right away: halves the live registers going into PV
#pragma unroll
            for (int n = 0; n < KB * 2; ++n) {
                const float e0 = ex2_approx(fmaf(s[n][0], scale_log2, -o0));
                const float e1 = ex2_approx(fmaf(s[n][1], scale_log2, -o0));
                const float e2 = ex2_approx(fmaf(s[n][2], scale_log2, -o1));
                const float e3 = ex2_approx(fmaf(s[n][3], scale_log2, -o1));
                sum0 += e0 + e1;
                sum1 += e2 + e3;
                pk[n][0] = pack_bf16x2(e0, e1);
                pk[n][1] = pack_bf16x2(e2, e3);
            }
            sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
            sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
            sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
            sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
            float o[DK / 8][4];
#pragma unroll
            for (int n = 0; n < DK / 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
            uint32_t vcol[DK / 16];
#pragma unroll
            for (int nb = 0; nb < DK / 16; ++nb) vcol[nb] = vrow_u + (((hc + nb * 2 + (lane >> 4)) ^ sw) << 4);
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
                for (int nb = 0; nb < DK / 16; ++nb) {
                    uint32_t b0, b1, b2, b3;
                    ldsm_x4_trans(vcol[nb] + kb * 16 * ATT2_ROW_BYTES, b0, b1, b2, b3);
                    mma_bf16_16816(o[2 * nb], pk[2 * kb][0], pk[2 * kb][1], pk[2 * kb + 1][0], pk[2 * kb + 1][1], b0, b1);
                    mma_bf16_16816(o[2 * nb + 1], pk[2 * kb][0], pk[2 * kb][1], pk[2 * kb + 1][0], pk[2 * kb + 1][1], b2, b3);
                }
            }
            const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
            const int col = grp * 64 + hh * DK + 2 * t;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int i = qt * 16 + g + half * 8;
                if (i < Lq && (i < p.q_rows[0] || p.out[1])) {  // a segment without an output is a conv halo only
                    __nv_bfloat16* orow = (i < p.q_rows[0])
                                              ? p.out[0] + ((size_t)clip * p.q_rows[0] + i) * p.out_ld[0]
                                              : p.out[1] + ((size_t)clip * p.q_rows[1] + (i - p.q_rows[0])) * p.out_ld[1];
                    const float inv = half ? inv1 : inv0;
#pragma unroll
                    for (int n = 0; n < DK / 8; ++n)
                        *reinterpret_cast<uint32_t*>(orow + col + n * 8) = pack_bf16x2(o[n][2 * half] * inv, o[n][2 * half + 1] * inv);
                }
            }
        };
        // The same tile restricted to `nkb` (<= PKB) 16-key blocks starting at block kb_lo: a partial task of the split tail
        // below.  Its normalised bf16 rows go into warp `sub`'s own (finished) Q tile of cv_q, the row statistics next to the
        // barriers; 128-B rows with the 16-byte chunks XOR-swizzled by the row, so the quad store pattern spreads over the banks.
        constexpr int PKB = (KB + 3) / 4;
        auto run_partial = [&](const int qt, const int kb_lo, const int nkb, const int sub) {
            float s[PKB * 2][4];
#pragma unroll
            for (int n = 0; n < PKB * 2; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
            const uint32_t qaddr = qrow_u + qt * 16 * ATT2_ROW_BYTES;
#pragma unroll
            for (int kk = 0; kk < DK / 16; ++kk) {
                uint32_t a0, a1, a2, a3;
                ldsm_x4(qaddr + (((kk * 2 + (lane >> 4)) ^ sw) << 4), a0, a1, a2, a3);
                const uint32_t kaddr = krow_u + (((kk * 2 + ((lane >> 3) & 1)) ^ sw) << 4) + kb_lo * 16 * ATT2_ROW_BYTES;
#pragma unroll
                for (int j = 0; j < PKB; ++j) {
                    if (j >= nkb) continue;  // warp-uniform
                    uint32_t b0, b1, b2, b3;
                    ldsm_x4(kaddr + j * 16 * ATT2_ROW_BYTES, b0, b1, b2, b3);
                    mma_bf16_16816(s[2 * j], a0, a1, a2, a3, b0, b1);
                    mma_bf16_16816(s[2 * j + 1], a0, a1, a2, a3, b2, b3);
                }
            }
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int n = 0; n < PKB * 2; ++n) {
                if (n >= 2 * nkb) continue;
                const int j = (kb_lo * 2 + n) * 8 + 2 * t;  // key index of s[n][0]
                if (j >= Lk) s[n][0] = s[n][2] = -INFINITY;
                if (j + 1 >= Lk) s[n][1] = s[n][3] = -INFINITY;
                m0 = fmaxf(m0, fmaxf(s[n][0], s[n][1]));
                m1 = fmaxf(m1, fmaxf(s[n][2], s[n][3]));
            }
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
            m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
            m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
            const float o0 = m0 * scale_log2, o1 = m1 * scale_log2;
            float sum0 = 0.f, sum1 = 0.f;
            uint32_t pk[PKB * 2][2];
#pragma unroll
            for (int n = 0; n < PKB * 2; ++n) {
                pk[n][0] = pk[n][1] = 0u;
                if (n >= 2 * nkb) continue;
                const float e0 = ex2_approx(fmaf(s[n][0], scale_log2, -o0));
                const float e1 = ex2_approx(fmaf(s[n][1], scale_log2, -o0));
                const float e2 = ex2_approx(fmaf(s[n][2], scale_log2, -o1));
                const float e3 = ex2_approx(fmaf(s[n][3], scale_log2, -o1));
                sum0 += e0 + e1;
                sum1 += e2 + e3;
                pk[n][0] = pack_bf16x2(e0, e1);
                pk[n][1] = pack_bf16x2(e2, e3);
            }
            sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
            sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
            sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
            sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
            float o[DK / 8][4];
#pragma unroll
            for (int n = 0; n < DK / 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
            for (int j = 0; j < PKB; ++j) {
                if (j >= nkb) continue;
#pragma unroll
                for (int nb = 0; nb < DK / 16; ++nb) {
                    uint32_t b0, b1, b2, b3;
                    ldsm_x4_trans(vrow_u + (((nb * 2 + (lane >> 4)) ^ sw) << 4) + (kb_lo + j) * 16 * ATT2_ROW_BYTES, b0, b1, b2, b3);
                    mma_bf16_16816(o[2 * nb], pk[2 * j][0], pk[2 * j][1], pk[2 * j + 1][0], pk[2 * j + 1][1], b0, b1);
                    mma_bf16_16816(o[2 * nb + 1], pk[2 * j][0], pk[2 * j][1], pk[2 * j + 1][0], pk[2 * j + 1][1], b2, b3);
                }
            }
            const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
            uint8_t* part = cv_q + sub * 16 * ATT2_ROW_BYTES;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int r = g + half * 8;
                const float inv = half ? inv1 : inv0;
#pragma unroll
                for (int n = 0; n < DK / 8; ++n)
                    *reinterpret_cast<uint32_t*>(part + r * ATT2_ROW_BYTES + ((n ^ (r & 7)) << 4) + 4 * t) =
                        pack_bf16x2(o[n][2 * half] * inv, o[n][2 * half + 1] * inv);
                if (t == 0) {
                    tail_stats[(sub * 16 + r) * 2] = half ? m1 : m0;
                    tail_stats[(sub * 16 + r) * 2 + 1] = half ? sum1 : sum0;
                }
            }
        };
        const int n_warps = nthr >> 5, n_tasks = HG * q_tiles;
        // One task more than warps (the 138-token joint attention of tedexp: nine 16-query tiles on eight warps): the last tile
        // would keep one warp busy for a whole extra round while seven wait at the barrier (25 % of the kernel's stall samples,
        // profiles/r02_ncu_attention_full.txt).  It is split by KEY ranges over four warps instead and the four partial
        // softmax results are merged (each warp merges 16 of the 64 output columns).
        const bool split_tail = TAIL_SPLIT && HG == 1 && KB >= 4 && n_tasks == n_warps + 1 && n_warps >= 4;
        for (int task = warp; task < (split_tail ? n_warps : n_tasks); task += n_warps) run_tile(task / q_tiles, task % q_tiles);
        if (split_tail && warp < 4) {
            const int qt = q_tiles - 1;
            run_partial(qt, (warp * KB) / 4, ((warp + 1) * KB) / 4 - (warp * KB) / 4, warp);
            asm volatile("bar.sync 1, 128;\n" ::: "memory");  // warps 0..3: partial rows and statistics are in shared memory
            const int r = lane & 15, h = lane >> 4;
            float mx = -INFINITY, ms[4], ls[4];
#pragma unroll
            for (int sidx = 0; sidx < 4; ++sidx) {
                ms[sidx] = tail_stats[(sidx * 16 + r) * 2];
                ls[sidx] = tail_stats[(sidx * 16 + r) * 2 + 1];
                mx = fmaxf(mx, ms[sidx]);
            }
            float wsum = 0.f, acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int sidx = 0; sidx < 4; ++sidx) {
                const float w = ls[sidx] * ex2_approx((ms[sidx] - mx) * scale_log2);
                wsum += w;
                const uint4 u = *reinterpret_cast<const uint4*>(cv_q + (sidx * 16 + r) * ATT2_ROW_BYTES + (((2 * warp + h) ^ (r & 7)) << 4));
                const uint32_t wd[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    acc[2 * e] = fmaf(w, __uint_as_float(wd[e] << 16), acc[2 * e]);
                    acc[2 * e + 1] = fmaf(w, __uint_as_float(wd[e] & 0xffff0000u), acc[2 * e + 1]);
                }
            }
            const int i = qt * 16 + r;
            if (i < Lq && (i < p.q_rows[0] || p.out[1])) {
                __nv_bfloat16* orow = (i < p.q_rows[0])
                                          ? p.out[0] + ((size_t)clip * p.q_rows[0] + i) * p.out_ld[0]
                                          : p.out[1] + ((size_t)clip * p.q_rows[1] + (i - p.q_rows[0])) * p.out_ld[1];
                const float inv = 1.0f / wsum;
                uint4 o4;
                o4.x = pack_bf16x2(acc[0] * inv, acc[1] * inv);
                o4.y = pack_bf16x2(acc[2] * inv, acc[3] * inv);
                o4.z = pack_bf16x2(acc[4] * inv, acc[5] * inv);
                o4.w = pack_bf16x2(acc[6] * inv, acc[7] * inv);
                *reinterpret_cast<uint4*>(orow + grp * 64 + warp * 16 + h * 8) = o4;
            }
        }
        __syncthreads();  // cv may be overwritten by the next item's conv
    }
    GD_TRACE_MARK(trace_slot, 8);
}

int make_rows_tmap(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    static PFN_encodeTiled encode = get_encode_tiled();
    if (!encode) return set_error(GD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(GD_ERR_CUDA, "cuTensorMapEncodeTiled (attention) failed (CUresult %d)", (int)r);
    return GD_OK;
}

// Conv arithmetic of the bf16 path: packed bf16x2 (default; in the chain -1.8 % tedexp / -5 % beat step time, final-pose
// error vs the reference 4.45e-3 -> 4.63e-3 tedexp, 4.58e-3 -> 5.32e-3 beat, tolerance 2e-2) or fp32 FMA (GD_ATTN_CONV=f32).
static bool conv_in_bf16() {
    const char* e = getenv("GD_ATTN_CONV");
    return !(e && e[0] == 'f');
}

template <int DK, int KB, bool CONV_BF16>
static int launch_attention_tma(const AttnParams& p, int n_clips, cudaStream_t s) {
    constexpr int HG = 64 / DK;
    // 8 warps x 128 registers: two CTAs fill the register file of an SM exactly.  (Round 2, measured and reverted: 9 warps x 112
    // registers for the 138-token joint attention - nine 16-query tiles, so on 8 warps the second MMA round runs one warp out
    // of eight - was SLOWER in the chain: attention 1.45 -> 1.60 ms per tedexp step; 40 bytes of spill and 18 instead of 16
    // warps fighting for the same issue slots cost more than the idle round.)
    constexpr int MAXT = 256;
    // short key ranges need fewer registers: 20 instead of 16 warps per SM; the beat windows (d_k = 32, <= 48 keys) compile to 80
    // registers without spilling: four CTAs per SM instead of three (at 128 clips per GPU every CTA then has ONE item)
    constexpr int MAXREG = (DK == 32 && KB <= 3) ? 80 : (KB <= 5 ? 96 : 128);
    // Split tail (one 16-query tile more than warps): OFF by default.  Measured on B200 in the tedexp chain (GD_ATTN_TAIL=1,
    // profiles/r02_ab_attention_tail_split.jsonl): 5.25-5.28 vs 5.13-5.15 ms/step - slower.  The warps that wait at the barrier
    // while one warp finishes the ninth tile are not lost time: the second resident CTA uses the issue slots, and the merge
    // plus 116 bytes of spill cost more than the idle round.  Same conclusion as the 9-warp variant.
    constexpr bool TAIL = (DK == 64 && KB >= 4);
    static const bool tail_on = getenv("GD_ATTN_TAIL") && getenv("GD_ATTN_TAIL")[0] == '1';
    auto kern = (TAIL && tail_on) ? dconv_attention_tma_kernel<DK, KB, MAXT, MAXREG, CONV_BF16, TAIL>
                                  : dconv_attention_tma_kernel<DK, KB, MAXT, MAXREG, CONV_BF16, false>;
    const int Lq_pad = (p.Lq + 15) & ~15, Lk_pad = KB * 16;
    Attn2Geom geo{};
    geo.groups = p.heads / HG;
    geo.n_items = n_clips * geo.groups;
    geo.raw_k_off = (Lq_pad + 2) * ATT2_ROW_BYTES;
    geo.raw_v_off = geo.raw_k_off + (Lk_pad + 2) * ATT2_ROW_BYTES;
    geo.raw_stage_bytes = geo.raw_v_off + (Lk_pad + 2) * ATT2_ROW_BYTES;
    const int cv_bytes = (Lq_pad + 2 * Lk_pad) * ATT2_ROW_BYTES;
    const int fixed_bytes = cv_bytes + 3 * DK * 4 * (int)sizeof(float) + 16 + 128;
    // a second raw stage (two items in flight) whenever four CTAs per SM still fit: the small-item launches
    // (34..40 tokens) are bound by the latency of their loads, not by shared-memory capacity
    geo.raw_stages = (2 * geo.raw_stage_bytes + fixed_bytes <= 56 * 1024) ? 2 : 1;
    geo.raw_bytes = geo.raw_stages * geo.raw_stage_bytes;
    geo.cv_q_off = geo.raw_bytes;
    geo.cv_k_off = geo.cv_q_off + Lq_pad * ATT2_ROW_BYTES;
    geo.cv_v_off = geo.cv_k_off + Lk_pad * ATT2_ROW_BYTES;
    geo.taps_off = geo.cv_v_off + Lk_pad * ATT2_ROW_BYTES;
    geo.bar_off = geo.taps_off + 3 * DK * 4 * (int)sizeof(float);
    geo.tx_bytes = (uint32_t)(p.Lq + 2 * p.Lk) * ATT2_ROW_BYTES;
    const size_t smem = (size_t)geo.bar_off + 16 + 512 /*tail-split row statistics*/ + 128 /*base alignment slack*/;
    CUtensorMap tq[2], tk[2], tv[2];
    for (int sgi = 0; sgi < 2; ++sgi) {
        const int has_q = p.q_rows[sgi] > 0, has_k = p.kv_rows[sgi] > 0;
        const int qs = has_q ? sgi : 0;
        int rc = make_rows_tmap(&tq[sgi], p.q[qs], (uint64_t)(n_clips - 1) * p.q_stride[qs] + p.q_rows[qs],
                                (uint64_t)p.heads * DK, p.q_ld[qs], p.q_rows[qs]);
        if (rc) return rc;
        rc = make_rows_tmap(&tk[sgi], has_k ? p.k[sgi] : p.k[0], (uint64_t)n_clips * p.kv_rows[has_k ? sgi : 0],
                            (uint64_t)p.heads * DK, p.kv_ld[has_k ? sgi : 0], p.kv_rows[has_k ? sgi : 0]);
        if (rc) return rc;
        rc = make_rows_tmap(&tv[sgi], has_k ? p.v[sgi] : p.v[0], (uint64_t)n_clips * p.kv_rows[has_k ? sgi : 0],
                            (uint64_t)p.heads * DK, p.kv_ld[has_k ? sgi : 0], p.kv_rows[has_k ? sgi : 0]);
        if (rc) return rc;
    }
    static size_t configured_dev[GD_MAX_DEVICES] = {};  // function attributes are per device
    size_t& configured = configured_dev[current_device()];
    if (smem > configured) {
        GD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
        GD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        configured = smem;
    }
    // warps: one per (head, 16-query tile) task, but enough of them to make the conv pass of K/V-heavy items quick
    const int tasks = HG * (Lq_pad / 16);
    int warps = tasks;
    const int conv_warps = (Lq_pad + 2 * Lk_pad) / 48;  // ~48 token rows of conv per warp
    if (warps < conv_warps) warps = conv_warps;
    if (warps < 4) warps = 4;
    if (warps > MAXT / 32) warps = MAXT / 32;
    int per_sm = 0;
    GD_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem));
    if (per_sm < 1) return set_error(GD_ERR_CUDA, "gd_dconv_attention: kernel does not fit on an SM (smem %zu B)", smem);
    int grid = per_sm * ((p.max_sms > 0 && p.max_sms < sm_count()) ? p.max_sms : sm_count());
    if (grid > geo.n_items) grid = geo.n_items;
    GD_CUDA_CHECK(launch_k(kern, grid, warps * 32, smem, s, 1, tq[0], tq[1], tk[0], tk[1], tv[0], tv[1], p, geo));
    count_launch();
    GD_CUDA_CHECK(cudaGetLastError());
    return GD_OK;
}

template <int DK>
static int dispatch_kb_tma(const AttnParams& p, int n_clips, cudaStream_t s) {
    const bool bf = conv_in_bf16();
    switch ((p.Lk + 15) / 16) {
        case 1: return bf ? launch_attention_tma<DK, 1, true>(p, n_clips, s) : launch_attention_tma<DK, 1, false>(p, n_clips, s);
        case 2: return bf ? launch_attention_tma<DK, 2, true>(p, n_clips, s) : launch_attention_tma<DK, 2, false>(p, n_clips, s);
        case 3: return bf ? launch_attention_tma<DK, 3, true>(p, n_clips, s) : launch_attention_tma<DK, 3, false>(p, n_clips, s);
        case 4: return bf ? launch_attention_tma<DK, 4, true>(p, n_clips, s) : launch_attention_tma<DK, 4, false>(p, n_clips, s);
        case 5: return bf ? launch_attention_tma<DK, 5, true>(p, n_clips, s) : launch_attention_tma<DK, 5, false>(p, n_clips, s);
        case 6: return bf ? launch_attention_tma<DK, 6, true>(p, n_clips, s) : launch_attention_tma<DK, 6, false>(p, n_clips, s);
        case 7: return bf ? launch_attention_tma<DK, 7, true>(p, n_clips, s) : launch_attention_tma<DK, 7, false>(p, n_clips, s);
        case 8: return bf ? launch_attention_tma<DK, 8, true>(p, n_clips, s) : launch_attention_tma<DK, 8, false>(p, n_clips, s);
        case 9: return bf ? launch_attention_tma<DK, 9, true>(p, n_clips, s) : launch_attention_tma<DK, 9, false>(p, n_clips, s);
        case 10: return bf ? launch_attention_tma<DK, 10, true>(p, n_clips, s) : launch_attention_tma<DK, 10, false>(p, n_clips, s);
    }
    return set_error(GD_ERR_INVALID, "gd_dconv_attention: need 0 < keys <= 160 (got %d)", p.Lk);
}

// GD_ATTN selects the kernel generation: v1 = one CTA per (clip, head), mma.sync; v2 (default) = persistent TMA-fed
// mma.sync kernel; v3 = warp-specialised tcgen05/TMEM pipeline for d_k = 64 (attention_tc.cu) - correct, but on B200 it
// only ties v2 on the 138-token joint attention (71 us) and loses on the shorter windows, so it stays opt-in.
static int attention_variant() {
    const char* e = getenv("GD_ATTN");
    if (e && e[0] == 'v' && e[1] >= '1' && e[1] <= '3') return e[1] - '0';
    return 2;
}

static int run_attention(const gd_attn_desc* d, bool fp32_in, void* stream) {
    if (!d) return set_error(GD_ERR_INVALID, "gd_dconv_attention: null descriptor");
    if (!d->q[0] || !d->k[0] || !d->v[0] || !d->out[0]) return set_error(GD_ERR_INVALID, "gd_dconv_attention: segment 0 missing");
    if (!d->conv_wq || !d->conv_bq || !d->conv_wk || !d->conv_bk || !d->conv_wv || !d->conv_bv)
        return set_error(GD_ERR_INVALID, "gd_dconv_attention: conv taps missing");
    if (d->n_clips <= 0 || d->heads <= 0) return set_error(GD_ERR_INVALID, "gd_dconv_attention: bad clip/head count");
    if (d->k[1] && !d->v[1]) return set_error(GD_ERR_INVALID, "gd_dconv_attention: segment 1 incomplete");
    KindScope kind_scope("attn");
    AttnParams p{};
    const int align = fp32_in ? 4 : 8;  // 16-byte row loads
    for (int s = 0; s < 2; ++s) {
        p.q[s] = d->q[s], p.k[s] = d->k[s], p.v[s] = d->v[s];
        p.out[s] = reinterpret_cast<__nv_bfloat16*>(d->out[s]);
        p.q_rows[s] = d->q[s] ? d->q_rows[s] : 0;
        p.kv_rows[s] = d->k[s] ? d->kv_rows[s] : 0;
        p.q_ld[s] = d->q_ld[s], p.kv_ld[s] = d->kv_ld[s], p.out_ld[s] = d->out_ld[s];
        p.q_stride[s] = d->q_clip_stride[s] > 0 ? d->q_clip_stride[s] : p.q_rows[s];
        if (p.q_stride[s] < p.q_rows[s]) return set_error(GD_ERR_INVALID, "gd_dconv_attention: q_clip_stride < q_rows");
        if ((d->q[s] && (d->q_ld[s] % align || (reinterpret_cast<uintptr_t>(d->q[s]) & 15) || d->out_ld[s] % 2)) ||
            (d->k[s] && (d->kv_ld[s] % align || ((reinterpret_cast<uintptr_t>(d->k[s]) | reinterpret_cast<uintptr_t>(d->v[s])) & 15))))
            return set_error(GD_ERR_INVALID, "gd_dconv_attention: q/k/v rows must be 16-byte aligned");
    }
    p.wq = d->conv_wq, p.bq = d->conv_bq, p.wk = d->conv_wk, p.bk = d->conv_bk, p.wv = d->conv_wv, p.bv = d->conv_bv;
    p.heads = d->heads;
    p.max_sms = d->max_ctas_sms;
    p.scale_log2 = d->scale * 1.4426950408889634f;
    p.Lq = p.q_rows[0] + p.q_rows[1];
    p.Lk = p.kv_rows[0] + p.kv_rows[1];
    if (p.Lq <= 0 || p.Lk <= 0 || p.Lk > 160 || p.Lq > 1024)
        return set_error(GD_ERR_INVALID, "gd_dconv_attention: need 0 < keys <= 160 (got %d), queries <= 1024", p.Lk);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    // bf16 rows whose segments fit a TMA box go to the persistent TMA-fed kernel
    const int variant = attention_variant();
    const bool halo = (p.q_rows[1] > 0 && !p.out[1]) || p.q_stride[0] != p.q_rows[0] || p.q_stride[1] != p.q_rows[1];
    if (!fp32_in && variant == 3 && !halo && d->d_k == 64 && p.q_rows[0] <= 256 && p.q_rows[1] <= 256 && p.kv_rows[0] <= 256 &&
        p.kv_rows[1] <= 256 && p.Lq <= 144 && attention_tc_smem_bytes(p) <= 232448)
        return launch_attention_tc(p, d->n_clips, s);
    const bool tma_ok = !fp32_in && variant >= 2 && p.q_rows[0] <= 256 && p.q_rows[1] <= 256 &&
                        p.Lq <= 256 && (64 % d->d_k) == 0 && d->heads % (64 / d->d_k) == 0;
    if (tma_ok) {
        if (d->d_k == 32) return dispatch_kb_tma<32>(p, d->n_clips, s);
        if (d->d_k == 64) return dispatch_kb_tma<64>(p, d->n_clips, s);
    }
    if (d->d_k == 32) return fp32_in ? dispatch_kb<float, 32>(p, d->n_clips, s) : dispatch_kb<__nv_bfloat16, 32>(p, d->n_clips, s);
    if (d->d_k == 64) return fp32_in ? dispatch_kb<float, 64>(p, d->n_clips, s) : dispatch_kb<__nv_bfloat16, 64>(p, d->n_clips, s);
    return set_error(GD_ERR_INVALID, "gd_dconv_attention: d_k=%d unsupported (32/64)", d->d_k);
}

#ifdef GD_TRACE
void set_trace_attention(unsigned long long* buf) { cudaMemcpyToSymbol(t_trace_buf, &buf, sizeof(buf)); }
#endif

}  // namespace gd

extern "C" int gd_dconv_attention(const gd_attn_desc* d, void* stream) { return gd::run_attention(d, false, stream); }
extern "C" int gd_dconv_attention_f32in(const gd_attn_desc* d, void* stream) { return gd::run_attention(d, true, stream); }
