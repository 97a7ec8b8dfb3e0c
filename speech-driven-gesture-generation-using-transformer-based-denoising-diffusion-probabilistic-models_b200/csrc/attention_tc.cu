// MultiDConvHeadAttention core on the 5th-generation tensor cores (d_k = 64, bf16 rows, <= 144 queries, <= 160 keys):
// tcgen05.mma for S = Q·Kᵀ and O = P·V with the accumulators in TMEM, one thread per query row for the softmax, and a
// warp-specialised software pipeline over the work items (clip, head) - opt-in with GD_ATTN=v3.
//
// One CTA per SM, 16 warps, three concurrent stages connected by mbarriers, so that the latency chain of an item
// (TMA -> conv -> MMA -> softmax -> MMA -> epilogue) overlaps with its neighbours':
//   conv warps 8..13(+15)  raw(i) -> operand tiles cv[i&1]: depth-wise conv3 over tokens (fp32 FMA); Q, K -> K-major
//                          SWIZZLE_128B tiles, V -> Vᵀ tile (d_k rows, keys contiguous) so that both MMAs take plain
//                          K-major descriptors; then the TMA load of raw(i+1) is issued
//   control thread (w14)   S[i&1] = Q Kᵀ as soon as cv[i&1] is complete - one item ahead of the softmax -, O[i&1] = P V as
//                          soon as P(i) is complete; tcgen05.commit hands buffers back (cv, P) and on (S, O)
//   softmax warps 0..7     thread = TMEM lane = query row (warps w / w+4 split the keys and the output columns; partial
//                          row max / sum meet in spare TMEM columns): two sweeps of tcgen05.ld over S, P -> shared memory
//                          as the bf16 A operand; later O -> scaled by 1/sum -> 64-byte row segments to global
//   tail warp 15           query rows 128..143 (the 138-token joint attention has ten) with mma.sync from cv[i&1], instead of
//                          a second, almost empty 128-row tensor-core tile
// Shared memory: raw 54 KB + 2 x 61 KB operand tiles + 48 KB P = 221 KB for the joint attention; TMEM: S and O
// double-buffered (2 x 160 + 2 x 64 columns) + 8 exchange columns.
//
// Status: validated against the fp32 reference (tests/test_kernels_gpu.py) but NOT the default.  B200, 256 clips x 8
// heads per launch (v2 = mma.sync kernel in attention.cu): pose 34 tokens 33 us (v2 21), memory 104 tokens 52 us (v2 47),
// joint 138 tokens 71 us (v2 71), 34x138 52 us (v2 38).  The matrix products are free here, but per item the conv
// (~6.8 k warp-instructions) and the softmax sweeps (~5.6 k) cost as much as the whole mma.sync kernel (13.7 k), the
// single raw buffer exposes one TMA round trip per item on the conv stage, and at 34..138 tokens the fixed per-item
// latencies dominate.  What would make it win: P kept in TMEM as the A operand (frees 48 KB for a second raw buffer),
// packed bf16 conv arithmetic, several heads per 128-row tile for the short windows.
#include "attention_common.cuh"

namespace gd {


__device__ __forceinline__ void tmem_alloc_dyn(uint32_t* smem_result, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_result)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_dyn(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void tmem_st_32x32b_x1(uint32_t taddr, float v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};\n" ::"r"(taddr), "f"(v) : "memory");
}
__device__ __forceinline__ float tmem_ld_32x32b_x1(uint32_t taddr) {
    float v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n" : "=f"(v) : "r"(taddr) : "memory");
    return v;
}
__device__ __forceinline__ void tmem_st_wait_tc() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

struct Taps4 {
    float4 w0, w1, w2, b;  // taps / bias of four consecutive channels
};
// conv.weight is [d_k, 3]: twelve consecutive floats hold the three taps of four channels
__device__ __forceinline__ Taps4 load_taps4(const float* w, const float* b, int c) {
    const float4 t0 = __ldg(reinterpret_cast<const float4*>(w + c * 3));
    const float4 t1 = __ldg(reinterpret_cast<const float4*>(w + c * 3) + 1);
    const float4 t2 = __ldg(reinterpret_cast<const float4*>(w + c * 3) + 2);
    Taps4 t;
    t.w0 = make_float4(t0.x, t0.w, t1.z, t2.y);
    t.w1 = make_float4(t0.y, t1.x, t1.w, t2.z);
    t.w2 = make_float4(t0.z, t1.y, t2.x, t2.w);
    t.b = __ldg(reinterpret_cast<const float4*>(b + c));
    return t;
}

// conv3 of 16 tokens x 4 channels into a K-major SWIZZLE_128B tile whose rows are tokens (Q and K operands)
__device__ __forceinline__ void conv16_rows(const uint8_t* src, uint8_t* dst, int chunk, const Taps4& t) {
    uint2 raw[18];
#pragma unroll
    for (int s = 0; s < 18; ++s) raw[s] = *reinterpret_cast<const uint2*>(src + s * ATT_ROW_BYTES);
    float4 prev = unpack_bf16x4(raw[0]), cur = unpack_bf16x4(raw[1]);
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        const float4 nxt = unpack_bf16x4(raw[s + 2]);
        uint2 o;
        o.x = pack_bf16x2(fmaf(t.w0.x, prev.x, fmaf(t.w1.x, cur.x, fmaf(t.w2.x, nxt.x, t.b.x))),
                          fmaf(t.w0.y, prev.y, fmaf(t.w1.y, cur.y, fmaf(t.w2.y, nxt.y, t.b.y))));
        o.y = pack_bf16x2(fmaf(t.w0.z, prev.z, fmaf(t.w1.z, cur.z, fmaf(t.w2.z, nxt.z, t.b.z))),
                          fmaf(t.w0.w, prev.w, fmaf(t.w1.w, cur.w, fmaf(t.w2.w, nxt.w, t.b.w))));
        *reinterpret_cast<uint2*>(dst + s * ATT_ROW_BYTES + ((chunk ^ (s & 7)) << 4)) = o;  // token row = p0 + s, p0 % 16 == 0
        prev = cur;
        cur = nxt;
    }
}

// conv3 of 16 tokens x 4 channels written TRANSPOSED: Vᵀ tile, row = channel, 16 consecutive keys = two 16-B chunks of
// key block p0/64.  Keys >= Lk are written as zeros (their P is zero; this keeps 0 * x finite).
__device__ __forceinline__ void conv16_vt(const uint8_t* src, uint8_t* vt, int p0, int c, int Lk, const Taps4& t) {
    uint2 raw[18];
#pragma unroll
    for (int s = 0; s < 18; ++s) raw[s] = *reinterpret_cast<const uint2*>(src + s * ATT_ROW_BYTES);
    uint32_t w[4][8];  // per channel: eight bf16x2 words = sixteen consecutive keys
    float4 prev = unpack_bf16x4(raw[0]), cur = unpack_bf16x4(raw[1]);
    float4 even = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        const float4 nxt = unpack_bf16x4(raw[s + 2]);
        float4 r;
        r.x = fmaf(t.w0.x, prev.x, fmaf(t.w1.x, cur.x, fmaf(t.w2.x, nxt.x, t.b.x)));
        r.y = fmaf(t.w0.y, prev.y, fmaf(t.w1.y, cur.y, fmaf(t.w2.y, nxt.y, t.b.y)));
        r.z = fmaf(t.w0.z, prev.z, fmaf(t.w1.z, cur.z, fmaf(t.w2.z, nxt.z, t.b.z)));
        r.w = fmaf(t.w0.w, prev.w, fmaf(t.w1.w, cur.w, fmaf(t.w2.w, nxt.w, t.b.w)));
        if (p0 + s >= Lk) r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s & 1) {
            w[0][s >> 1] = pack_bf16x2(even.x, r.x);
            w[1][s >> 1] = pack_bf16x2(even.y, r.y);
            w[2][s >> 1] = pack_bf16x2(even.z, r.z);
            w[3][s >> 1] = pack_bf16x2(even.w, r.w);
        } else {
            even = r;
        }
        prev = cur;
        cur = nxt;
    }
    uint8_t* blk = vt + (p0 >> 6) * 8192;
    const int ch0 = (p0 & 63) >> 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = c + i;
        uint8_t* rp = blk + row * ATT_ROW_BYTES;
        *reinterpret_cast<uint4*>(rp + ((ch0 ^ (row & 7)) << 4)) = make_uint4(w[i][0], w[i][1], w[i][2], w[i][3]);
        *reinterpret_cast<uint4*>(rp + (((ch0 + 1) ^ (row & 7)) << 4)) = make_uint4(w[i][4], w[i][5], w[i][6], w[i][7]);
    }
}

__device__ __forceinline__ void softmax_warps_sync() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }  // warps 0..7
__device__ __forceinline__ void ldsm_x4_tc(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void mma_16816_tc(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Query rows 128..143 (the 138-token joint attention has ten of them) would cost a second, almost empty 128-row tensor
// core tile; one extra warp does them with mma.sync straight from the same operand tiles instead, concurrently with the
// tcgen05 tile: 16-key blocks with a running max / sum, K and Vᵀ both read as [n][k] B operands through ldmatrix.
__device__ __forceinline__ void tail_rows_mma_sync(const uint8_t* cvq, const uint8_t* cvk, const uint8_t* vt, int nsk, int Lq,
                                                   int Lk, float c_log2, int lane, const AttnParams& p, int clip, int head) {
    const int g = lane >> 2, t = lane & 3, sw = lane & 7;
    uint32_t a[4][4];
    const uint32_t qrow = smem_u32(cvq) + (128 + (lane & 15)) * ATT_ROW_BYTES;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) ldsm_x4_tc(qrow + (((kk * 2 + (lane >> 4)) ^ sw) << 4), a[kk][0], a[kk][1], a[kk][2], a[kk][3]);
    const uint32_t krow = smem_u32(cvk) + ((lane & 7) + (lane >> 4) * 8) * ATT_ROW_BYTES;
    const uint32_t vrow = smem_u32(vt) + ((lane & 7) + (lane >> 4) * 8) * ATT_ROW_BYTES;
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, sum0 = 0.f, sum1 = 0.f;
    for (int nb = 0; nb < nsk; ++nb) {
        float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4_tc(krow + nb * 16 * ATT_ROW_BYTES + (((kk * 2 + ((lane >> 3) & 1)) ^ sw) << 4), b0, b1, b2, b3);
            mma_16816_tc(s0, a[kk], b0, b1);
            mma_16816_tc(s1, a[kk], b2, b3);
        }
        const int j = nb * 16 + 2 * t;  // keys j, j+1 (s0) and j+8, j+9 (s1)
        if (j >= Lk) s0[0] = s0[2] = -INFINITY;
        if (j + 1 >= Lk) s0[1] = s0[3] = -INFINITY;
        if (j + 8 >= Lk) s1[0] = s1[2] = -INFINITY;
        if (j + 9 >= Lk) s1[1] = s1[3] = -INFINITY;
        float b0m = fmaxf(fmaxf(s0[0], s0[1]), fmaxf(s1[0], s1[1])), b1m = fmaxf(fmaxf(s0[2], s0[3]), fmaxf(s1[2], s1[3]));
        b0m = fmaxf(b0m, __shfl_xor_sync(0xffffffffu, b0m, 1));
        b0m = fmaxf(b0m, __shfl_xor_sync(0xffffffffu, b0m, 2));
        b1m = fmaxf(b1m, __shfl_xor_sync(0xffffffffu, b1m, 1));
        b1m = fmaxf(b1m, __shfl_xor_sync(0xffffffffu, b1m, 2));
        const float n0 = fmaxf(m0, b0m), n1 = fmaxf(m1, b1m);  // finite from the first block on (it holds valid keys)
        const float r0 = ex2_approx((m0 - n0) * c_log2), r1 = ex2_approx((m1 - n1) * c_log2);
        m0 = n0, m1 = n1;
        sum0 *= r0, sum1 *= r1;
#pragma unroll
        for (int n = 0; n < 8; ++n) o[n][0] *= r0, o[n][1] *= r0, o[n][2] *= r1, o[n][3] *= r1;
        const float f0 = m0 * c_log2, f1 = m1 * c_log2;
        const float e00 = ex2_approx(fmaf(s0[0], c_log2, -f0)), e01 = ex2_approx(fmaf(s0[1], c_log2, -f0));
        const float e02 = ex2_approx(fmaf(s0[2], c_log2, -f1)), e03 = ex2_approx(fmaf(s0[3], c_log2, -f1));
        const float e10 = ex2_approx(fmaf(s1[0], c_log2, -f0)), e11 = ex2_approx(fmaf(s1[1], c_log2, -f0));
        const float e12 = ex2_approx(fmaf(s1[2], c_log2, -f1)), e13 = ex2_approx(fmaf(s1[3], c_log2, -f1));
        sum0 += (e00 + e01) + (e10 + e11);
        sum1 += (e02 + e03) + (e12 + e13);
        const uint32_t pa[4] = {pack_bf16x2(e00, e01), pack_bf16x2(e02, e03), pack_bf16x2(e10, e11), pack_bf16x2(e12, e13)};
        // P·V for this 16-key block: Vᵀ rows are channels, keys are contiguous (64 keys per 8-KB block)
        const uint32_t vblk = vrow + ((nb * 16) >> 6) * 8192;
        const int kc = ((nb * 16) & 63) >> 3;
#pragma unroll
        for (int nd = 0; nd < 4; ++nd) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4_tc(vblk + nd * 16 * ATT_ROW_BYTES + (((kc + ((lane >> 3) & 1)) ^ sw) << 4), b0, b1, b2, b3);
            mma_16816_tc(o[2 * nd], pa, b0, b1);
            mma_16816_tc(o[2 * nd + 1], pa, b2, b3);
        }
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int qi = 128 + g + half * 8;
        if (qi < Lq) {
            __nv_bfloat16* orow = (qi < p.q_rows[0]) ? p.out[0] + ((size_t)clip * p.q_rows[0] + qi) * p.out_ld[0]
                                                     : p.out[1] + ((size_t)clip * p.q_rows[1] + (qi - p.q_rows[0])) * p.out_ld[1];
            const float inv = half ? inv1 : inv0;
#pragma unroll
            for (int n = 0; n < 8; ++n)
                *reinterpret_cast<uint32_t*>(orow + head * 64 + n * 8 + 2 * t) = pack_bf16x2(o[n][2 * half] * inv, o[n][2 * half + 1] * inv);
        }
    }
}

struct TcpGeom {
    int n_items, Lq16, Lk16, kblocks;
    int raw_k_off, raw_v_off, raw_bytes;
    int cv_off, cv_stride, cvk_rel, vt_rel;  // operand tile set b at cv_off + b * cv_stride: [cvq | cvk | vt]
    int p_off, bar_off;
    uint32_t tx_bytes;
};
constexpr int TCP_THREADS = 512;
constexpr int TCP_CONV_WARPS = 6;
constexpr uint32_t TCP_S_COL = 0, TCP_S_STRIDE = 160, TCP_O_COL = 320, TCP_O_STRIDE = 64, TCP_X_COL = 448;


__global__ void __launch_bounds__(TCP_THREADS, 1)
dconv_attention_tc_kernel(const __grid_constant__ CUtensorMap tm_q0, const __grid_constant__ CUtensorMap tm_q1,
                           const __grid_constant__ CUtensorMap tm_k0, const __grid_constant__ CUtensorMap tm_k1,
                           const __grid_constant__ CUtensorMap tm_v0, const __grid_constant__ CUtensorMap tm_v1,
                           const AttnParams p, const TcpGeom g) {
    extern __shared__ __align__(1024) uint8_t smem_tc[];
    uint8_t* smem = smem_tc;
    uint8_t* raw_q = smem;
    uint8_t* raw_k = smem + g.raw_k_off;
    uint8_t* raw_v = smem + g.raw_v_off;
    uint8_t* pbuf = smem + g.p_off;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.bar_off);
    uint64_t* raw_full = bars;          // TMA -> conv warps  (bars[1], bars[2] unused)
    uint64_t* cv_full = bars + 3;       // [2] conv warps -> control, tail
    uint64_t* cv_empty = bars + 5;      // [2] P·V retired (+ tail warp done) -> conv warps
    uint64_t* s_full = bars + 7;        // [2] Q·Kᵀ retired -> softmax warps
    uint64_t* s_empty = bars + 9;       // [2] softmax warps have read S -> control
    uint64_t* p_full = bars + 11;       // softmax warps wrote P -> control
    uint64_t* p_empty = bars + 12;      // P·V retired -> softmax warps
    uint64_t* o_full = bars + 13;       // [2] P·V retired -> softmax warps
    uint64_t* o_empty = bars + 15;      // [2] softmax warps have read O -> control
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Lq = p.Lq, Lk = p.Lk;
    const bool has_tail = Lq > 128;
    const int n_conv_warps = has_tail ? TCP_CONV_WARPS : TCP_CONV_WARPS + 1;  // warp 15 convolves when there are no tail rows

    if (tid == 0) {
        if (smem_u32(smem) & 1023) __trap();
        prefetch_tensormap(&tm_q0), prefetch_tensormap(&tm_k0), prefetch_tensormap(&tm_v0);
        if (p.q_rows[1]) prefetch_tensormap(&tm_q1);
        if (p.kv_rows[1]) prefetch_tensormap(&tm_k1), prefetch_tensormap(&tm_v1);
        mbar_init(raw_full, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&cv_full[b], n_conv_warps);
            mbar_init(&cv_empty[b], has_tail ? 2 : 1);
            mbar_init(&s_full[b], 1);
            mbar_init(&s_empty[b], 8);
            mbar_init(&o_full[b], 1);
            mbar_init(&o_empty[b], 8);
        }
        mbar_init(p_full, 8);
        mbar_init(p_empty, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_dyn(tmem_slot, 512);
    for (int i = tid; i < g.raw_bytes / 16; i += TCP_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before_sync();
    pdl_launch_dependents();
    pdl_wait();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int nsq = g.Lq16 / 16, nsk = g.Lk16 / 16;
    const float c_log2 = p.scale_log2;
    const int n_mine = (g.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // items of this CTA

    auto issue_load = [&](int item) {  // one thread; token rows start at row 1 of each raw block
        const int clip = item / p.heads, col0 = (item % p.heads) * 64;
        mbar_arrive_expect_tx(raw_full, g.tx_bytes);
        tma_load_2d(raw_q + ATT_ROW_BYTES, &tm_q0, raw_full, col0, clip * p.q_rows[0]);
        if (p.q_rows[1]) tma_load_2d(raw_q + (1 + p.q_rows[0]) * ATT_ROW_BYTES, &tm_q1, raw_full, col0, clip * p.q_rows[1]);
        tma_load_2d(raw_k + ATT_ROW_BYTES, &tm_k0, raw_full, col0, clip * p.kv_rows[0]);
        tma_load_2d(raw_v + ATT_ROW_BYTES, &tm_v0, raw_full, col0, clip * p.kv_rows[0]);
        if (p.kv_rows[1]) {
            tma_load_2d(raw_k + (1 + p.kv_rows[0]) * ATT_ROW_BYTES, &tm_k1, raw_full, col0, clip * p.kv_rows[1]);
            tma_load_2d(raw_v + (1 + p.kv_rows[0]) * ATT_ROW_BYTES, &tm_v1, raw_full, col0, clip * p.kv_rows[1]);
        }
    };

    if ((warp >= 8 && warp < 8 + TCP_CONV_WARPS) || (warp == 15 && !has_tail)) {
        // ------------------------------------------------------------------ conv warps
        const int cw = warp == 15 ? TCP_CONV_WARPS : warp - 8;
        const int ctid = cw * 32 + lane, cthreads = n_conv_warps * 32;
        const int units = (nsq + 2 * nsk) * 16;
        if (ctid == 0 && n_mine > 0) issue_load(blockIdx.x);
        for (int k = 0; k < n_mine; ++k) {
            const int b = k & 1;
            const uint32_t ph2 = (k >> 1) & 1;
            uint8_t* cvq = smem + g.cv_off + b * g.cv_stride;
            uint8_t* cvk = cvq + g.cvk_rel;
            uint8_t* vt = cvq + g.vt_rel;
            mbar_wait(raw_full, k & 1);
            mbar_wait(&cv_empty[b], ph2 ^ 1);  // the MMAs (and the tail warp) of item k-2 are done with this tile set
            for (int it = ctid; it < units; it += cthreads) {
                const int hc4 = it & 15, sg = it >> 4;
                const int which = sg < nsq ? 0 : (sg < nsq + nsk ? 1 : 2);
                const int p0 = (sg - (which == 0 ? 0 : (which == 1 ? nsq : nsq + nsk))) * 16;
                const Taps4 t = load_taps4(which == 0 ? p.wq : (which == 1 ? p.wk : p.wv),
                                           which == 0 ? p.bq : (which == 1 ? p.bk : p.bv), hc4 * 4);
                const uint8_t* src = (which == 0 ? raw_q : (which == 1 ? raw_k : raw_v)) + p0 * ATT_ROW_BYTES + hc4 * 8;
                if (which == 2)
                    conv16_vt(src, vt, p0, hc4 * 4, Lk, t);
                else
                    conv16_rows(src, (which == 0 ? cvq : cvk) + p0 * ATT_ROW_BYTES + (hc4 & 1) * 8, hc4 >> 1, t);
            }
            fence_proxy_async();  // ordinary stores -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&cv_full[b]);
            asm volatile("bar.sync 2, %0;\n" ::"r"(cthreads) : "memory");  // every conv warp has finished reading raw
            if (ctid == 0 && k + 1 < n_mine) issue_load(blockIdx.x + (k + 1) * gridDim.x);
        }
    } else if (warp == 14) {
        // ------------------------------------------------------------------ control thread: both matrix products
        if (lane == 0) {
            const uint32_t idesc_qk = umma_idesc_bf16(128, g.Lk16), idesc_pv = umma_idesc_bf16(128, 64);
            auto issue_qk = [&](int k) {
                const int b = k & 1;
                const uint32_t ph2 = (k >> 1) & 1;
                uint8_t* cvq = smem + g.cv_off + b * g.cv_stride;
                mbar_wait(&cv_full[b], ph2);
                mbar_wait(&s_empty[b], ph2 ^ 1);
                tc_fence_after_sync();
                const uint64_t da = umma_desc_k_sw128(smem_u32(cvq));
                const uint64_t db = umma_desc_k_sw128(smem_u32(cvq + g.cvk_rel));
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16_ss(tmem_base + TCP_S_COL + b * TCP_S_STRIDE, da + 2 * kk, db + 2 * kk, idesc_qk, kk != 0 ? 1u : 0u);
                umma_commit(&s_full[b]);
            };
            if (n_mine > 0) issue_qk(0);
            for (int k = 0; k < n_mine; ++k) {
                const int b = k & 1;
                const uint32_t ph2 = (k >> 1) & 1;
                if (k + 1 < n_mine) issue_qk(k + 1);  // keeps the softmax warps one item ahead of P·V
                uint8_t* vt = smem + g.cv_off + b * g.cv_stride + g.vt_rel;
                mbar_wait(p_full, k & 1);
                mbar_wait(&o_empty[b], ph2 ^ 1);
                tc_fence_after_sync();
                for (int ks = 0; ks < nsk; ++ks) {
                    const uint64_t da = umma_desc_k_sw128(smem_u32(pbuf + (ks >> 2) * 16384)) + 2 * (ks & 3);
                    const uint64_t db = umma_desc_k_sw128(smem_u32(vt + (ks >> 2) * 8192)) + 2 * (ks & 3);
                    umma_bf16_ss(tmem_base + TCP_O_COL + b * TCP_O_STRIDE, da, db, idesc_pv, ks != 0 ? 1u : 0u);
                }
                umma_commit(&o_full[b]);   // each commit fires once every MMA issued so far has retired
                umma_commit(p_empty);
                umma_commit(&cv_empty[b]);
            }
        }
    } else if (warp == 15) {
        // ------------------------------------------------------------------ tail warp: query rows 128..143
        {
            for (int k = 0; k < n_mine; ++k) {
                const int b = k & 1;
                const int item = blockIdx.x + k * gridDim.x;
                uint8_t* cvq = smem + g.cv_off + b * g.cv_stride;
                mbar_wait(&cv_full[b], (k >> 1) & 1);
                tail_rows_mma_sync(cvq, cvq + g.cvk_rel, cvq + g.vt_rel, nsk, Lq, Lk, c_log2, lane, p, item / p.heads, item % p.heads);
                __syncwarp();
                if (lane == 0) mbar_arrive(&cv_empty[b]);
            }
        }
    } else if (warp < 8) {
        // ------------------------------------------------------------------ softmax / output warps
        const int rows_valid = min(128, Lq);
        const int quad = warp & 3, khalf = warp >> 2;
        const int row = quad * 32 + lane;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        const int c_split = (nsk + 1) >> 1;
        const int c_lo = khalf ? c_split : 0, c_hi = khalf ? nsk : c_split;
        const bool active = quad * 32 < rows_valid;  // warps without valid rows still take part in every barrier
        for (int k = 0; k < n_mine; ++k) {
            const int b = k & 1;
            const uint32_t ph2 = (k >> 1) & 1;
            const int item = blockIdx.x + k * gridDim.x;
            const int clip = item / p.heads, head = item % p.heads;
            const uint32_t s_addr = t_lane + TCP_S_COL + b * TCP_S_STRIDE;
            const uint32_t x_addr = t_lane + TCP_X_COL + b * 4;
            mbar_wait(&s_full[b], ph2);
            tc_fence_after_sync();
            float m = -INFINITY;
            if (active) {
                for (int c = c_lo; c < c_hi; ++c) {
                    uint32_t v[16];
                    tmem_ld_32x32b_x16(s_addr + c * 16, v);
                    tmem_ld_wait();
                    if (c * 16 + 16 <= Lk) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) m = fmaxf(m, __uint_as_float(v[j]));
                    } else {
                        const int nvalid = Lk - c * 16;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (j < nvalid) m = fmaxf(m, __uint_as_float(v[j]));
                    }
                }
            }
            tmem_st_32x32b_x1(x_addr + khalf, m);
            tmem_st_wait_tc();
            tc_fence_before_sync();
            softmax_warps_sync();  // both halves of every row have published their maximum
            tc_fence_after_sync();
            const float max_a = tmem_ld_32x32b_x1(x_addr), max_b = tmem_ld_32x32b_x1(x_addr + 1);
            tmem_ld_wait();
            m = fmaxf(max_a, max_b);
            mbar_wait(p_empty, (k & 1) ^ 1);  // P·V of the previous item has retired: the P tile may be overwritten
            float part_sum = 0.f;
            if (active) {
                const float mo = m * c_log2;
                uint8_t* prow = pbuf + row * ATT_ROW_BYTES;
                for (int c = c_lo; c < c_hi; ++c) {
                    uint32_t v[16];
                    tmem_ld_32x32b_x16(s_addr + c * 16, v);
                    tmem_ld_wait();
                    float e[16];
                    if (c * 16 + 16 <= Lk) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) e[j] = ex2_approx(fmaf(__uint_as_float(v[j]), c_log2, -mo));
                    } else {
                        const int nvalid = Lk - c * 16;
#pragma unroll
                        for (int j = 0; j < 16; ++j) e[j] = j < nvalid ? ex2_approx(fmaf(__uint_as_float(v[j]), c_log2, -mo)) : 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) part_sum += e[j];
                    uint8_t* blk = prow + (c >> 2) * 16384;
                    const int cc = (c & 3) * 2;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint4 o;
                        o.x = pack_bf16x2(e[8 * h + 0], e[8 * h + 1]);
                        o.y = pack_bf16x2(e[8 * h + 2], e[8 * h + 3]);
                        o.z = pack_bf16x2(e[8 * h + 4], e[8 * h + 5]);
                        o.w = pack_bf16x2(e[8 * h + 6], e[8 * h + 7]);
                        *reinterpret_cast<uint4*>(blk + (((cc + h) ^ (row & 7)) << 4)) = o;
                    }
                }
            }
            tmem_st_32x32b_x1(x_addr + 2 + khalf, part_sum);
            tmem_st_wait_tc();
            fence_proxy_async();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&s_empty[b]);
                mbar_arrive(p_full);
            }
            // ---- output: O[b] is complete once P·V has retired
            mbar_wait(&o_full[b], ph2);
            tc_fence_after_sync();
            const float sum_a = tmem_ld_32x32b_x1(x_addr + 2), sum_b = tmem_ld_32x32b_x1(x_addr + 3);
            tmem_ld_wait();
            const float inv_sum = 1.0f / (sum_a + sum_b);
            __nv_bfloat16* orow = nullptr;
            if (active && row < rows_valid)
                orow = ((row < p.q_rows[0]) ? p.out[0] + ((size_t)clip * p.q_rows[0] + row) * p.out_ld[0]
                                            : p.out[1] + ((size_t)clip * p.q_rows[1] + (row - p.q_rows[0])) * p.out_ld[1]) +
                       head * 64 + khalf * 32;
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) {
                uint32_t v[16];
                tmem_ld_32x32b_x16(t_lane + TCP_O_COL + b * TCP_O_STRIDE + (khalf * 2 + c2) * 16, v);
                tmem_ld_wait();
                if (orow) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint4 o;
                        o.x = pack_bf16x2(__uint_as_float(v[8 * h + 0]) * inv_sum, __uint_as_float(v[8 * h + 1]) * inv_sum);
                        o.y = pack_bf16x2(__uint_as_float(v[8 * h + 2]) * inv_sum, __uint_as_float(v[8 * h + 3]) * inv_sum);
                        o.z = pack_bf16x2(__uint_as_float(v[8 * h + 4]) * inv_sum, __uint_as_float(v[8 * h + 5]) * inv_sum);
                        o.w = pack_bf16x2(__uint_as_float(v[8 * h + 6]) * inv_sum, __uint_as_float(v[8 * h + 7]) * inv_sum);
                        *reinterpret_cast<uint4*>(orow + c2 * 16 + h * 8) = o;
                    }
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_empty[b]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc_dyn(tmem_base, 512);
}

// shared memory the pipelined kernel needs for this shape (it runs one CTA per SM: <= 227 KB)
size_t attention_tc_smem_bytes(const AttnParams& p) {
    const int Lq16 = (p.Lq + 15) & ~15, Lk16 = (p.Lk + 15) & ~15, kblocks = (Lk16 + 63) / 64;
    const int raw = (p.Lq + 2 + 2 * (p.Lk + 2)) * ATT_ROW_BYTES;
    int cv_stride = (Lq16 + Lk16) * ATT_ROW_BYTES + kblocks * 8192;
    if (cv_stride < 16384) cv_stride = 16384;
    return (size_t)((raw + 1023) & ~1023) + 2 * (size_t)cv_stride + (size_t)kblocks * 16384 + 256;
}

int launch_attention_tc(const AttnParams& p, int n_clips, cudaStream_t s) {
    TcpGeom g{};
    g.n_items = n_clips * p.heads;
    g.Lq16 = (p.Lq + 15) & ~15, g.Lk16 = (p.Lk + 15) & ~15;
    g.kblocks = (g.Lk16 + 63) / 64;
    if (g.Lk16 > (int)TCP_S_STRIDE) return set_error(GD_ERR_INVALID, "gd_dconv_attention: more than 160 keys");
    g.raw_k_off = (p.Lq + 2) * ATT_ROW_BYTES;
    g.raw_v_off = g.raw_k_off + (p.Lk + 2) * ATT_ROW_BYTES;
    g.raw_bytes = g.raw_v_off + (p.Lk + 2) * ATT_ROW_BYTES;
    g.cv_off = (g.raw_bytes + 1023) & ~1023;
    g.cvk_rel = g.Lq16 * ATT_ROW_BYTES;
    g.vt_rel = g.cvk_rel + g.Lk16 * ATT_ROW_BYTES;
    g.cv_stride = g.vt_rel + g.kblocks * 8192;
    if (g.cv_stride < 16384) g.cv_stride = 16384;  // the query tile is read as 128 rows
    g.p_off = g.cv_off + 2 * g.cv_stride;
    g.bar_off = g.p_off + g.kblocks * 16384;
    g.tx_bytes = (uint32_t)(p.Lq + 2 * p.Lk) * ATT_ROW_BYTES;
    const size_t smem = (size_t)g.bar_off + 256;
    if (smem > 232448) return set_error(GD_ERR_CUDA, "gd_dconv_attention: pipelined tcgen05 kernel needs %zu B of shared memory", smem);
    if (reinterpret_cast<uintptr_t>(p.wq) & 15 || reinterpret_cast<uintptr_t>(p.wk) & 15 || reinterpret_cast<uintptr_t>(p.wv) & 15 ||
        reinterpret_cast<uintptr_t>(p.bq) & 15 || reinterpret_cast<uintptr_t>(p.bk) & 15 || reinterpret_cast<uintptr_t>(p.bv) & 15)
        return set_error(GD_ERR_INVALID, "gd_dconv_attention: conv taps must be 16-byte aligned");
    CUtensorMap tq[2], tk[2], tv[2];
    for (int sgi = 0; sgi < 2; ++sgi) {
        const int has_q = p.q_rows[sgi] > 0, has_k = p.kv_rows[sgi] > 0;
        int rc = make_rows_tmap(&tq[sgi], has_q ? p.q[sgi] : p.q[0], (uint64_t)n_clips * p.q_rows[has_q ? sgi : 0],
                                (uint64_t)p.heads * 64, p.q_ld[has_q ? sgi : 0], p.q_rows[has_q ? sgi : 0]);
        if (rc) return rc;
        rc = make_rows_tmap(&tk[sgi], has_k ? p.k[sgi] : p.k[0], (uint64_t)n_clips * p.kv_rows[has_k ? sgi : 0],
                            (uint64_t)p.heads * 64, p.kv_ld[has_k ? sgi : 0], p.kv_rows[has_k ? sgi : 0]);
        if (rc) return rc;
        rc = make_rows_tmap(&tv[sgi], has_k ? p.v[sgi] : p.v[0], (uint64_t)n_clips * p.kv_rows[has_k ? sgi : 0],
                            (uint64_t)p.heads * 64, p.kv_ld[has_k ? sgi : 0], p.kv_rows[has_k ? sgi : 0]);
        if (rc) return rc;
    }
    static size_t configured_dev[GD_MAX_DEVICES] = {};  // function attributes are per device
    size_t& configured = configured_dev[current_device()];
    if (smem > configured) {
        GD_CUDA_CHECK(cudaFuncSetAttribute(dconv_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
        configured = smem;
    }
    int grid = sm_count();
    if (grid > g.n_items) grid = g.n_items;
    GD_CUDA_CHECK(launch_k(dconv_attention_tc_kernel, grid, TCP_THREADS, smem, s, 1, tq[0], tq[1], tk[0], tk[1], tv[0], tv[1], p, g));
    count_launch();
    GD_CUDA_CHECK(cudaGetLastError());
    return GD_OK;
}

}  // namespace gd
