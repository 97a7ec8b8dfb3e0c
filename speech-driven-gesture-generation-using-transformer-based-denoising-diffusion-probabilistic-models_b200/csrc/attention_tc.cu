// MultiDConvHeadAttention core on the 5th-generation tensor cores (d_k = 64, bf16 rows): tcgen05.mma for S = Q·Kᵀ and
// O = P·V with the accumulators in TMEM, one thread per query row for the softmax.
//
// Work item = (clip, head), persistent CTAs of 9 warps, two per SM.  Per item:
//   TMA      raw Q/K/V row blocks -> shared memory (dense 128-B rows, [zero | L tokens | zero] per tensor)
//   conv     depth-wise conv3 over tokens (fp32 FMA): Q, K -> K-major SWIZZLE_128B operand tiles, V -> Vᵀ tile (d_k rows,
//            keys contiguous) so that both MMAs take plain K-major descriptors
//   S        one thread issues 4 x tcgen05.mma (M = 128 query rows, N = keys padded to 16, K = 16) -> TMEM
//   softmax  a thread owns TMEM lane r = query row r (warps w and w+4 split the keys of quadrant w&3): two sweeps of
//            tcgen05.ld (row max, then exp / sum), partial results meet in spare TMEM columns; no shuffles, no ldmatrix,
//            no per-warp MMA fragments; P goes to shared memory as the bf16 A operand
//   O        keys/16 x tcgen05.mma (N = 64) -> TMEM; the row's thread scales by 1/sum, a swizzled staging tile makes the
//            global stores full 128-B rows
// The raw blocks are dead after the conv and the P tile is needed only between the two MMAs, so P aliases the raw
// region (115 KB per CTA for the 138-token joint attention instead of 164 KB: that is what keeps two CTAs on an SM);
// the next item's TMA load is issued as soon as the last P·V has retired.
// Status: numerically validated against the fp32 reference (tests/test_kernels_gpu.py, GD_ATTN=v3) but NOT the default.
// Measured on B200, 256 clips x 8 heads per launch (v2 = mma.sync kernel in attention.cu):
//     pose 34 tokens 36 us (v2 21), memory 104 tokens 61 us (v2 49), joint 138 tokens 75 us (v2 74), 34x138 59 us (v2 41)
// The matrix products cost nothing here, but an item is a chain of dependent latencies - TMA, conv, MMA, TMEM loads,
// shared-memory P, MMA, TMEM loads - with three CTA-wide barriers, and at 34..138 tokens there is too little work per
// item to hide it with two CTAs per SM (ncu: 2.3-6.3 barrier-stall cycles per issued instruction).  Next step for this
// kernel: several items in flight per CTA (conv'd operand tiles double-buffered, P kept in TMEM as the A operand).
#include "attention_common.cuh"

namespace gd {

struct TcGeom {
    int n_items, Lq16, Lk16, n_tiles, kblocks;
    int raw_k_off, raw_v_off, raw_bytes;     // packed raw blocks, rows of 128 B
    int cvq_off, cvk_off, vt_off, bar_off;   // operand tiles (1024-B aligned), barriers
    uint32_t tx_bytes, tmem_cols, o_col, x_col;  // x_col: four spare TMEM columns for the row max / sum exchange
};

constexpr int TC_THREADS = 288;   // 8 warps for the tcgen05 tile + warp 8 for query rows 128..143 (mma.sync)
constexpr int TC_MMA_THREAD = 128;  // warp 4, lane 0

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

__device__ __forceinline__ float bf16_lane(const uint2& u, int i) {  // i-th of the four bf16 packed in u
    const uint32_t w = (i < 2) ? u.x : u.y;
    return __uint_as_float((i & 1) ? (w & 0xffff0000u) : (w << 16));
}

// conv3 of 16 tokens x 4 channels written TRANSPOSED: Vᵀ tile, row = channel, 16 consecutive keys = two 16-B chunks of
// key block p0/64.  Keys >= Lk are written as zeros (their P is zero; this keeps 0 * x finite).
__device__ __forceinline__ void conv16_vt(const uint8_t* src, uint8_t* vt, int p0, int c, int Lk, const Taps4& t) {
    uint2 raw[18];
#pragma unroll
    for (int s = 0; s < 18; ++s) raw[s] = *reinterpret_cast<const uint2*>(src + s * ATT_ROW_BYTES);
    const float w0[4] = {t.w0.x, t.w0.y, t.w0.z, t.w0.w}, w1[4] = {t.w1.x, t.w1.y, t.w1.z, t.w1.w};
    const float w2[4] = {t.w2.x, t.w2.y, t.w2.z, t.w2.w}, bb[4] = {t.b.x, t.b.y, t.b.z, t.b.w};
    uint8_t* blk = vt + (p0 >> 6) * 8192;
    const int ch0 = (p0 & 63) >> 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float r[16];
#pragma unroll
        for (int s = 0; s < 16; ++s) {
            r[s] = fmaf(w0[i], bf16_lane(raw[s], i), fmaf(w1[i], bf16_lane(raw[s + 1], i), fmaf(w2[i], bf16_lane(raw[s + 2], i), bb[i])));
            if (p0 + s >= Lk) r[s] = 0.f;
        }
        const int row = c + i;
        uint8_t* rp = blk + row * ATT_ROW_BYTES;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint4 o;
            o.x = pack_bf16x2(r[8 * h + 0], r[8 * h + 1]);
            o.y = pack_bf16x2(r[8 * h + 2], r[8 * h + 3]);
            o.z = pack_bf16x2(r[8 * h + 4], r[8 * h + 5]);
            o.w = pack_bf16x2(r[8 * h + 6], r[8 * h + 7]);
            *reinterpret_cast<uint4*>(rp + (((ch0 + h) ^ (row & 7)) << 4)) = o;
        }
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

__global__ void __launch_bounds__(TC_THREADS, 2)
dconv_attention_tc_kernel(const __grid_constant__ CUtensorMap tm_q0, const __grid_constant__ CUtensorMap tm_q1,
                          const __grid_constant__ CUtensorMap tm_k0, const __grid_constant__ CUtensorMap tm_k1,
                          const __grid_constant__ CUtensorMap tm_v0, const __grid_constant__ CUtensorMap tm_v1,
                          const AttnParams p, const TcGeom g) {
    extern __shared__ __align__(1024) uint8_t smem_tc[];
    uint8_t* smem = smem_tc;
    uint8_t* raw_q = smem;
    uint8_t* raw_k = smem + g.raw_k_off;
    uint8_t* raw_v = smem + g.raw_v_off;
    uint8_t* pbuf = smem;  // P (and the staging tile of non-final query tiles) alias the raw blocks
    uint8_t* cvq = smem + g.cvq_off;
    uint8_t* cvk = smem + g.cvk_off;
    uint8_t* vt = smem + g.vt_off;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + g.bar_off);
    uint64_t* mma_bar = full_bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Lq = p.Lq, Lk = p.Lk;

    if (tid == 0) {
        if (smem_u32(smem) & 1023) __trap();  // the swizzled operand tiles rely on a 1024-B aligned base
        prefetch_tensormap(&tm_q0), prefetch_tensormap(&tm_k0), prefetch_tensormap(&tm_v0);
        if (p.q_rows[1]) prefetch_tensormap(&tm_q1);
        if (p.kv_rows[1]) prefetch_tensormap(&tm_k1), prefetch_tensormap(&tm_v1);
        mbar_init(full_bar, 1);
        mbar_init(mma_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_dyn(tmem_slot, g.tmem_cols);
    for (int i = tid; i < g.raw_bytes / 16; i += TC_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before_sync();
    pdl_launch_dependents();
    pdl_wait();  // set-up touched only shared memory / TMEM; Q/K/V come from the previous kernel
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    auto issue_load = [&](int item) {  // one thread; token rows start at row 1 of each raw block
        const int clip = item / p.heads, col0 = (item % p.heads) * 64;
        mbar_arrive_expect_tx(full_bar, g.tx_bytes);
        tma_load_2d(raw_q + ATT_ROW_BYTES, &tm_q0, full_bar, col0, clip * p.q_rows[0]);
        if (p.q_rows[1]) tma_load_2d(raw_q + (1 + p.q_rows[0]) * ATT_ROW_BYTES, &tm_q1, full_bar, col0, clip * p.q_rows[1]);
        tma_load_2d(raw_k + ATT_ROW_BYTES, &tm_k0, full_bar, col0, clip * p.kv_rows[0]);
        tma_load_2d(raw_v + ATT_ROW_BYTES, &tm_v0, full_bar, col0, clip * p.kv_rows[0]);
        if (p.kv_rows[1]) {
            tma_load_2d(raw_k + (1 + p.kv_rows[0]) * ATT_ROW_BYTES, &tm_k1, full_bar, col0, clip * p.kv_rows[1]);
            tma_load_2d(raw_v + (1 + p.kv_rows[0]) * ATT_ROW_BYTES, &tm_v1, full_bar, col0, clip * p.kv_rows[1]);
        }
    };

    int item = blockIdx.x;
    if (tid == 0 && item < g.n_items) issue_load(item);
    uint32_t lphase = 0, mphase = 0;
    const uint32_t idesc_qk = umma_idesc_bf16(128, g.Lk16), idesc_pv = umma_idesc_bf16(128, 64);
    const int nsq = g.Lq16 / 16, nsk = g.Lk16 / 16;
    const int units = (nsq + 2 * nsk) * 16;
    const float c_log2 = p.scale_log2;

    for (; item < g.n_items; item += gridDim.x) {
        const int clip = item / p.heads, head = item % p.heads;
        mbar_wait(full_bar, lphase);
        lphase ^= 1;
        // ---------------- depth-wise conv3 over tokens, raw -> operand tiles
        for (int it = tid; it < units; it += TC_THREADS) {
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
        fence_proxy_async();  // operand tiles were written by ordinary stores; the tensor core reads them through the async proxy
        __syncthreads();

        // Eight warps share the 128 rows of the tensor-core tile: warp w and w + 4 both own TMEM lane quadrant w & 3 (query
        // rows 32(w&3)..+31); warp w < 4 takes the first half of the 16-key chunks / of the output columns, warp w + 4 the
        // second; the two partial row maxima and sums meet in four spare TMEM columns.  Warp 8 does rows 128..143.
        const int rows_valid = min(128, Lq);
        const int quad = warp & 3, khalf = (warp >> 2) & 1;
        const bool active = warp < 8 && quad * 32 < rows_valid;
        const int row = quad * 32 + lane;  // TMEM lane = query row
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        const int c_split = (nsk + 1) >> 1;
        const int c_lo = khalf ? c_split : 0, c_hi = khalf ? nsk : c_split;
        // ---------------- S = Q Kᵀ (rows 0..127)
        if (tid == TC_MMA_THREAD) {
            tc_fence_after_sync();
            const uint64_t da = umma_desc_k_sw128(smem_u32(cvq));
            const uint64_t db = umma_desc_k_sw128(smem_u32(cvk));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base, da + 2 * k, db + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
            umma_commit(mma_bar);
        }
        __syncwarp();
        if (warp == 8 && Lq > 128) tail_rows_mma_sync(cvq, cvk, vt, nsk, Lq, Lk, c_log2, lane, p, clip, head);
        if (active) {
            mbar_wait(mma_bar, mphase);
            tc_fence_after_sync();
            // sweep 1: maximum over this thread's share of the valid keys
            float m = -INFINITY;
            for (int c = c_lo; c < c_hi; ++c) {
                uint32_t v[16];
                tmem_ld_32x32b_x16(t_lane + c * 16, v);
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
            tmem_st_32x32b_x1(t_lane + g.x_col + khalf, m);
            tmem_st_wait_tc();
            tc_fence_before_sync();
        }
        mphase ^= 1;
        if (warp < 8) softmax_warps_sync();  // partial maxima published (the tail warp is not part of this barrier)
        if (active) {
            tc_fence_after_sync();
            const float m = fmaxf(tmem_ld_32x32b_x1(t_lane + g.x_col), tmem_ld_32x32b_x1(t_lane + g.x_col + 1));
            tmem_ld_wait();
            // sweep 2: p = exp2(s*c - m*c), row sum, P -> shared memory (bf16, K-major SWIZZLE_128B, 64 keys per block)
            const float mo = m * c_log2;
            float part_sum = 0.f;
            uint8_t* prow = pbuf + row * ATT_ROW_BYTES;
            for (int c = c_lo; c < c_hi; ++c) {
                uint32_t v[16];
                tmem_ld_32x32b_x16(t_lane + c * 16, v);
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
            tmem_st_32x32b_x1(t_lane + g.x_col + 2 + khalf, part_sum);
            tmem_st_wait_tc();
            fence_proxy_async();
            tc_fence_before_sync();
        }
        if (warp < 8) softmax_warps_sync();  // P complete, every read of S retired, partial sums published
        // ---------------- O = P V
        if (tid == TC_MMA_THREAD) {
            tc_fence_after_sync();
            for (int ks = 0; ks < nsk; ++ks) {
                const uint64_t da = umma_desc_k_sw128(smem_u32(pbuf + (ks >> 2) * 16384)) + 2 * (ks & 3);
                const uint64_t db = umma_desc_k_sw128(smem_u32(vt + (ks >> 2) * 8192)) + 2 * (ks & 3);
                umma_bf16_ss(tmem_base + g.o_col, da, db, idesc_pv, ks != 0 ? 1u : 0u);
            }
            umma_commit(mma_bar);
        }
        __syncwarp();
        if (active) {
            mbar_wait(mma_bar, mphase);
            tc_fence_after_sync();
            if (warp == 0) {
                // P is dead (its MMA has retired): restore the zero padding rows it covered, then fetch the next item
                const int zr[6] = {0, Lq + 1, 0, Lk + 1, 0, Lk + 1};
                uint8_t* const zb[6] = {raw_q, raw_q, raw_k, raw_k, raw_v, raw_v};
                for (int i = lane; i < 48; i += 32)
                    *reinterpret_cast<uint4*>(zb[i >> 3] + zr[i >> 3] * ATT_ROW_BYTES + (i & 7) * 16) = make_uint4(0, 0, 0, 0);
                __syncwarp();
                if (lane == 0 && item + (int)gridDim.x < g.n_items) issue_load(item + gridDim.x);
            }
            const float inv_sum = 1.0f / (tmem_ld_32x32b_x1(t_lane + g.x_col + 2) + tmem_ld_32x32b_x1(t_lane + g.x_col + 3));
            // this thread's 32 of the row's 64 output columns: 64 contiguous bytes = two full 32-B sectors
            __nv_bfloat16* orow = nullptr;
            if (row < rows_valid)
                orow = ((row < p.q_rows[0]) ? p.out[0] + ((size_t)clip * p.q_rows[0] + row) * p.out_ld[0]
                                            : p.out[1] + ((size_t)clip * p.q_rows[1] + (row - p.q_rows[0])) * p.out_ld[1]) +
                       head * 64 + khalf * 32;
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) {
                uint32_t v[16];
                tmem_ld_32x32b_x16(t_lane + g.o_col + (khalf * 2 + c2) * 16, v);
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
        }
        mphase ^= 1;
        __syncthreads();  // operand tiles, S and O are free for the next item
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc_dyn(tmem_base, g.tmem_cols);
}

int launch_attention_tc(const AttnParams& p, int n_clips, cudaStream_t s) {
    TcGeom g{};
    g.n_items = n_clips * p.heads;
    g.Lq16 = (p.Lq + 15) & ~15, g.Lk16 = (p.Lk + 15) & ~15;
    g.n_tiles = 1;  // rows 0..127 on the tensor cores, rows 128..143 by the mma.sync warp
    g.kblocks = (g.Lk16 + 63) / 64;
    // raw blocks: [zero | L rows | zero] per tensor, packed; the conv of the last 16-token segment may read up to 15 rows
    // past a block (finite data of the next block / tile: results land in padded rows that are masked or zeroed)
    g.raw_k_off = (p.Lq + 2) * ATT_ROW_BYTES;
    g.raw_v_off = g.raw_k_off + (p.Lk + 2) * ATT_ROW_BYTES;
    g.raw_bytes = g.raw_v_off + (p.Lk + 2) * ATT_ROW_BYTES;
    const int p_bytes = g.kblocks * 16384;
    int region0 = g.raw_bytes > p_bytes ? g.raw_bytes : p_bytes;
    region0 = (region0 + 1023) & ~1023;
    g.cvq_off = region0;
    g.cvk_off = g.cvq_off + g.Lq16 * ATT_ROW_BYTES;
    g.vt_off = g.cvk_off + g.Lk16 * ATT_ROW_BYTES;
    g.bar_off = g.vt_off + g.kblocks * 8192;
    // every query tile is read as 128 rows (and the last one doubles as a 16-KB staging tile): keep that in bounds
    if (g.bar_off < g.cvq_off + g.n_tiles * 16384) g.bar_off = g.cvq_off + g.n_tiles * 16384;
    g.tx_bytes = (uint32_t)(p.Lq + 2 * p.Lk) * ATT_ROW_BYTES;
    g.o_col = (uint32_t)((g.Lk16 + 31) & ~31);
    g.x_col = g.o_col + 64;
    uint32_t cols = 32;
    while (cols < g.x_col + 4) cols <<= 1;
    g.tmem_cols = cols;
    // the mbarriers live in the padding behind the raw blocks when there is room (the 138-token joint attention then
    // fits two CTAs per SM to the byte: 2 x (115 712 + 1 024) = 233 472)
    size_t smem = (size_t)g.bar_off + 64;
    if (region0 - g.raw_bytes >= 64 && g.raw_bytes >= p_bytes) {
        smem = (size_t)g.bar_off;
        g.bar_off = g.raw_bytes;
    }
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
    static size_t configured = 0;
    if (smem > configured) {
        GD_CUDA_CHECK(cudaFuncSetAttribute(dconv_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
        GD_CUDA_CHECK(cudaFuncSetAttribute(dconv_attention_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                           cudaSharedmemCarveoutMaxShared));
        configured = smem;
    }
    // CTAs per SM: 128 registers x 256 threads -> two by the register file; shared memory (228 KB per SM, 1 KB reserved
    // per CTA) and TMEM (512 columns per SM) can only lower that
    int per_sm = 2;  // 96 registers x 288 threads
    if ((smem + 1024) * 2 > 233472) per_sm = 1;
    if (per_sm * (int)g.tmem_cols > 512) per_sm = 512 / (int)g.tmem_cols;
    if (smem + 1024 > 233472) return set_error(GD_ERR_CUDA, "gd_dconv_attention: tcgen05 kernel does not fit on an SM (smem %zu B)", smem);
    int grid = per_sm * sm_count();
    if (grid > g.n_items) grid = g.n_items;
    GD_CUDA_CHECK(launch_k(dconv_attention_tc_kernel, grid, TC_THREADS, smem, s, 1, tq[0], tq[1], tk[0], tk[1], tv[0], tv[1], p, g));
    count_launch();
    GD_CUDA_CHECK(cudaGetLastError());
    return GD_OK;
}

}  // namespace gd
