// MultiDConvHeadAttention core on the 5th-generation tensor cores (d_k = 64, bf16 rows): tcgen05.mma for S = Q·Kᵀ and
// O = P·V with the accumulators in TMEM, one thread per query row for the softmax.
//
// Work item = (clip, head), persistent CTAs of 8 warps, two per SM.  Per item:
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
// Status: numerically validated against the fp32 reference (tests/test_kernels_gpu.py, GD_ATTN=v3) but NOT the default:
// on B200 it runs the 138-token joint attention of 256 clips in 86 us where the mma.sync kernel (attention.cu, v2) needs
// 74 us - five CTA-wide barriers and two MMA round trips per query tile leave the SM idle (ncu: 2.9 barrier-stall
// cycles per issued instruction), and the second query tile (10 valid rows of 128) costs a full round.  The design
// notes above are the starting point for the next iteration (both tiles in flight, softmax overlapped with the conv
// of the next item).
#include "attention_common.cuh"

namespace gd {

struct TcGeom {
    int n_items, Lq16, Lk16, n_tiles, kblocks;
    int raw_k_off, raw_v_off, raw_bytes;     // packed raw blocks, rows of 128 B
    int cvq_off, cvk_off, vt_off, bar_off;   // operand tiles (1024-B aligned), barriers
    uint32_t tx_bytes, tmem_cols, o_col, x_col;  // x_col: four spare TMEM columns for the row max / sum exchange
};

constexpr int TC_THREADS = 256;
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

        for (int tile = 0; tile < g.n_tiles; ++tile) {
            const int rows_valid = min(128, Lq - tile * 128);
            const bool last = tile == g.n_tiles - 1;
            // Eight warps share the 128 rows: warp w and w + 4 both own TMEM lane quadrant w & 3 (query rows 32(w&3)..+31);
            // warp w < 4 takes the first half of the 16-key chunks / of the output columns, warp w + 4 the second.  The two
            // partial row maxima and sums meet in four spare TMEM columns.
            const int quad = warp & 3, khalf = warp >> 2;
            const bool active = quad * 32 < rows_valid;
            const int row = quad * 32 + lane;  // TMEM lane = query row within the tile
            const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
            const int c_split = (nsk + 1) >> 1;
            const int c_lo = khalf ? c_split : 0, c_hi = khalf ? nsk : c_split;
            // ---------------- S = Q Kᵀ
            if (tid == TC_MMA_THREAD) {
                tc_fence_after_sync();
                const uint64_t da = umma_desc_k_sw128(smem_u32(cvq + tile * 16384));
                const uint64_t db = umma_desc_k_sw128(smem_u32(cvk));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base, da + 2 * k, db + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
                umma_commit(mma_bar);
            }
            __syncwarp();
            if (active) {
                mbar_wait(mma_bar, mphase);
                tc_fence_after_sync();
                // sweep 1: maximum over this thread's share of the valid keys
                float m = -INFINITY;
                for (int c = c_lo; c < c_hi; ++c) {
                    uint32_t v[16];
                    tmem_ld_32x32b_x16(t_lane + c * 16, v);
                    tmem_ld_wait();
                    const int nvalid = Lk - c * 16;  // >= 16 except in the last chunk
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (j < nvalid) m = fmaxf(m, __uint_as_float(v[j]));
                }
                tmem_st_32x32b_x1(t_lane + g.x_col + khalf, m);
                tmem_st_wait_tc();
                tc_fence_before_sync();
            }
            mphase ^= 1;
            __syncthreads();  // partial maxima published
            float part_sum = 0.f;
            if (active) {
                tc_fence_after_sync();
                const float m = fmaxf(tmem_ld_32x32b_x1(t_lane + g.x_col), tmem_ld_32x32b_x1(t_lane + g.x_col + 1));
                tmem_ld_wait();
                // sweep 2: p = exp2(s*c - m*c), row sum, P -> shared memory (bf16, K-major SWIZZLE_128B, 64 keys per block)
                const float mo = (m == -INFINITY) ? 0.f : m * c_log2;  // a half without valid keys contributes nothing
                uint8_t* prow = pbuf + row * ATT_ROW_BYTES;
                for (int c = c_lo; c < c_hi; ++c) {
                    uint32_t v[16];
                    tmem_ld_32x32b_x16(t_lane + c * 16, v);
                    tmem_ld_wait();
                    const int nvalid = Lk - c * 16;
                    float e[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        e[j] = j < nvalid ? ex2_approx(fmaf(__uint_as_float(v[j]), c_log2, -mo)) : 0.f;
                        part_sum += e[j];
                    }
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
            __syncthreads();  // P complete, every read of S retired, partial sums published
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
            // staging tile for coalesced stores: the P tile once it is dead, or (last query tile) the Q tile, so that
            // the raw/P region can already take the next item's TMA load
            uint8_t* stage = last ? cvq : pbuf;
            if (active) {
                mbar_wait(mma_bar, mphase);
                tc_fence_after_sync();
                const float inv_sum = 1.0f / (tmem_ld_32x32b_x1(t_lane + g.x_col + 2) + tmem_ld_32x32b_x1(t_lane + g.x_col + 3));
#pragma unroll
                for (int c2 = 0; c2 < 2; ++c2) {
                    const int c = khalf * 2 + c2;  // this warp's half of the 64 output columns
                    uint32_t v[16];
                    tmem_ld_32x32b_x16(t_lane + g.o_col + c * 16, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint4 o;
                        o.x = pack_bf16x2(__uint_as_float(v[8 * h + 0]) * inv_sum, __uint_as_float(v[8 * h + 1]) * inv_sum);
                        o.y = pack_bf16x2(__uint_as_float(v[8 * h + 2]) * inv_sum, __uint_as_float(v[8 * h + 3]) * inv_sum);
                        o.z = pack_bf16x2(__uint_as_float(v[8 * h + 4]) * inv_sum, __uint_as_float(v[8 * h + 5]) * inv_sum);
                        o.w = pack_bf16x2(__uint_as_float(v[8 * h + 6]) * inv_sum, __uint_as_float(v[8 * h + 7]) * inv_sum);
                        *reinterpret_cast<uint4*>(stage + row * ATT_ROW_BYTES + (((c * 2 + h) ^ (row & 7)) << 4)) = o;
                    }
                }
                tc_fence_before_sync();
                if (last && warp == 0) {
                    // P is dead (its MMA has retired): restore the zero padding rows it covered, then fetch the next item
                    const int zr[6] = {0, Lq + 1, 0, Lk + 1, 0, Lk + 1};
                    uint8_t* const zb[6] = {raw_q, raw_q, raw_k, raw_k, raw_v, raw_v};
                    for (int i = lane; i < 48; i += 32)
                        *reinterpret_cast<uint4*>(zb[i >> 3] + zr[i >> 3] * ATT_ROW_BYTES + (i & 7) * 16) = make_uint4(0, 0, 0, 0);
                    __syncwarp();
                    if (lane == 0 && item + (int)gridDim.x < g.n_items) issue_load(item + gridDim.x);
                }
            }
            mphase ^= 1;
            __syncthreads();  // staging tile complete
            for (int i = tid; i < rows_valid * 8; i += TC_THREADS) {
                const int r = i >> 3, ch = i & 7;
                const int qi = tile * 128 + r;
                __nv_bfloat16* orow = (qi < p.q_rows[0])
                                          ? p.out[0] + ((size_t)clip * p.q_rows[0] + qi) * p.out_ld[0]
                                          : p.out[1] + ((size_t)clip * p.q_rows[1] + (qi - p.q_rows[0])) * p.out_ld[1];
                *reinterpret_cast<uint4*>(orow + head * 64 + ch * 8) =
                    *reinterpret_cast<const uint4*>(stage + r * ATT_ROW_BYTES + ((ch ^ (r & 7)) << 4));
            }
            __syncthreads();  // staging tile, S and O are free again
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc_dyn(tmem_base, g.tmem_cols);
}

int launch_attention_tc(const AttnParams& p, int n_clips, cudaStream_t s) {
    TcGeom g{};
    g.n_items = n_clips * p.heads;
    g.Lq16 = (p.Lq + 15) & ~15, g.Lk16 = (p.Lk + 15) & ~15;
    g.n_tiles = (p.Lq + 127) / 128;
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
    int per_sm = 2;
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
