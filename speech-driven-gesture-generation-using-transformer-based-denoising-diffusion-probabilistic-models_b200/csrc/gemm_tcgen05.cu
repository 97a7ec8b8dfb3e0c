// Persistent, warp-specialised bf16 GEMM for sm_100a:  D[M,N] = A[M,K] · W[N,K]ᵀ  (+ fused epilogue)
//
//   warp 0      TMA producer   (cp.async.bulk.tensor 2-D, SWIZZLE_128B tiles, mbarrier ring)
//   warp 1      MMA issuer     (one thread: tcgen05.mma cta_group::1 kind::f16, 128 x BN x 16)
//   warp 2      TMEM allocator (512 columns = two BN-wide fp32 accumulator stages for BN <= 256)
//   warps 4..11 epilogue       (tcgen05.ld 32x32b -> registers -> bias/activation -> swizzled smem -> TMA store)
//
// The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile
// i+1; K is short on this workload (128..2048), so that overlap is where the time is.  Eight epilogue warps
// (two per scheduler; warp w owns TMEM lane quadrant w%4 and column half w/4) write 32x32 chunks into a
// 128B/64B-swizzled staging tile and hand them to the TMA engine: plain tensor stores for bf16 outputs, and
// `cp.reduce.async.bulk.tensor ... add.f32` for the in-place residual GEMMs (H += A·Wᵀ + b), so the residual is
// added inside L2 and the SMs never load it.  Odd cases (positional rowbias, out-of-place residual, dual
// outputs) take the simple per-row path (MODE_DIRECT).
// MODE_DDPM turns the epilogue of the final projection into the DDPM ancestral update
// (gd_b200.h: gd_linear_ddpm): eps stays in registers.
#include "common.cuh"
#include "ddpm_math.cuh"
#include "host_util.h"
#include <cstdlib>

namespace gd {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 384;  // 4 control warps + 8 epilogue warps
constexpr int EPI_WARPS = 8;
constexpr int MODE_DIRECT = 0;    // per-row stores straight from registers (any epilogue feature)
constexpr int MODE_DDPM = 1;      // final projection + DDPM update
constexpr int MODE_TMA_BF16 = 2;  // bf16 output through TMA stores
constexpr int MODE_TMA_F32 = 3;   // fp32 output through TMA stores, or TMA reduce-add when accumulating in place
constexpr int MODE_CONV = 4;      // implicit-GEMM convolution over pixel rows (gd_conv_taps_bf16): shifted A tiles per tap
// Split-precision convolutions with operand reuse: a pipeline stage holds the tiles of one (tap, 64-channel block) and the
// MMA issuer forms all three bf16x3 products from them, so no tile is loaded twice (-1/3 shared-memory fill):
constexpr int MODE_CONV_S1 = 5;   // rows [hi(32) | lo(32)]: 1 A tile, B tiles [Whi|Whi] and [Wlo|0]      -> A*B0 + A*B1
constexpr int MODE_CONV_S2 = 6;
// staging bytes per epilogue warp of the fp32 (reduce-add) epilogue: 4096 = one 32x32 fp32 chunk, 8192 = a ring of two
#ifndef GD_F32_STAGING
#define GD_F32_STAGING 4096
#endif   // rows [hi(c) | lo(c)], c % 64 == 0: A tiles hi, lo; B tiles Whi, Wlo  -> hi*Whi + lo*Whi + hi*Wlo

struct GemmParams {
    int M, N, K;
    const float* bias;
    const float* rowbias;
    int rowbias_period, rowbias_offset;
    const float* residual;
    int ldr;
    int act;
    float* out_f32;
    int ldo_f32;
    __nv_bfloat16* out_bf16;
    int ldo_bf16;
    int reduce_add;  // MODE_TMA_F32: 1 = out += result (in-place residual), 0 = out = result
    int max_ctas;    // host only: cap of the persistent grid (0 = every SM)
    gd_ddpm_desc ddpm;
    // MODE_CONV: k-block kb reads A rows m0 + tap_shift[kb / kb_per_tap], columns ((kb % kb_per_tap) * 64) mod a_cols
    int kb_per_tap, a_cols;
    int split_out, c_store;  // split_out: channels [0, c_store) stored as [hi(c_store) | lo(c_store)] bf16 pairs
    int tap_shift[GD_CONV_MAX_TAPS];
    const float* scale;   // per output channel, applied after bias (+ReLU): folded BatchNorm
    const float* shift;
    int relu;
    int grid_h, grid_w;   // pixel grid of one image (rows per image = grid_h * grid_w)
    int y0, y1, x0, x1, stride;
    int out_img_stride, out_y_stride, out_x_stride, out_offset;  // output row of a kept pixel
};

template <int BN, int CL, int NA = 1, int NB = 1, int STG = 4096>
struct GemmCfg {
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_BYTES = (BN / CL) * BLOCK_K * 2;  // a CTA of a pair stores half of the W tile
    static constexpr int STAGE_BYTES = NA * A_BYTES + NB * B_BYTES;  // NA / NB tiles per stage (MODE_CONV_S*: operand reuse)
    // per epilogue warp one 32-row x 32-column fp32 staging tile.  A deeper ring and a shared-memory copy of the bias were
    // measured and changed nothing (the epilogue is not waiting on them) while costing a pipeline stage, so: 4 KB / warp.
    static constexpr int STAGING_PER_WARP = STG;
    static constexpr int STAGING_BYTES = EPI_WARPS * STAGING_PER_WARP;
    static constexpr int SMEM_LIMIT = 227 * 1024;
    static constexpr int STAGES_FIT = (SMEM_LIMIT - 2048 - STAGING_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*barriers*/ + STAGING_BYTES + 1024 /*align slack*/;
    static constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
};

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == GD_ACT_RELU2) {
        float r = fmaxf(v, 0.0f);
        return r * r;
    }
    if (act == GD_ACT_SILU) return v / (1.0f + __expf(-v));
    return v;
}

__device__ __forceinline__ void add_bias_chunk(float (&r)[32], const uint32_t (&v)[32], const float4 (&b)[8], bool has_bias) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        r[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + (has_bias ? b[j].x : 0.f);
        r[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + (has_bias ? b[j].y : 0.f);
        r[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + (has_bias ? b[j].z : 0.f);
        r[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + (has_bias ? b[j].w : 0.f);
    }
}

// Positional row bias of the TMA epilogues: r[j] += rowbias[(row % period + offset) * N + col0 + j]  (emb_x / emb_mem + PE)
__device__ __forceinline__ void add_rowbias_chunk(float (&r)[32], const GemmParams& p, int row, int col0) {
    const int pos = (row % p.rowbias_period) + p.rowbias_offset;
    const float4* b4 = reinterpret_cast<const float4*>(p.rowbias + (size_t)pos * p.N + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 b = __ldg(b4 + j);
        r[4 * j + 0] += b.x, r[4 * j + 1] += b.y, r[4 * j + 2] += b.z, r[4 * j + 3] += b.w;
    }
}

// MODE_DIRECT: one thread = one output row, 32 consecutive columns starting at col0, every epilogue feature.
__device__ __forceinline__ void epilogue_direct_chunk(const GemmParams& p, int row, int col0, uint32_t (&v)[32]) {
    float r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = __uint_as_float(v[j]);
    if (p.bias) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 b = __ldg(b4 + j);
            r[4 * j + 0] += b.x, r[4 * j + 1] += b.y, r[4 * j + 2] += b.z, r[4 * j + 3] += b.w;
        }
    }
    if (p.rowbias) {
        int pos = (row % p.rowbias_period) + p.rowbias_offset;
        const float4* b4 = reinterpret_cast<const float4*>(p.rowbias + (size_t)pos * p.N + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 b = __ldg(b4 + j);
            r[4 * j + 0] += b.x, r[4 * j + 1] += b.y, r[4 * j + 2] += b.z, r[4 * j + 3] += b.w;
        }
    }
    if (p.act != GD_ACT_NONE) {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = apply_act(r[j], p.act);
    }
    if (p.residual) {
        const float4* q4 = reinterpret_cast<const float4*>(p.residual + (size_t)row * p.ldr + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 q = __ldcg(q4 + j);
            r[4 * j + 0] += q.x, r[4 * j + 1] += q.y, r[4 * j + 2] += q.z, r[4 * j + 3] += q.w;
        }
    }
    if (p.out_f32) {
        float4* o4 = reinterpret_cast<float4*>(p.out_f32 + (size_t)row * p.ldo_f32 + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) o4[j] = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    }
    if (p.out_bf16) {
        uint4* o4 = reinterpret_cast<uint4*>(p.out_bf16 + (size_t)row * p.ldo_bf16 + col0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint4 w;
            w.x = pack_bf16x2(r[8 * j + 0], r[8 * j + 1]);
            w.y = pack_bf16x2(r[8 * j + 2], r[8 * j + 3]);
            w.z = pack_bf16x2(r[8 * j + 4], r[8 * j + 5]);
            w.w = pack_bf16x2(r[8 * j + 6], r[8 * j + 7]);
            o4[j] = w;
        }
    }
}

// MODE_CONV: v = (relu?)(acc + bias) * scale + shift -> bf16, 32 channels of one kept pixel (64 contiguous bytes)
__device__ __forceinline__ void epilogue_conv_chunk(const GemmParams& p, size_t out_row, int col0, const uint32_t (&v)[32]) {
    float r[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 b = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        r[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + b.x;
        r[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b.y;
        r[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b.z;
        r[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b.w;
    }
    if (p.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = fmaxf(r[j], 0.f);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + col0) + j);
        const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + col0) + j);
        r[4 * j + 0] = fmaf(r[4 * j + 0], sc.x, sh.x);
        r[4 * j + 1] = fmaf(r[4 * j + 1], sc.y, sh.y);
        r[4 * j + 2] = fmaf(r[4 * j + 2], sc.z, sh.z);
        r[4 * j + 3] = fmaf(r[4 * j + 3], sc.w, sh.w);
    }
    uint4* o4 = reinterpret_cast<uint4*>(p.out_bf16 + out_row * p.ldo_bf16 + col0);
    uint32_t hi[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) hi[j] = pack_bf16x2(r[2 * j], r[2 * j + 1]);
#pragma unroll
    for (int j = 0; j < 4; ++j) o4[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
    if (p.split_out) {  // second bf16 plane: what the first one lost (v = hi + lo to ~2^-17)
        uint32_t lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
            lo[j] = pack_bf16x2(r[2 * j] - __uint_as_float(hi[j] << 16), r[2 * j + 1] - __uint_as_float(hi[j] & 0xffff0000u));
        uint4* l4 = o4 + (p.c_store >> 3);
#pragma unroll
        for (int j = 0; j < 4; ++j) l4[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
    }
}

// DDPM epilogue: row = clip*T + frame, columns = pose channels.  Lanes of a warp hold consecutive frames, so for a fixed
// channel the (N,C,T) accesses of a warp are contiguous.  x, the noise slab of this step and the optional eps/x0
// outputs share one 32-bit element offset per column (off0 + j*T); AUX / INPAINT are compile-time so the common
// sampling step carries no dead predicates (the first version spent ~94 instructions per element, mostly on 64-bit
// index arithmetic and per-element pointer tests).
// x and the step's noise of one 32-column chunk, issued BEFORE the warp waits for the accumulator (they do not depend on it):
// the two HBM round trips of the update then overlap the tile's TMA loads and MMAs instead of following them.
__device__ __forceinline__ void ddpm_prefetch_chunk(const GemmParams& p, const float* tape_t, int row, int col0, float (&xv)[32],
                                                    float (&zv)[32]) {
    const gd_ddpm_desc& u = p.ddpm;
    const int T = u.T;
    const int clip = row / T;
    const uint32_t off0 = static_cast<uint32_t>((clip * u.C + col0) * T + (row - clip * T));
    const int ncol = u.C - col0;
    const float* const xp = u.x + off0;
    const float* const zp = tape_t ? tape_t + off0 : nullptr;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const bool ok = j < ncol;
        xv[j] = ok ? __ldcg(xp + j * T) : 0.f;
        zv[j] = (ok && zp) ? __ldg(zp + j * T) : 0.f;
    }
}

template <bool AUX, bool INPAINT, bool PRE = false>
__device__ __forceinline__ void epilogue_ddpm_chunk(const GemmParams& p, const DdpmStepCoefs& cf, const float* tape_t,
                                                    int row, int col0, const uint32_t (&v)[32], const float* xpre = nullptr,
                                                    const float* zpre = nullptr) {
    const gd_ddpm_desc& u = p.ddpm;
    const int T = u.T;
    const int clip = row / T;
    const int frame = row - clip * T;
    const uint32_t off0 = static_cast<uint32_t>((clip * u.C + col0) * T + frame);
    const int ncol = u.C - col0;  // columns >= ncol of this chunk are padding (warp-uniform)
    float* const xp = u.x + off0;
    const float* const zp = tape_t ? tape_t + off0 : nullptr;
    const float* const xa_add = u.xa_add ? u.xa_add + off0 : nullptr;  // Inpaint model: loop-invariant input offset
    float* const eps_o = (AUX && u.eps_out) ? u.eps_out + off0 : nullptr;
    float* const x0_o = (AUX && u.x0_out) ? u.x0_out + off0 : nullptr;
    float* const mean_o = (AUX && u.mean_out) ? u.mean_out + off0 : nullptr;
    float* const raw_o = (AUX && u.raw_x0_out) ? u.raw_x0_out + off0 : nullptr;
    float m = 0.f, f = 0.f;
    const float* sp = nullptr;
    if (INPAINT) {
        m = __ldg(u.inpaint_mask + clip * T + frame);
        f = __ldg(u.inpaint_factor + frame);
        sp = u.inpaint_seed + (static_cast<size_t>(clip) * T + frame) * u.C + col0;
    }
    // per 16-column half: all loads first (x aliases the stores below, so the compiler cannot hoist them itself)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float xv[16], zv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int j = h * 16 + i;
            const bool ok = j < ncol;
            if (PRE) {
                xv[i] = xpre[j];
                zv[i] = zpre[j];
            } else {
                xv[i] = ok ? __ldcg(xp + j * T) : 0.f;
                zv[i] = (ok && zp) ? __ldg(zp + j * T) : 0.f;
            }
        }
        float xn[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int j = h * 16 + i;
            xn[i] = 0.0f;
            if (j < ncol) {
                const float eps = __uint_as_float(v[j]) + __ldg(p.bias + col0 + j);
                const float sv = INPAINT ? __ldg(sp + j) : 0.f;
                float x0, mean, raw;
                const float xnext = ddpm_update_elem(cf, xv[i], eps, zv[i], INPAINT, sv, m, f, u.clip_x0, &x0, &mean, &raw);
                xp[j * T] = xnext;
                if (AUX) {
                    if (eps_o) eps_o[j * T] = eps;
                    if (x0_o) x0_o[j * T] = x0;
                    if (mean_o) mean_o[j * T] = mean;
                    if (raw_o) raw_o[j * T] = raw;
                }
                xn[i] = xa_add ? xnext + __ldg(xa_add + j * T) : xnext;
            }
        }
        if (u.xa_bf16) {
            uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(u.xa_bf16) + (size_t)row * u.ld_xa + col0 + h * 16);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint4 w;
                w.x = pack_bf16x2(xn[8 * j + 0], xn[8 * j + 1]);
                w.y = pack_bf16x2(xn[8 * j + 2], xn[8 * j + 3]);
                w.z = pack_bf16x2(xn[8 * j + 4], xn[8 * j + 5]);
                w.w = pack_bf16x2(xn[8 * j + 6], xn[8 * j + 7]);
                o4[j] = w;
            }
        }
    }
}

// CL = 1: one CTA per 128 x BN tile.  CL = 2: a CTA pair (cta_group::2) owns a 256 x BN tile - each CTA stages its own
// 128 rows of A and HALF of the W tile, the leader CTA issues one M=256 MMA that reads both halves, and every CTA
// keeps the accumulator rows of its own m-tile in its own TMEM.  The big GEMMs are bound by what an SM can ingest
// from L2 (~42 B/clk measured: 48 KB per k-block vs 512 MMA clocks); pairing cuts that to 32 KB per k-block.
template <int BN, int MODE, int CL>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out, const GemmParams p) {
    constexpr bool IS_CONV = MODE >= MODE_CONV;
    constexpr bool PREFETCH_W = (CL == 1) && !IS_CONV;  // single-CTA tiles of the Linear layers
    constexpr int NA = (MODE == MODE_CONV_S2) ? 2 : 1, NB = (MODE >= MODE_CONV_S1) ? 2 : 1;
    using Cfg = GemmCfg<BN, CL, NA, NB, (MODE == MODE_TMA_F32) ? GD_F32_STAGING : 4096>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * NA * Cfg::A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);  // 256 B reserved
    uint64_t* full_bar = bars;                    // [STAGES]  TMA -> MMA
    uint64_t* empty_bar = bars + STAGES;          // [STAGES]  MMA -> TMA
    uint64_t* acc_full_bar = bars + 2 * STAGES;   // [2]       MMA -> epilogue
    uint64_t* acc_empty_bar = acc_full_bar + 2;   // [2]       epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty_bar + 2);
    uint8_t* staging = smem + STAGES * Cfg::STAGE_BYTES + 1024;  // 1024-B aligned: TMA swizzle atoms

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int cta_rank = (CL > 1) ? static_cast<int>(cluster_ctarank()) : 0;
    const int cluster_id = blockIdx.x / CL, num_clusters = gridDim.x / CL;
    const int m_groups = ((p.M + BLOCK_M - 1) / BLOCK_M + CL - 1) / CL;  // groups of CL consecutive m-tiles
    const int n_tiles = p.N / BN;
    const int num_tiles = m_groups * n_tiles;  // work items per cluster-wide scheduler
    const int k_blocks = p.K / (BLOCK_K * NB);  // pipeline stages per tile (a MODE_CONV_S* stage covers NB k-blocks of W)
    // work item `tile` -> this CTA's output tile; an m-tile past the end (odd tile count) is all out-of-bounds:
    // TMA zero-fills its loads and clips its stores, so the CTA just keeps the cluster protocol going
#define GD_TILE_M0(tile) ((((tile) / n_tiles) * CL + cta_rank) * BLOCK_M)

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tmap_a);
        prefetch_tensormap(&tmap_b);
        if (MODE == MODE_TMA_BF16 || MODE == MODE_TMA_F32) prefetch_tensormap(&tmap_out);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);   // CL = 2: only the leader's is used (both CTAs' TMA bytes land on it)
            mbar_init(&empty_bar[s], 1);  // released by the (pair-wide) MMA commit
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full_bar[s], 1);
            mbar_init(&acc_empty_bar[s], CL * EPI_WARPS);  // one arrive per epilogue warp of every CTA of the pair
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        if (CL == 1)
            tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
        else
            tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot);
    }
#ifdef GD_TRACE
    __shared__ int trace_slot_s;
    if (threadIdx.x == 0) {
        trace_slot_s = GD_TRACE_OPEN(100 + MODE);
        GD_TRACE_MARK(trace_slot_s, 0);  // entry (before barrier init / TMEM allocation complete)
    }
#endif
    tc_fence_before_sync();
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // the peer's barriers and TMEM exist before any cross-CTA traffic
    tc_fence_after_sync();
#ifdef GD_TRACE
    const int trace_slot = trace_slot_s;
    if (threadIdx.x == 0) GD_TRACE_MARK(trace_slot, 1);  // prologue done
#else
    const int trace_slot = -1;
    (void)trace_slot;
#endif
    pdl_launch_dependents();
    // W is a weight matrix: it does not depend on the previous kernel of the chain, so the producer puts the W tiles of its first
    // pipeline stages in flight BEFORE it waits for that kernel (the CTA is resident and idle for 2-7 us at small batches,
    // profiles/r02_kernel_timeline_*.json); after the wait only the activation tiles remain to be fetched for those stages.
    int w_preissued = 0;
    if (PREFETCH_W && warp == 0 && lane == 0) {
        for (int tile = cluster_id; tile < num_tiles && w_preissued < STAGES; tile += num_clusters) {
            const int n0 = (tile % n_tiles) * BN;
            for (int kb = 0; kb < k_blocks && w_preissued < STAGES; ++kb, ++w_preissued) {
                mbar_arrive_expect_tx(&full_bar[w_preissued], Cfg::STAGE_BYTES);
                tma_load_2d(smem_b + w_preissued * Cfg::B_BYTES, &tmap_b, &full_bar[w_preissued], kb * BLOCK_K, n0);
            }
        }
    }
    pdl_wait();  // activations and outputs belong to the chain: nothing below may touch them before the previous kernel is done
    if (threadIdx.x == 0) GD_TRACE_MARK(trace_slot, 2);  // predecessor complete
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int issued = 0;  // pipeline stages filled so far; the first `w_preissued` already have their W tile in flight
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                const int m0 = GD_TILE_M0(tile);
                const int n0 = (tile % n_tiles) * BN;
                for (int kb = 0; kb < k_blocks; ++kb, ++issued) {
                    if (PREFETCH_W && issued < w_preissued) {  // first pass over the ring: slot free, W on its way, A missing
                        tma_load_2d(smem_a + stage * NA * Cfg::A_BYTES, &tmap_a, &full_bar[stage], kb * BLOCK_K, m0);
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                        continue;
                    }
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (CL == 1) {
                        mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                        int a_col = kb * BLOCK_K, a_row = m0;
                        if (IS_CONV) {  // tap-shifted pixel rows; rows outside [0, M) are zero-filled by TMA
                            const int tap = kb / p.kb_per_tap;
                            a_col = (kb - tap * p.kb_per_tap) * BLOCK_K;
                            if (MODE == MODE_CONV) a_col %= p.a_cols;  // wraps: split rows are walked hi, lo, hi
                            a_row = m0 + p.tap_shift[tap];
                        }
                        tma_load_2d(smem_a + stage * NA * Cfg::A_BYTES, &tmap_a, &full_bar[stage], a_col, a_row);
                        if (NA == 2)  // the lo plane of the same 64 channels
                            tma_load_2d(smem_a + (stage * NA + 1) * Cfg::A_BYTES, &tmap_a, &full_bar[stage], a_col + (p.a_cols >> 1), a_row);
#pragma unroll
                        for (int b = 0; b < NB; ++b)
                            tma_load_2d(smem_b + (stage * NB + b) * Cfg::B_BYTES, &tmap_b, &full_bar[stage], (kb * NB + b) * BLOCK_K, n0);
                    } else {
                        // both CTAs fill their own slot; all bytes are counted on the leader's barrier
                        if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], CL * Cfg::STAGE_BYTES);
                        const uint32_t lead_bar = map_to_cta(smem_u32(&full_bar[stage]), 0);
                        tma_load_2d_pair(smem_a + stage * Cfg::A_BYTES, &tmap_a, lead_bar, kb * BLOCK_K, m0);
                        tma_load_2d_pair(smem_b + stage * Cfg::B_BYTES, &tmap_b, lead_bar, kb * BLOCK_K,
                                         n0 + cta_rank * (BN / CL));
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && cta_rank == 0) {  // the leader CTA issues the MMAs of the whole pair
            constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M * CL, BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&acc_empty_bar[acc], acc_phase ^ 1);
                tc_fence_after_sync();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after_sync();
                    if (it == 0 && kb == 0) GD_TRACE_MARK(trace_slot, 3);  // first operands landed
                    if constexpr (NB == 2) {
                        // operand reuse: every (A, B) product of the stage from tiles that were loaded once
                        constexpr int NPAIR = NA + 1;  // S1: (A,B0) (A,B1);  S2: (hi,Whi) (lo,Whi) (hi,Wlo)
#pragma unroll
                        for (int pr = 0; pr < NPAIR; ++pr) {
                            const int ai = (NA == 2 && pr == 1) ? 1 : 0, bi = (pr == NPAIR - 1) ? 1 : 0;
                            const uint64_t da = umma_desc_k_sw128(smem_u32(smem_a + (stage * NA + ai) * Cfg::A_BYTES));
                            const uint64_t db = umma_desc_k_sw128(smem_u32(smem_b + (stage * NB + bi) * Cfg::B_BYTES));
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                                umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | pr | k) != 0 ? 1u : 0u);
                        }
                    } else {
                    const uint64_t da = umma_desc_k_sw128(smem_u32(smem_a + stage * Cfg::A_BYTES));
                    const uint64_t db = umma_desc_k_sw128(smem_u32(smem_b + stage * Cfg::B_BYTES));
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // +32 B per UMMA_K step inside the 128-B swizzle row: start-address field += 2
                        if (CL == 1)
                            umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        else
                            umma_bf16_ss_pair(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    }
                    // frees the smem slot (of both CTAs of a pair) when these MMAs retire
                    if (CL == 1)
                        umma_commit(&empty_bar[stage]);
                    else
                        umma_commit_pair(&empty_bar[stage], 0b11);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (CL == 1)  // accumulator complete -> epilogue (of both CTAs of a pair)
                    umma_commit(&acc_full_bar[acc]);
                else
                    umma_commit_pair(&acc_full_bar[acc], 0b11);
                GD_TRACE_MARK(trace_slot, 4);  // MMAs of the (last) tile issued
            }
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;
        const int quad = ew & 3;    // TMEM lanes [32*quad, 32*quad+32)  (a warp may only touch quadrant warp%4)
        const int half = ew >> 2;   // which half of the tile's columns
        constexpr int WCOLS = BN / 2;
        int t = 0;
        DdpmStepCoefs cf;
        const float* tape_t = nullptr;
        bool ddpm_aux = false, ddpm_inpaint = false;
        if (MODE == MODE_DDPM) {
            t = load_step(p.ddpm.step_ptr);
            cf = ddpm_load_coefs(p.ddpm, t);
            if (p.ddpm.noise_tape) tape_t = p.ddpm.noise_tape + (size_t)t * p.ddpm.n_clips * p.ddpm.C * p.ddpm.T;
            ddpm_aux = p.ddpm.eps_out != nullptr || p.ddpm.x0_out != nullptr || p.ddpm.mean_out != nullptr || p.ddpm.raw_x0_out != nullptr;
            if (ddpm_aux && p.ddpm.aux_step_ptr) {  // optional outputs only in the step that asks for them
                const int aux_t = load_step(p.ddpm.aux_step_ptr);
                ddpm_aux = aux_t < 0 || aux_t == t;
            }
            ddpm_inpaint = p.ddpm.inpaint_seed != nullptr;
        }
        uint8_t* stg_base = staging + ew * Cfg::STAGING_PER_WARP;
        constexpr int STG_CHUNK = (MODE == MODE_TMA_F32) ? 4096 : 2048;
        constexpr int STG_RING = Cfg::STAGING_PER_WARP / STG_CHUNK;
        uint32_t chunk_ctr = 0;  // chunks this warp has handed to the TMA engine
        const bool has_bias = p.bias != nullptr;
        int it = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
            const int m0 = GD_TILE_M0(tile);
            const int n0 = (tile % n_tiles) * BN;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int row0 = m0 + quad * 32;
            const int row = row0 + lane;
            constexpr bool DDPM_PRE = (MODE == MODE_DDPM && BN == 64);  // one 32-column chunk per warp and tile
            float xpre[DDPM_PRE ? 32 : 1], zpre[DDPM_PRE ? 32 : 1];
            if constexpr (DDPM_PRE) {
                if (row < p.M && !ddpm_aux && !ddpm_inpaint) ddpm_prefetch_chunk(p, tape_t, row, n0 + half * (BN / 2), xpre, zpre);
            }
            mbar_wait(&acc_full_bar[acc], acc_phase);
            tc_fence_after_sync();
            if (ew == 0 && lane == 0) GD_TRACE_MARK(trace_slot, 5);  // accumulator of the (last) tile complete
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + half * WCOLS;
            // the accumulator stage goes back to the MMA issuer as soon as this warp's last TMEM load has landed in
            // registers - not after the stores - so tile i+2 can start while tile i is still being written out
            auto release_accumulator = [&]() {
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) {
                    if (CL == 1)
                        mbar_arrive(&acc_empty_bar[acc]);
                    else  // the MMA issuer lives in the leader CTA
                        mbar_arrive_cluster(map_to_cta(smem_u32(&acc_empty_bar[acc]), 0));
                }
            };
            bool conv_keep = false;
            size_t conv_out_row = 0;
            if (IS_CONV) {  // pixel row -> (image, y, x); only pixels inside the window (and on the stride) are stored
                const int gsz = p.grid_h * p.grid_w;
                const int img = row / gsz;
                const int rem = row - img * gsz;
                const int y = rem / p.grid_w, x = rem - y * p.grid_w;
                int dy = y - p.y0, dx = x - p.x0;
                conv_keep = row < p.M && y >= p.y0 && y <= p.y1 && x >= p.x0 && x <= p.x1;
                if (p.stride == 2) {
                    conv_keep = conv_keep && ((dy | dx) & 1) == 0;
                    dy >>= 1, dx >>= 1;
                }
                conv_out_row = (size_t)img * p.out_img_stride + (size_t)(dy * p.out_y_stride + dx * p.out_x_stride + p.out_offset);
            }
            if constexpr (MODE == MODE_TMA_BF16 && WCOLS >= 64) {
                // bf16 output, 64 columns (one 128-B swizzled row per lane) per TMA store: half as many fences / stores /
                // staging hand-offs as 32-column chunks, and the next TMEM load is in flight while this one is stored
                constexpr int NSC = WCOLS / 64;
                uint32_t v[64];
                tmem_ld_32x32b_x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
                tmem_ld_32x32b_x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
#pragma unroll 1
                for (int sc = 0; sc < NSC; ++sc) {
                    const int col0 = n0 + half * WCOLS + sc * 64;
                    tmem_ld_wait();
                    uint32_t w[32];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float4 b[8];
                        if (has_bias) {
                            const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0 + h * 32);
#pragma unroll
                            for (int j = 0; j < 8; ++j) b[j] = __ldg(b4 + j);
                        }
                        float r[32];
                        add_bias_chunk(r, *reinterpret_cast<const uint32_t(*)[32]>(&v[32 * h]), b, has_bias);
                        if (p.rowbias) add_rowbias_chunk(r, p, row, col0 + h * 32);
                        if (p.act == GD_ACT_RELU2) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const float m = fmaxf(r[j], 0.f);
                                r[j] = m * m;
                            }
                        } else if (p.act == GD_ACT_SILU) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) r[j] = apply_act(r[j], GD_ACT_SILU);
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) w[16 * h + j] = pack_bf16x2(r[2 * j], r[2 * j + 1]);
                    }
                    if (sc + 1 < NSC) {
                        tmem_ld_32x32b_x32(taddr + (sc + 1) * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
                        tmem_ld_32x32b_x32(taddr + (sc + 1) * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
                    } else {
                        release_accumulator();
                    }
                    if (lane == 0) bulk_wait_group_read0();  // the previous store has finished READING the staging tile
                    __syncwarp();
                    uint4* st4 = reinterpret_cast<uint4*>(stg_base);  // 128-B rows, SWIZZLE_128B: chunk ^= row & 7
#pragma unroll
                    for (int j = 0; j < 8; ++j) st4[lane * 8 + (j ^ (lane & 7))] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
                    fence_proxy_async();  // make the generic-proxy smem writes visible to the TMA engine
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmap_out, stg_base, col0, row0);
                        bulk_commit_group();
                    }
                }
            } else {
#pragma unroll 1
            for (int c = 0; c < WCOLS / 32; ++c) {
                const int col0 = n0 + half * WCOLS + c * 32;
                uint32_t v[32];
                tmem_ld_32x32b_x32(taddr + c * 32, v);
                if (MODE == MODE_TMA_BF16 || MODE == MODE_TMA_F32) {
                    // bias loads are issued while the TMEM load is in flight
                    float4 b[8];
                    if (has_bias) {
                        const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) b[j] = __ldg(b4 + j);
                    }
                    tmem_ld_wait();
                    if (c == WCOLS / 32 - 1) release_accumulator();
                    float r[32];
                    add_bias_chunk(r, v, b, has_bias);
                    if (p.rowbias) add_rowbias_chunk(r, p, row, col0);
                    if (p.act == GD_ACT_RELU2) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float m = fmaxf(r[j], 0.f);
                            r[j] = m * m;
                        }
                    } else if (p.act == GD_ACT_SILU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) r[j] = apply_act(r[j], GD_ACT_SILU);
                    }
                    // ring slot: the store issued STG_RING chunks ago must have finished READING it
                    uint8_t* stg = stg_base + (chunk_ctr % STG_RING) * STG_CHUNK;
                    ++chunk_ctr;
                    if (lane == 0) bulk_wait_group_read<STG_RING - 1>();
                    __syncwarp();
                    if (MODE == MODE_TMA_F32) {
                        float4* st4 = reinterpret_cast<float4*>(stg);  // 128-B rows, SWIZZLE_128B: chunk ^= row & 7
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            st4[lane * 8 + (j ^ (lane & 7))] = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                    } else {
                        uint4* st4 = reinterpret_cast<uint4*>(stg);  // 64-B rows, SWIZZLE_64B: chunk ^= (row >> 1) & 3
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 w;
                            w.x = pack_bf16x2(r[8 * j + 0], r[8 * j + 1]);
                            w.y = pack_bf16x2(r[8 * j + 2], r[8 * j + 3]);
                            w.z = pack_bf16x2(r[8 * j + 4], r[8 * j + 5]);
                            w.w = pack_bf16x2(r[8 * j + 6], r[8 * j + 7]);
                            st4[lane * 4 + (j ^ ((lane >> 1) & 3))] = w;
                        }
                    }
                    fence_proxy_async();  // make the generic-proxy smem writes visible to the TMA engine
                    __syncwarp();
                    if (lane == 0) {
                        if (MODE == MODE_TMA_F32 && p.reduce_add)
                            tma_reduce_add_2d(&tmap_out, stg, col0, row0);
                        else
                            tma_store_2d(&tmap_out, stg, col0, row0);
                        bulk_commit_group();
                    }
                } else if (IS_CONV) {
                    tmem_ld_wait();
                    if (c == WCOLS / 32 - 1) release_accumulator();
                    if (conv_keep && col0 < p.c_store) epilogue_conv_chunk(p, conv_out_row, col0, v);
                } else {
                    tmem_ld_wait();
                    if (row < p.M) {
                        if (MODE == MODE_DDPM) {
                            if (ddpm_inpaint)
                                epilogue_ddpm_chunk<true, true>(p, cf, tape_t, row, col0, v);
                            else if (ddpm_aux)
                                epilogue_ddpm_chunk<true, false>(p, cf, tape_t, row, col0, v);
                            else if constexpr (DDPM_PRE)
                                epilogue_ddpm_chunk<false, false, true>(p, cf, tape_t, row, col0, v, xpre, zpre);
                            else
                                epilogue_ddpm_chunk<false, false>(p, cf, tape_t, row, col0, v);
                        } else
                            epilogue_direct_chunk(p, row, col0, v);
                    }
                    if (c == WCOLS / 32 - 1) release_accumulator();
                }
            }
            }
        }
        if (ew == 0 && lane == 0) GD_TRACE_MARK(trace_slot, 6);  // last store issued
        if ((MODE == MODE_TMA_BF16 || MODE == MODE_TMA_F32) && lane == 0) bulk_wait_group0();
        if (ew == 0 && lane == 0) GD_TRACE_MARK(trace_slot, 7);  // stores drained
    }

    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) GD_TRACE_MARK(trace_slot, 8);  // exit
    if (CL > 1) cluster_sync_all();  // no CTA leaves while a peer can still multicast into it / arrive on its barriers
    if (warp == 2) {
        if (CL == 1)
            tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
        else
            tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
    }
#undef GD_TILE_M0
}

// ------------------------------------------------------------------------------------------ host
int make_tmap_2d(CUtensorMap* m, CUtensorMapDataType dt, int elt_bytes, const void* base, uint64_t rows, uint64_t cols,
                 uint64_t ld_elems, uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle swz) {
    static PFN_encodeTiled encode = get_encode_tiled();
    if (!encode) return set_error(GD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld_elems * elt_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(m, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(GD_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return GD_OK;
}
static int make_tmap_2d_bf16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                             uint32_t box_rows) {
    return make_tmap_2d(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, rows, cols, ld_elems, BLOCK_K, box_rows,
                        CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int BN, int MODE, int CL>
static int launch_gemm_cl(const GemmParams& p, const void* A, int lda, const void* W, int ldw, cudaStream_t stream) {
    using Cfg = GemmCfg<BN, CL, 1, 1, (MODE == MODE_TMA_F32) ? GD_F32_STAGING : 4096>;
    CUtensorMap ta, tb, tout;
    int rc = make_tmap_2d_bf16(&ta, A, p.M, p.K, lda, BLOCK_M);
    if (rc) return rc;
    rc = make_tmap_2d_bf16(&tb, W, p.N, p.K, ldw, BN / CL);
    if (rc) return rc;
    tout = ta;  // unused unless a TMA epilogue is selected
    if (MODE == MODE_TMA_BF16)  // 64-column (128-B) boxes whenever an epilogue warp owns at least 64 columns of the tile
        rc = make_tmap_2d(&tout, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.out_bf16, p.M, p.N, p.ldo_bf16, BN / 2 >= 64 ? 64 : 32, 32,
                          BN / 2 >= 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
    else if (MODE == MODE_TMA_F32)
        rc = make_tmap_2d(&tout, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p.out_f32, p.M, p.N, p.ldo_f32, 32, 32,
                          CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    static bool attr_set[GD_MAX_DEVICES] = {};  // function attributes are per device
    const int dev_idx = current_device();
    if (!attr_set[dev_idx]) {
        GD_CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, MODE, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Cfg::SMEM_BYTES));
        attr_set[dev_idx] = true;
    }
    const int m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
    const int work = ((m_tiles + CL - 1) / CL) * (p.N / BN);  // cluster-level work items
    const int sms = (p.max_ctas > 0 && p.max_ctas < sm_count()) ? p.max_ctas : sm_count();
    const int max_clusters = sms / CL > 0 ? sms / CL : 1;
    const int grid = (work < max_clusters ? work : max_clusters) * CL;
    GD_CUDA_CHECK(launch_k(gemm_bf16_tn_kernel<BN, MODE, CL>, grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream, CL, ta, tb, tout, p));
    count_launch();
    GD_CUDA_CHECK(cudaGetLastError());
    return GD_OK;
}

// GD_GEMM_PAIR: 0 = never pair, 2 = pair whenever the shape allows (tests, A/B runs); default = long K only.
static int cta_pair_policy() {
    const char* e = getenv("GD_GEMM_PAIR");
    return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
}

// CTA pairs (256 x BN tiles).  Measured on B200 (profiles/r01_kernel_bench_*.jsonl): +7 % on the K = 2048 down-projections
// (69 vs 74 us), nothing at K = 512 (the tile is then bound by its epilogue/output traffic, as cuBLAS is: 1.07 PFLOP/s
// there for both), slower at K = 256 - so pairs are used for long K only.
template <int BN, int MODE>
static int launch_gemm(const GemmParams& p, const void* A, int lda, const void* W, int ldw, cudaStream_t stream) {
    const int m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
    const int policy = cta_pair_policy();
    if (BN >= 128 && MODE != MODE_DDPM && MODE != MODE_DIRECT && MODE < MODE_CONV && policy > 0 && (p.K >= 1024 || policy == 2) &&
        ((m_tiles + 1) / 2) * (p.N / BN) >= sm_count() / 2)
        return launch_gemm_cl<BN, MODE, 2>(p, A, lda, W, ldw, stream);
    return launch_gemm_cl<BN, MODE, 1>(p, A, lda, W, ldw, stream);
}

// MODE_CONV: the A operand is the pixel-row tensor [rows, c_in]; every tap reads the same columns at shifted rows, and
// W is [c_out, n_taps * c_in] with the taps along K.
template <int BN, int MODE>
static int launch_conv(const GemmParams& p, const void* A, int c_in, const void* W, cudaStream_t stream) {
    using Cfg = GemmCfg<BN, 1, (MODE == MODE_CONV_S2) ? 2 : 1, (MODE >= MODE_CONV_S1) ? 2 : 1>;
    CUtensorMap ta, tb;
    int rc = make_tmap_2d_bf16(&ta, A, p.M, c_in, c_in, BLOCK_M);
    if (rc) return rc;
    rc = make_tmap_2d_bf16(&tb, W, p.N, p.K, p.K, BN);
    if (rc) return rc;
    static bool attr_set[GD_MAX_DEVICES] = {};  // function attributes are per device
    const int dev_idx = current_device();
    if (!attr_set[dev_idx]) {
        GD_CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, MODE, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Cfg::SMEM_BYTES));
        attr_set[dev_idx] = true;
    }
    const int work = ((p.M + BLOCK_M - 1) / BLOCK_M) * (p.N / BN);
    const int grid = work < sm_count() ? work : sm_count();
    GD_CUDA_CHECK(launch_k(gemm_bf16_tn_kernel<BN, MODE, 1>, grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream, 1, ta, tb, ta, p));
    count_launch();
    GD_CUDA_CHECK(cudaGetLastError());
    return GD_OK;
}

static int validate_linear(const gd_linear_desc* d) {
    if (!d || !d->A || !d->W) return set_error(GD_ERR_INVALID, "gd_linear: null descriptor/operand");
    if (d->M <= 0 || d->N <= 0 || d->K <= 0) return set_error(GD_ERR_INVALID, "gd_linear: non-positive shape");
    if (d->K % BLOCK_K) return set_error(GD_ERR_INVALID, "gd_linear: K=%d must be a multiple of 64", d->K);
    if (d->N % 64) return set_error(GD_ERR_INVALID, "gd_linear: N=%d must be a multiple of 64", d->N);
    if (d->lda % 8 || d->ldw % 8 || d->lda < d->K || d->ldw < d->K)
        return set_error(GD_ERR_INVALID, "gd_linear: lda/ldw must be >= K and multiples of 8");
    if ((reinterpret_cast<uintptr_t>(d->A) | reinterpret_cast<uintptr_t>(d->W)) & 15)
        return set_error(GD_ERR_INVALID, "gd_linear: A/W must be 16-byte aligned");
    return GD_OK;
}

#ifdef GD_TRACE
void set_trace_gemm(unsigned long long* buf) { cudaMemcpyToSymbol(t_trace_buf, &buf, sizeof(buf)); }
#endif

}  // namespace gd

using namespace gd;

extern "C" int gd_linear_bf16(const gd_linear_desc* d, void* stream) {
    gd::KindScope kind_scope("gemm");
    int rc = validate_linear(d);
    if (rc) return rc;
    if (!d->out_f32 && !d->out_bf16) return set_error(GD_ERR_INVALID, "gd_linear_bf16: no output buffer");
    if ((d->out_f32 && d->ldo_f32 % 4) || (d->out_bf16 && d->ldo_bf16 % 8) || (d->residual && d->ldr % 4))
        return set_error(GD_ERR_INVALID, "gd_linear_bf16: output/residual row strides must keep 16-byte alignment");
    if (d->rowbias && d->rowbias_period <= 0) return set_error(GD_ERR_INVALID, "gd_linear_bf16: rowbias_period <= 0");
    rc = check_device();
    if (rc) return rc;
    GemmParams p{};
    p.M = d->M, p.N = d->N, p.K = d->K;
    p.bias = d->bias, p.rowbias = d->rowbias;
    p.rowbias_period = d->rowbias_period, p.rowbias_offset = d->rowbias_offset;
    p.residual = d->residual, p.ldr = d->ldr, p.act = d->act;
    p.out_f32 = d->out_f32, p.ldo_f32 = d->ldo_f32;
    p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(d->out_bf16), p.ldo_bf16 = d->ldo_bf16;
    p.max_ctas = d->max_ctas;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int m_tiles = (d->M + BLOCK_M - 1) / BLOCK_M;
    // epilogue flavour: TMA stores for the two hot shapes, the per-row path for everything else
    int mode = MODE_DIRECT;
    const bool aligned16 = ((reinterpret_cast<uintptr_t>(d->out_f32) | reinterpret_cast<uintptr_t>(d->out_bf16)) & 15) == 0;
    if (aligned16) {  // (the positional rowbias is supported by the TMA epilogues too)
        if (d->out_bf16 && !d->out_f32 && !d->residual)
            mode = MODE_TMA_BF16;
        else if (d->out_f32 && !d->out_bf16 && (!d->residual || (d->residual == d->out_f32 && d->ldr == d->ldo_f32))) {
            mode = MODE_TMA_F32;
            p.reduce_add = d->residual != nullptr;
            p.residual = nullptr;
        }
    }
    // Widest tile that still gives every SM work; narrow tiles when the problem is small.
    int bn = (d->N % 256 == 0 && m_tiles * (d->N / 256) >= sm_count()) ? 256
             : (d->N % 128 == 0 && m_tiles * (d->N / 128) >= sm_count()) ? 128 : 64;
    {
        // Problems of up to two waves of m-tiles are bound by the operand bytes the busiest CTA has to pull from L2, not by
        // throughput (profiles/r02_kernel_timeline_beat128.json: ~100 GB/s per SM): pick the tile width that minimises
        // rounds x (A tile + W tile) bytes; ties go to the wider tile.  Measured (r02_ab_small_problem_tile_rule*.jsonl): beat 128
        // clips -4.7 %, 256 clips -2.4 %, beat-4x 64 clips -1.4 %, tedexp 32 clips -2.5 %, larger batches unchanged (the rule
        // then picks what the old one did).  GD_GEMM_SMALL=0: old rule; GD_GEMM_SMALL_M: m-tile limit.
        static const bool small_rule = !(getenv("GD_GEMM_SMALL") && getenv("GD_GEMM_SMALL")[0] == '0');
        static const int small_max = getenv("GD_GEMM_SMALL_M") ? atoi(getenv("GD_GEMM_SMALL_M")) : sm_count() * 2;
        if (small_rule && m_tiles <= small_max) {
            long best = -1;
            for (int cand = 256; cand >= 64; cand >>= 1) {
                if (d->N % cand) continue;
                const long tiles = (long)m_tiles * (d->N / cand);
                const long rounds = (tiles + sm_count() - 1) / sm_count();
                const long bytes = rounds * ((long)d->K * (BLOCK_M + cand) * 2);
                if (best < 0 || bytes < best) best = bytes, bn = cand;
            }
        }
    }
    {  // GD_GEMM_BN=128|64 caps the tile width (A/B runs: wave quantisation against per-tile efficiency)
        static const int bn_cap = getenv("GD_GEMM_BN") ? atoi(getenv("GD_GEMM_BN")) : 0;
        if (bn_cap >= 64 && bn > bn_cap) bn = bn_cap;
    }
#define GD_LAUNCH(BN_)                                                                                   \
    (mode == MODE_TMA_BF16  ? launch_gemm<BN_, MODE_TMA_BF16>(p, d->A, d->lda, d->W, d->ldw, s)          \
     : mode == MODE_TMA_F32 ? launch_gemm<BN_, MODE_TMA_F32>(p, d->A, d->lda, d->W, d->ldw, s)           \
                            : launch_gemm<BN_, MODE_DIRECT>(p, d->A, d->lda, d->W, d->ldw, s))
    if (bn == 256) return GD_LAUNCH(256);
    if (bn == 128) return GD_LAUNCH(128);
    return GD_LAUNCH(64);
#undef GD_LAUNCH
}

extern "C" int gd_linear_ddpm(const gd_linear_desc* d, const gd_ddpm_desc* u, void* stream) {
    gd::KindScope kind_scope("ddpm");
    int rc = validate_linear(d);
    if (rc) return rc;
    rc = validate_ddpm(u);
    if (rc) return rc;
    if (d->N != 128 || u->C > 128) return set_error(GD_ERR_INVALID, "gd_linear_ddpm: N must be 128 (padded d_pose)");
    if (d->M != u->n_clips * u->T) return set_error(GD_ERR_INVALID, "gd_linear_ddpm: M != n_clips*T");
    if (!d->bias) return set_error(GD_ERR_INVALID, "gd_linear_ddpm: bias required");
    if ((int64_t)u->n_clips * u->C * u->T >= (int64_t)1 << 31)
        return set_error(GD_ERR_INVALID, "gd_linear_ddpm: n_clips*C*T must stay below 2^31 elements");
    if (u->xa_bf16 && (u->ld_xa % 8 || u->ld_xa < 128))
        return set_error(GD_ERR_INVALID, "gd_linear_ddpm: ld_xa must be >= 128 and a multiple of 8");
    rc = check_device();
    if (rc) return rc;
    GemmParams p{};
    p.M = d->M, p.N = d->N, p.K = d->K;
    p.bias = d->bias;
    p.ddpm = *u;
    // 128 x 64 tiles: twice as many CTAs share the (epilogue-bound) update - the whole pose matrix is only 68..320 m-tiles
    return launch_gemm<64, MODE_DDPM>(p, d->A, d->lda, d->W, d->ldw, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int gd_conv_taps_bf16(const gd_conv_desc* d, void* stream) {
    gd::KindScope kind_scope("speech");
    if (!d || !d->in || !d->W || !d->out || !d->scale || !d->shift)
        return set_error(GD_ERR_INVALID, "gd_conv_taps_bf16: null descriptor/operand");
    if (d->n_images <= 0 || d->grid_h <= 0 || d->grid_w <= 0) return set_error(GD_ERR_INVALID, "gd_conv_taps_bf16: empty grid");
    if (d->in_ld <= 0 || d->in_ld % BLOCK_K || d->k_per_tap <= 0 || d->k_per_tap % BLOCK_K || d->c_out <= 0 || d->c_out % 64)
        return set_error(GD_ERR_INVALID, "gd_conv_taps_bf16: in_ld=%d / k_per_tap=%d / c_out=%d must be multiples of 64", d->in_ld,
                         d->k_per_tap, d->c_out);
    if (d->walk < 0 || d->walk > 2 || (d->walk == 1 && (d->in_ld != 64 || d->k_per_tap != 128)) ||
        (d->walk == 2 && (d->in_ld % 128 || d->k_per_tap != d->in_ld)))
        return set_error(GD_ERR_INVALID, "gd_conv_taps_bf16: walk=%d does not fit in_ld=%d / k_per_tap=%d", d->walk, d->in_ld, d->k_per_tap);
    if (d->c_store <= 0 || d->c_store % 32 || d->c_store > d->c_out)
        return set_error(GD_ERR_INVALID, "gd_conv_taps_bf16: c_store=%d must be a multiple of 32 and <= c_out", d->c_store);
    if (d->n_taps < 1 || d->n_taps > GD_CONV_MAX_TAPS) return set_error(GD_ERR_INVALID, "gd_conv_taps_bf16: n_taps out of range");
    if (d->stride != 1 && d->stride != 2) return set_error(GD_ERR_INVALID, "gd_conv_taps_bf16: stride must be 1 or 2");
    if (d->y0 < 0 || d->y1 >= d->grid_h || d->x0 < 0 || d->x1 >= d->grid_w || d->y0 > d->y1 || d->x0 > d->x1)
        return set_error(GD_ERR_INVALID, "gd_conv_taps_bf16: output window outside the grid");
    if (d->out_ld % 8 || d->out_ld < (d->split_out ? 2 : 1) * d->c_store)
        return set_error(GD_ERR_INVALID, "gd_conv_taps_bf16: out_ld must hold the stored channels and be a multiple of 8");
    const int64_t rows = (int64_t)d->n_images * d->grid_h * d->grid_w;
    if (rows >= ((int64_t)1 << 31) - 65536) return set_error(GD_ERR_INVALID, "gd_conv_taps_bf16: too many pixel rows");
    if ((reinterpret_cast<uintptr_t>(d->in) | reinterpret_cast<uintptr_t>(d->W) | reinterpret_cast<uintptr_t>(d->out)) & 15)
        return set_error(GD_ERR_INVALID, "gd_conv_taps_bf16: in/W/out must be 16-byte aligned");
    int rc = check_device();
    if (rc) return rc;
    GemmParams p{};
    p.M = (int)rows, p.N = d->c_out, p.K = d->n_taps * d->k_per_tap;
    p.bias = d->bias, p.scale = d->scale, p.shift = d->shift, p.relu = d->relu;
    p.kb_per_tap = d->k_per_tap / (BLOCK_K * (d->walk ? 2 : 1)), p.a_cols = d->in_ld;  // pipeline stages per tap
    p.split_out = d->split_out, p.c_store = d->c_store;
    for (int t = 0; t < d->n_taps; ++t) p.tap_shift[t] = d->tap_shift[t];
    p.grid_h = d->grid_h, p.grid_w = d->grid_w;
    p.y0 = d->y0, p.y1 = d->y1, p.x0 = d->x0, p.x1 = d->x1, p.stride = d->stride;
    p.out_img_stride = d->out_img_stride, p.out_y_stride = d->out_y_stride, p.out_x_stride = d->out_x_stride;
    p.out_offset = d->out_offset;
    p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(d->out), p.ldo_bf16 = d->out_ld;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
    // operand-reuse stages hold up to four tiles, so those kernels stop at 128-wide tiles (>= 3 stages in 227 KB)
    const int bn = (d->walk == 0 && d->c_out % 256 == 0 && m_tiles * (d->c_out / 256) >= sm_count()) ? 256
                   : (d->c_out % 128 == 0 && m_tiles * (d->c_out / 128) >= sm_count()) ? 128 : 64;
    if (d->walk == 1) {
        if (bn == 128) return launch_conv<128, MODE_CONV_S1>(p, d->in, d->in_ld, d->W, s);
        return launch_conv<64, MODE_CONV_S1>(p, d->in, d->in_ld, d->W, s);
    }
    if (d->walk == 2) {
        if (bn == 128) return launch_conv<128, MODE_CONV_S2>(p, d->in, d->in_ld, d->W, s);
        return launch_conv<64, MODE_CONV_S2>(p, d->in, d->in_ld, d->W, s);
    }
    if (bn == 256) return launch_conv<256, MODE_CONV>(p, d->in, d->in_ld, d->W, s);
    if (bn == 128) return launch_conv<128, MODE_CONV>(p, d->in, d->in_ld, d->W, s);
    return launch_conv<64, MODE_CONV>(p, d->in, d->in_ld, d->W, s);
}
