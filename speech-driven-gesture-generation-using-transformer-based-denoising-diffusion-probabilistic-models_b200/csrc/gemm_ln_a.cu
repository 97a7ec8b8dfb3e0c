// LayerNorm as the PROLOGUE of an A-stationary GEMM (gd_linear_ln_bf16):
//
//     out[M, N] = act( LayerNorm(H[M, D]; gamma, beta) · W[N, D]ᵀ + bias )        bf16 out, D = 256 or 512
//
// replaces `self.norm_*(x)` followed by the fused Q|K|V projection or the first FeedForward layer
// (models/nn.py:97-124,158-173 -> transformer.py:51,57,146-152): the normalised bf16 rows never exist in HBM and
// the stand-alone LayerNorm launch disappears.
//
// A work unit is (m-tile of 128 rows, group of consecutive 128-column n-tiles).  Per unit the eight TRANSFORM warps read
// the 128 x D fp32 rows (L2-coherent loads, four rows in flight per warp), normalise them with exactly the arithmetic
// of layernorm_rows_kernel (common.cuh: ln_row_stats / ln_apply_pack) and write the bf16 result straight into the
// K-major, 128-byte-swizzled shared-memory operand layout that tcgen05.mma reads - D/64 slots of 16 KB, the same bytes
// a SWIZZLE_128B TMA load of a bf16 tensor would have produced.  The A tile then stays put while the n-tiles of the
// group sweep over it: warp 0 streams W tiles through a 4-stage TMA ring, warp 1 issues tcgen05.mma (128 x 128 x 16,
// accumulators double-buffered in TMEM), warps 4..11 run the bias / activation / bf16 epilogue with TMA stores.
// D = 256: two A buffers, so the transform of unit i+1 overlaps the MMAs of unit i.  D = 512: the 128 KB tile leaves
// room for one buffer only; the transform of the next unit starts when the last MMA of the current one has retired.
#include "common.cuh"
#include "host_util.h"
#include <cstdlib>

namespace gd {

constexpr int LA_BLOCK_M = 128;
constexpr int LA_BLOCK_K = 64;
constexpr int LA_BN = 128;
constexpr int LA_B_STAGES = 4;
constexpr int LA_EPI_WARPS = 8;
constexpr int LA_TR_WARPS = 8;
constexpr int LA_THREADS = (4 + LA_EPI_WARPS + LA_TR_WARPS) * 32;  // 640
constexpr int LA_A_SLOT = LA_BLOCK_M * LA_BLOCK_K * 2;             // 16 KB: one k-block of the A tile
constexpr int LA_B_STAGE = LA_BN * LA_BLOCK_K * 2;                 // 16 KB
constexpr int LA_STG_PER_WARP = 4096;                              // two 32x32 bf16 staging chunks per epilogue warp

struct LnGemmParams {
    const float* H;
    int ldh;
    int M, N;
    const float* gamma;
    const float* beta;
    float eps;
    const float* bias;
    int act;
    int tiles_per_unit;  // n-tiles swept over one resident A tile
    int n_split;         // units per m-tile
};

template <int D>
struct LnGemmCfg {
    static constexpr int KB = D / LA_BLOCK_K;
    static constexpr int A_BUFS = D <= 256 ? 2 : 1;
    static constexpr int A_BYTES = KB * LA_A_SLOT;
    static constexpr int SMEM_BYTES = A_BUFS * A_BYTES + LA_B_STAGES * LA_B_STAGE + LA_EPI_WARPS * LA_STG_PER_WARP + 1024 + 1024;
};

__device__ __forceinline__ float la_act(float v, int act) {
    if (act == GD_ACT_RELU2) {
        const float r = fmaxf(v, 0.0f);
        return r * r;
    }
    if (act == GD_ACT_SILU) return v / (1.0f + __expf(-v));
    return v;
}

template <int D>
__global__ void __launch_bounds__(LA_THREADS, 1)
gemm_ln_a_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_out, const LnGemmParams p) {
    using Cfg = LnGemmCfg<D>;
    constexpr int KB = Cfg::KB, A_BUFS = Cfg::A_BUFS;
    constexpr int V = D / 128;  // float4 per lane and row
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem_a + A_BUFS * Cfg::A_BYTES;
    uint8_t* staging = smem_b + LA_B_STAGES * LA_B_STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + LA_EPI_WARPS * LA_STG_PER_WARP);
    uint64_t* b_full = bars;                        // [B_STAGES] TMA -> MMA
    uint64_t* b_empty = b_full + LA_B_STAGES;       // [B_STAGES] MMA -> TMA
    uint64_t* a_full = b_empty + LA_B_STAGES;       // [2] transform -> MMA
    uint64_t* a_empty = a_full + 2;                 // [2] MMA -> transform
    uint64_t* acc_full = a_empty + 2;               // [2] MMA -> epilogue
    uint64_t* acc_empty = acc_full + 2;             // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (p.M + LA_BLOCK_M - 1) / LA_BLOCK_M;
    const int n_units = m_tiles * p.n_split;
    const int tpu = p.tiles_per_unit;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tmap_w);
        prefetch_tensormap(&tmap_out);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < LA_B_STAGES; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&a_full[s], LA_TR_WARPS);
            mbar_init(&a_empty[s], 1);
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], LA_EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<256>(tmem_slot);  // two 128-column fp32 accumulator stages
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    pdl_launch_dependents();
    pdl_wait();  // H and out belong to the chain
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ W producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int t0 = (u % p.n_split) * tpu;
                for (int t = 0; t < tpu; ++t) {
                    const int n0 = (t0 + t) * LA_BN;
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(&b_empty[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&b_full[stage], LA_B_STAGE);
                        tma_load_2d(smem_b + stage * LA_B_STAGE, &tmap_w, &b_full[stage], kb * LA_BLOCK_K, n0);
                        if (++stage == LA_B_STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(LA_BLOCK_M, LA_BN);
            int stage = 0;
            uint32_t phase = 0;
            int it_t = 0, it_u = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++it_u) {
                const int buf = it_u % A_BUFS;
                mbar_wait(&a_full[buf], (it_u / A_BUFS) & 1);
                tc_fence_after_sync();
                const uint8_t* a_tile = smem_a + buf * Cfg::A_BYTES;
                for (int t = 0; t < tpu; ++t, ++it_t) {
                    const int acc = it_t & 1;
                    mbar_wait(&acc_empty[acc], ((it_t >> 1) & 1) ^ 1);
                    tc_fence_after_sync();
                    const uint32_t tmem_d = tmem_base + acc * LA_BN;
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(&b_full[stage], phase);
                        tc_fence_after_sync();
                        const uint64_t da = umma_desc_k_sw128(smem_u32(a_tile + kb * LA_A_SLOT));
                        const uint64_t db = umma_desc_k_sw128(smem_u32(smem_b + stage * LA_B_STAGE));
#pragma unroll
                        for (int k = 0; k < LA_BLOCK_K / 16; ++k)
                            umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_commit(&b_empty[stage]);
                        if (++stage == LA_B_STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    umma_commit(&acc_full[acc]);
                }
                umma_commit(&a_empty[buf]);  // every MMA that reads this A buffer has retired when this arrives
            }
        }
    } else if (warp >= 4 && warp < 4 + LA_EPI_WARPS) {
        // ------------------------------------------------------------------ epilogue: TMEM -> bias/act -> bf16 -> TMA store
        const int ew = warp - 4;
        const int quad = ew & 3, half = ew >> 2;  // TMEM lane quadrant (= warp % 4) and column half of the tile
        constexpr int WCOLS = LA_BN / 2;
        uint8_t* stg_base = staging + ew * LA_STG_PER_WARP;
        uint32_t chunk_ctr = 0;
        int it_t = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const int m0 = (u / p.n_split) * LA_BLOCK_M;
            const int t0 = (u % p.n_split) * tpu;
            for (int t = 0; t < tpu; ++t, ++it_t) {
                const int n0 = (t0 + t) * LA_BN;
                const int acc = it_t & 1;
                mbar_wait(&acc_full[acc], (it_t >> 1) & 1);
                tc_fence_after_sync();
                const int row0 = m0 + quad * 32;
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * LA_BN + half * WCOLS;
#pragma unroll 1
                for (int c = 0; c < WCOLS / 32; ++c) {
                    const int col0 = n0 + half * WCOLS + c * 32;
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + c * 32, v);
                    float4 b[8];
                    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) b[j] = __ldg(b4 + j);
                    tmem_ld_wait();
                    if (c == WCOLS / 32 - 1) {  // accumulator stage back to the MMA issuer as soon as it is in registers
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[acc]);
                    }
                    float r[32];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        r[4 * j + 0] = la_act(__uint_as_float(v[4 * j + 0]) + b[j].x, p.act);
                        r[4 * j + 1] = la_act(__uint_as_float(v[4 * j + 1]) + b[j].y, p.act);
                        r[4 * j + 2] = la_act(__uint_as_float(v[4 * j + 2]) + b[j].z, p.act);
                        r[4 * j + 3] = la_act(__uint_as_float(v[4 * j + 3]) + b[j].w, p.act);
                    }
                    uint8_t* stg = stg_base + (chunk_ctr & 1) * 2048;
                    ++chunk_ctr;
                    if (lane == 0) bulk_wait_group_read<1>();  // the store issued two chunks ago has finished reading this slot
                    __syncwarp();
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
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmap_out, stg, col0, row0);
                        bulk_commit_group();
                    }
                }
            }
        }
        if (lane == 0) bulk_wait_group0();
    } else if (warp >= 4 + LA_EPI_WARPS) {
        // ------------------------------------------------------------------ transform: fp32 rows -> LayerNorm -> swizzled bf16 A tile
        const int tw = warp - 4 - LA_EPI_WARPS;
        constexpr int ROWS_PER_WARP = LA_BLOCK_M / LA_TR_WARPS;  // 16
        constexpr int RB = 4;                                    // rows in flight per warp
        // gamma / beta stay in L1 (re-read per row): keeping them in registers would push the D = 512 variant past the
        // 96 registers a 640-thread CTA leaves per thread
        const float4* g4 = reinterpret_cast<const float4*>(p.gamma) + lane;
        const float4* b4 = reinterpret_cast<const float4*>(p.beta) + lane;
        // float4 number (i*32 + lane) of a row = columns [128 i + 4 lane, +4): k-block 2i + lane/16, 16-byte chunk (lane%16)/2,
        // low / high half of the chunk = lane & 1
        const int kb_lane = lane >> 4, chunk = (lane & 15) >> 1, half8 = (lane & 1) * 8;
        int it_u = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++it_u) {
            const int m0 = (u / p.n_split) * LA_BLOCK_M;
            const int buf = it_u % A_BUFS;
            mbar_wait(&a_empty[buf], ((it_u / A_BUFS) & 1) ^ 1);
            uint8_t* a_tile = smem_a + buf * Cfg::A_BYTES;
#pragma unroll 1
            for (int rb = 0; rb < ROWS_PER_WARP / RB; ++rb) {
                float4 v[RB][V];
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    const int row = min(m0 + tw * ROWS_PER_WARP + rb * RB + j, p.M - 1);  // rows past M: any finite data, never stored
                    const float4* xr = reinterpret_cast<const float4*>(p.H + (size_t)row * p.ldh);
#pragma unroll
                    for (int i = 0; i < V; ++i) v[j][i] = __ldcg(xr + i * 32 + lane);
                }
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    float mean, rstd;
                    ln_row_stats<V>(v[j], p.eps, mean, rstd);
                    const int r = tw * ROWS_PER_WARP + rb * RB + j;  // row inside the tile
                    uint8_t* rowp = a_tile + (r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4) + half8;
#pragma unroll
                    for (int i = 0; i < V; ++i)
                        *reinterpret_cast<uint2*>(rowp + (2 * i + kb_lane) * LA_A_SLOT) = ln_apply_pack(v[j][i], mean, rstd, __ldg(g4 + i * 32), __ldg(b4 + i * 32));
                }
            }
            fence_proxy_async();  // generic-proxy writes of the A tile -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[buf]);
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc<256>(tmem_base);
}

// units per m-tile: a divisor of the n-tile count that balances rounds over the SMs against re-normalising the A tile
static int pick_n_split(int m_tiles, int n_tiles, int kb, bool overlapped) {
    const int sms = sm_count();
    const double t_tile = kb * 4 * 64 / 0.6;                          // clocks of one 128x128xD n-tile at ~60 % tensor-pipe
    const double t_ln = overlapped ? 600.0 : 128.0 * kb * 64 * 4 / 40;  // exposed clocks of the prologue (single-buffered: ~40 B/clk)
    int best = 1;
    double best_cost = 1e30;
    for (int ns = 1; ns <= n_tiles; ++ns) {
        if (n_tiles % ns) continue;
        const long units = (long)m_tiles * ns;
        const long rounds = (units + sms - 1) / sms;
        const double cost = rounds * ((n_tiles / ns) * t_tile + t_ln);
        if (cost < best_cost - 1e-9) best_cost = cost, best = ns;
    }
    return best;
}

template <int D>
static int launch_ln_gemm(LnGemmParams& p, const void* W, int ldw, void* out, int ldo, cudaStream_t stream) {
    using Cfg = LnGemmCfg<D>;
    CUtensorMap tw, tout;
    int rc = make_tmap_2d(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, p.N, D, ldw, LA_BLOCK_K, LA_BN, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tmap_2d(&tout, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, p.M, p.N, ldo, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    static bool attr_set[GD_MAX_DEVICES] = {};
    const int dev_idx = current_device();
    if (!attr_set[dev_idx]) {
        GD_CUDA_CHECK(cudaFuncSetAttribute(gemm_ln_a_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set[dev_idx] = true;
    }
    const int m_tiles = (p.M + LA_BLOCK_M - 1) / LA_BLOCK_M, n_tiles = p.N / LA_BN;
    const char* e = getenv("GD_LN_NSPLIT");
    int ns = (e && atoi(e) > 0 && n_tiles % atoi(e) == 0) ? atoi(e) : pick_n_split(m_tiles, n_tiles, Cfg::KB, Cfg::A_BUFS > 1);
    p.n_split = ns, p.tiles_per_unit = n_tiles / ns;
    const int units = m_tiles * ns;
    const int grid = units < sm_count() ? units : sm_count();
    GD_CUDA_CHECK(launch_k(gemm_ln_a_kernel<D>, grid, LA_THREADS, Cfg::SMEM_BYTES, stream, 1, tw, tout, p));
    count_launch();
    GD_CUDA_CHECK(cudaGetLastError());
    return GD_OK;
}

}  // namespace gd

using namespace gd;

extern "C" int gd_linear_ln_bf16(const gd_linear_desc* d, const float* gamma, const float* beta, float eps, void* stream) {
    KindScope kind_scope("gemm");
    if (!d || !d->A || !d->W || !gamma || !beta) return set_error(GD_ERR_INVALID, "gd_linear_ln_bf16: null descriptor/operand");
    if (d->M <= 0 || d->N <= 0) return set_error(GD_ERR_INVALID, "gd_linear_ln_bf16: non-positive shape");
    if (d->K != 256 && d->K != 512) return set_error(GD_ERR_INVALID, "gd_linear_ln_bf16: K=%d must be the model width (256 or 512)", d->K);
    if (d->N % LA_BN) return set_error(GD_ERR_INVALID, "gd_linear_ln_bf16: N=%d must be a multiple of 128", d->N);
    if (d->lda % 4 || d->lda < d->K || d->ldw % 8 || d->ldw < d->K)
        return set_error(GD_ERR_INVALID, "gd_linear_ln_bf16: lda (fp32 elements) / ldw must cover K and keep 16-byte alignment");
    if (!d->out_bf16 || d->out_f32 || d->residual || d->rowbias)
        return set_error(GD_ERR_INVALID, "gd_linear_ln_bf16: bf16 output only (no fp32 output / residual / rowbias)");
    if (!d->bias) return set_error(GD_ERR_INVALID, "gd_linear_ln_bf16: bias required");
    if (d->ldo_bf16 % 8 || d->ldo_bf16 < d->N) return set_error(GD_ERR_INVALID, "gd_linear_ln_bf16: bad output row stride");
    if ((reinterpret_cast<uintptr_t>(d->A) | reinterpret_cast<uintptr_t>(d->W) | reinterpret_cast<uintptr_t>(d->out_bf16) |
         reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) | reinterpret_cast<uintptr_t>(d->bias)) & 15)
        return set_error(GD_ERR_INVALID, "gd_linear_ln_bf16: pointers must be 16-byte aligned");
    int rc = check_device();
    if (rc) return rc;
    LnGemmParams p{};
    p.H = reinterpret_cast<const float*>(d->A), p.ldh = d->lda, p.M = d->M, p.N = d->N;
    p.gamma = gamma, p.beta = beta, p.eps = eps, p.bias = d->bias, p.act = d->act;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (d->K == 256) return launch_ln_gemm<256>(p, d->W, d->ldw, d->out_bf16, d->ldo_bf16, s);
    return launch_ln_gemm<512>(p, d->W, d->ldw, d->out_bf16, d->ldo_bf16, s);
}
