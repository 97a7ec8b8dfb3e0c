// Residual GEMM with the following LayerNorm fused into its epilogue (gd_b200.h: gd_linear_resid_ln):
//
//     H[M, d] += A[M, K] · W[d, K]ᵀ + bias            (attention out-projection / FFN down-projection + residual add)
//     xn[M, d] = LayerNorm_d(H) * gamma + beta  (bf16)  (the operand of the next block's first GEMM)
//
// A row of the residual stream (d = 256 or 512 fp32) has to be complete before it can be normalised, so one CTA
// (d = 256) or a cluster of two CTAs (d = 512, each owning 256 of the columns of the SAME 128 rows) covers full rows.
// Mainloop as in gemm_tcgen05.cu (TMA ring -> tcgen05.mma -> two TMEM accumulator stages).  Epilogue, 8 warps, a warp
// = 32 rows x 128 columns in four 32-column chunks:
//   pass 1  residual chunk arrives by TMA (issued one chunk ahead) -> x = acc + bias + h -> shifted row sums ->
//           x goes back into TMEM (tcgen05.st) and, through a swizzled staging tile, to H by TMA store
//   stats   per-row (mean, M2) of each 128-column part are exchanged through shared memory - of both CTAs of the
//           cluster (st.shared::cluster + cluster-scope mbarrier) - and merged with Chan's formula
//   pass 2  x from TMEM -> (x - mean) * rstd * gamma + beta -> bf16 -> staging -> TMA store to xn
// so the stand-alone LayerNorm kernel (read H, write xn) and the L2 reduce-add read-modify-write disappear.
#include "common.cuh"
#include "host_util.h"

namespace gd {

constexpr int RL_BLOCK_M = 128;
constexpr int RL_BN = 256;
constexpr int RL_BLOCK_K = 64;
constexpr int RL_UMMA_K = 16;
constexpr int RL_THREADS = 384;
constexpr int RL_EPI_WARPS = 8;
constexpr int RL_A_BYTES = RL_BLOCK_M * RL_BLOCK_K * 2;
constexpr int RL_B_BYTES = RL_BN * RL_BLOCK_K * 2;
constexpr int RL_STAGE_BYTES = RL_A_BYTES + RL_B_BYTES;
constexpr int RL_STAGES = 3;
constexpr int RL_STAGING_PER_WARP = 8192;  // [residual-in 4 KB | out 4 KB]
constexpr int RL_STAGING_BYTES = RL_EPI_WARPS * RL_STAGING_PER_WARP;
constexpr int RL_STATS_BYTES = 2 * 4 * RL_BLOCK_M * 8;  // [tile parity][column part][row] x (mean, M2)
constexpr int RL_VEC_BYTES = 5 * RL_BN * 4;  // this CTA's 256 columns of bias | gamma0 | beta0 | gamma1 | beta1
constexpr int RL_SMEM_BYTES = RL_STAGES * RL_STAGE_BYTES + 1024 + RL_STAGING_BYTES + RL_STATS_BYTES + RL_VEC_BYTES + 1024;

struct ResidLnParams {
    int M, N, K;
    const float* bias;
    const float *gamma0, *beta0, *gamma1, *beta1;  // rows < split_row use set 0, the others set 1
    int split_row;
    float eps;
};

__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const float (&x)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
        "f"(x[0]), "f"(x[1]), "f"(x[2]), "f"(x[3]), "f"(x[4]), "f"(x[5]), "f"(x[6]), "f"(x[7]), "f"(x[8]), "f"(x[9]),
        "f"(x[10]), "f"(x[11]), "f"(x[12]), "f"(x[13]), "f"(x[14]), "f"(x[15]), "f"(x[16]), "f"(x[17]), "f"(x[18]),
        "f"(x[19]), "f"(x[20]), "f"(x[21]), "f"(x[22]), "f"(x[23]), "f"(x[24]), "f"(x[25]), "f"(x[26]), "f"(x[27]),
        "f"(x[28]), "f"(x[29]), "f"(x[30]), "f"(x[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_bar_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_bar_addr) : "memory");
}
__device__ __forceinline__ void st_cluster_f2(uint32_t cluster_addr, float a, float b) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};\n" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}

template <int NSPLIT>
__global__ void __launch_bounds__(RL_THREADS, 1)
gemm_resid_ln_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const __grid_constant__ CUtensorMap tmap_h, const __grid_constant__ CUtensorMap tmap_xn,
                     const __grid_constant__ CUtensorMap tmap_hpf, const ResidLnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + RL_STAGES * RL_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RL_STAGES * RL_STAGE_BYTES);
    uint64_t* full_bar = bars;                        // [STAGES]
    uint64_t* empty_bar = bars + RL_STAGES;           // [STAGES]
    uint64_t* acc_full_bar = bars + 2 * RL_STAGES;    // [2]
    uint64_t* acc_empty_bar = acc_full_bar + 2;       // [2]
    uint64_t* resid_bar = acc_empty_bar + 2;          // [EPI_WARPS] residual chunk landed
    uint64_t* stats_bar = resid_bar + RL_EPI_WARPS;   // [1] all column parts of a tile have published their row stats
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stats_bar + 1);
    uint8_t* staging = smem + RL_STAGES * RL_STAGE_BYTES + 1024;
    float2* stats = reinterpret_cast<float2*>(staging + RL_STAGING_BYTES);
    // per-column vectors live in shared memory: with ~220 KB of it carved out there is no L1 left to cache them
    float* s_vec = reinterpret_cast<float*>(staging + RL_STAGING_BYTES + RL_STATS_BYTES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int cta_rank = (NSPLIT > 1) ? static_cast<int>(cluster_ctarank()) : 0;
    const int cluster_id = blockIdx.x / NSPLIT, num_clusters = gridDim.x / NSPLIT;
    const int m_tiles = (p.M + RL_BLOCK_M - 1) / RL_BLOCK_M;
    const int k_blocks = p.K / RL_BLOCK_K;
    const int n0 = cta_rank * RL_BN;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tmap_a);
        prefetch_tensormap(&tmap_b);
        prefetch_tensormap(&tmap_h);
        prefetch_tensormap(&tmap_xn);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < RL_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full_bar[s], 1);
            mbar_init(&acc_empty_bar[s], RL_EPI_WARPS);
        }
        for (int s = 0; s < RL_EPI_WARPS; ++s) mbar_init(&resid_bar[s], 1);
        mbar_init(stats_bar, RL_EPI_WARPS * NSPLIT);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_slot);
    for (int i = threadIdx.x; i < 5 * RL_BN; i += RL_THREADS) {
        const int which = i / RL_BN, c = n0 + i % RL_BN;
        const float* src = which == 0 ? p.bias : (which == 1 ? p.gamma0 : (which == 2 ? p.beta0 : (which == 3 ? p.gamma1 : p.beta1)));
        s_vec[i] = __ldg(src + c);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (NSPLIT > 1) cluster_sync_all();  // the peer's barriers exist before anything is sent to them
    tc_fence_after_sync();
    pdl_launch_dependents();
    pdl_wait();  // bias/gamma/beta above are weights; A, H and xn belong to the chain
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < m_tiles; tile += num_clusters) {
                const int m0 = tile * RL_BLOCK_M;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], RL_STAGE_BYTES);
                    tma_load_2d(smem_a + stage * RL_A_BYTES, &tmap_a, &full_bar[stage], kb * RL_BLOCK_K, m0);
                    tma_load_2d(smem_b + stage * RL_B_BYTES, &tmap_b, &full_bar[stage], kb * RL_BLOCK_K, n0);
                    if (++stage == RL_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(RL_BLOCK_M, RL_BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = cluster_id; tile < m_tiles; tile += num_clusters, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&acc_empty_bar[acc], acc_phase ^ 1);
                tc_fence_after_sync();
                const uint32_t tmem_d = tmem_base + acc * RL_BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after_sync();
                    const uint64_t da = umma_desc_k_sw128(smem_u32(smem_a + stage * RL_A_BYTES));
                    const uint64_t db = umma_desc_k_sw128(smem_u32(smem_b + stage * RL_B_BYTES));
#pragma unroll
                    for (int k = 0; k < RL_BLOCK_K / RL_UMMA_K; ++k)
                        umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(&empty_bar[stage]);
                    if (++stage == RL_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&acc_full_bar[acc]);
            }
        }
    } else if (warp == 3) {
        // Residual prefetcher: the epilogue's 4-KB residual chunks are latency-critical and only one per warp can be
        // in flight in shared memory, so the 128 x 256 fp32 block of H of the tile after next is pulled into L2 early
        // (one TMA prefetch instruction); the chunk loads then see L2 latency instead of HBM latency.
        if (lane == 0) {
            prefetch_tensormap(&tmap_hpf);
            int it = 0;
            for (int tile = cluster_id; tile < m_tiles; tile += num_clusters, ++it) {
                if (it == 0) {
                    tma_prefetch_l2_2d(&tmap_hpf, n0, tile * RL_BLOCK_M);
                    if (tile + num_clusters < m_tiles) tma_prefetch_l2_2d(&tmap_hpf, n0, (tile + num_clusters) * RL_BLOCK_M);
                }
                mbar_wait(&acc_full_bar[it & 1], (it >> 1) & 1);  // tile `it` enters its epilogue
                if (tile + 2 * num_clusters < m_tiles) tma_prefetch_l2_2d(&tmap_hpf, n0, (tile + 2 * num_clusters) * RL_BLOCK_M);
            }
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;
        const int quad = ew & 3;   // TMEM lanes [32*quad, +32)
        const int chalf = ew >> 2;  // which 128 columns of this CTA's 256
        const int part = cta_rank * 2 + chalf;
        const int colbase = n0 + chalf * 128;
        const int rloc = quad * 32 + lane;
        uint8_t* slot_in = staging + ew * RL_STAGING_PER_WARP;
        uint8_t* slot_out = slot_in + 4096;
        uint64_t* rbar = &resid_bar[ew];
        uint32_t rphase = 0, sphase = 0;
        const float inv_part = 1.0f / 128.0f;
        int tile = cluster_id;
        if (tile < m_tiles && lane == 0) {  // first residual chunk, while the first accumulator is being computed
            mbar_arrive_expect_tx(rbar, 4096);
            tma_load_2d(slot_in, &tmap_h, rbar, colbase, tile * RL_BLOCK_M + quad * 32);
        }
        int it = 0;
        for (; tile < m_tiles; tile += num_clusters, ++it) {
            const int m0 = tile * RL_BLOCK_M;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(&acc_full_bar[acc], acc_phase);
            tc_fence_after_sync();
            const int row0 = m0 + quad * 32;
            const int row = row0 + lane;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * RL_BN + chalf * 128;
            float s1 = 0.f, s2 = 0.f, shift = 0.f;
            // ---------------- pass 1: x = acc + bias + h; H <- x; row sums
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int col0 = colbase + c * 32;
                uint32_t v[32];
                tmem_ld_32x32b_x32(taddr + c * 32, v);
                float4 b[8];
                {
                    const float4* b4 = reinterpret_cast<const float4*>(s_vec + (col0 - n0));
#pragma unroll
                    for (int j = 0; j < 8; ++j) b[j] = b4[j];
                }
                mbar_wait(rbar, rphase);
                rphase ^= 1;
                float4 r[8];
                {
                    const float4* in4 = reinterpret_cast<const float4*>(slot_in);  // 128-B rows, SWIZZLE_128B
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[j] = in4[lane * 8 + (j ^ (lane & 7))];
                }
                __syncwarp();  // every lane has read the slot: it can be refilled
                if (lane == 0) {
                    int nc = c + 1, nt = tile;
                    if (nc == 4) nc = 0, nt = tile + num_clusters;
                    if (nt < m_tiles) {
                        mbar_arrive_expect_tx(rbar, 4096);
                        tma_load_2d(slot_in, &tmap_h, rbar, colbase + nc * 32, nt * RL_BLOCK_M + quad * 32);
                    }
                }
                tmem_ld_wait();
                float x[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    x[4 * j + 0] = (__uint_as_float(v[4 * j + 0]) + b[j].x) + r[j].x;
                    x[4 * j + 1] = (__uint_as_float(v[4 * j + 1]) + b[j].y) + r[j].y;
                    x[4 * j + 2] = (__uint_as_float(v[4 * j + 2]) + b[j].z) + r[j].z;
                    x[4 * j + 3] = (__uint_as_float(v[4 * j + 3]) + b[j].w) + r[j].w;
                }
                if (c == 0) shift = x[0];  // shifted-data sums: no cancellation when |mean| >> std
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float d = x[j] - shift;
                    s1 += d;
                    s2 = fmaf(d, d, s2);
                }
                tmem_st_32x32b_x32(taddr + c * 32, x);
                if (lane == 0) bulk_wait_group_read0();  // the previous store has finished reading slot_out
                __syncwarp();
                {
                    float4* st4 = reinterpret_cast<float4*>(slot_out);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        st4[lane * 8 + (j ^ (lane & 7))] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmap_h, slot_out, col0, row0);
                    bulk_commit_group();
                }
            }
            tmem_st_wait();
            // ---------------- row statistics of this 128-column part -> every CTA of the cluster
            const float mean_p = shift + s1 * inv_part;
            const float m2_p = fmaxf(s2 - s1 * s1 * inv_part, 0.f);
            const int par = it & 1;
            float2* mine = stats + (par * 4 + part) * RL_BLOCK_M + rloc;
            *mine = make_float2(mean_p, m2_p);
            if (NSPLIT > 1) st_cluster_f2(map_to_cta(smem_u32(mine), cta_rank ^ 1), mean_p, m2_p);
            __syncwarp();
            // Both CTAs count 16 arrivals per tile on their own barrier.  A CTA can only arrive for tile i+1 after it has
            // passed its own barrier of tile i, i.e. after all 16 warps have issued their arrives for tile i (a warp issues the
            // two back to back), so a phase is never completed by arrivals of the next one in practice: the peer's second
            // arrive would have to stay in flight for the thousands of cycles of pass 2 + the next tile's pass 1.
            if (lane == 0) {
                if (NSPLIT > 1) {  // one cluster-scope release fence, then two relaxed arrives
                    fence_acq_rel_cluster();
                    mbar_arrive_cluster_relaxed(map_to_cta(smem_u32(stats_bar), 0));
                    mbar_arrive_cluster_relaxed(map_to_cta(smem_u32(stats_bar), 1));
                } else {
                    mbar_arrive(stats_bar);
                }
            }
            mbar_wait(stats_bar, sphase);  // CTA-scope spin ...
            if (NSPLIT > 1) fence_acq_rel_cluster();  // ... then one cluster-scope acquire for the peer's statistics
            sphase ^= 1;
            float mean = 0.f;
            float2 st[2 * NSPLIT];
#pragma unroll
            for (int q = 0; q < 2 * NSPLIT; ++q) {
                st[q] = stats[(par * 4 + q) * RL_BLOCK_M + rloc];
                mean += st[q].x;
            }
            mean *= 1.0f / (2 * NSPLIT);
            float m2 = 0.f;
#pragma unroll
            for (int q = 0; q < 2 * NSPLIT; ++q) {
                const float dm = st[q].x - mean;
                m2 += st[q].y + 128.0f * dm * dm;
            }
            const float rstd = rsqrtf(m2 * (1.0f / (256.0f * NSPLIT)) + p.eps);
            const float* gam = s_vec + (row < p.split_row ? 1 : 3) * RL_BN - n0;  // indexed by global column below
            const float* bet = gam + RL_BN;
            // ---------------- pass 2: xn = (x - mean) * rstd * gamma + beta
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int col0 = colbase + c * 32;
                uint32_t v[32];
                tmem_ld_32x32b_x32(taddr + c * 32, v);
                float4 g4[8], e4[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    g4[j] = reinterpret_cast<const float4*>(gam + col0)[j];
                    e4[j] = reinterpret_cast<const float4*>(bet + col0)[j];
                }
                tmem_ld_wait();
                float y[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    y[4 * j + 0] = fmaf((__uint_as_float(v[4 * j + 0]) - mean) * rstd, g4[j].x, e4[j].x);
                    y[4 * j + 1] = fmaf((__uint_as_float(v[4 * j + 1]) - mean) * rstd, g4[j].y, e4[j].y);
                    y[4 * j + 2] = fmaf((__uint_as_float(v[4 * j + 2]) - mean) * rstd, g4[j].z, e4[j].z);
                    y[4 * j + 3] = fmaf((__uint_as_float(v[4 * j + 3]) - mean) * rstd, g4[j].w, e4[j].w);
                }
                if (lane == 0) bulk_wait_group_read0();
                __syncwarp();
                {
                    uint4* st4 = reinterpret_cast<uint4*>(slot_out);  // 64-B rows, SWIZZLE_64B: chunk ^= (row >> 1) & 3
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 w;
                        w.x = pack_bf16x2(y[8 * j + 0], y[8 * j + 1]);
                        w.y = pack_bf16x2(y[8 * j + 2], y[8 * j + 3]);
                        w.z = pack_bf16x2(y[8 * j + 4], y[8 * j + 5]);
                        w.w = pack_bf16x2(y[8 * j + 6], y[8 * j + 7]);
                        st4[lane * 4 + (j ^ ((lane >> 1) & 3))] = w;
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmap_xn, slot_out, col0, row0);
                    bulk_commit_group();
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty_bar[acc]);
        }
        if (lane == 0) bulk_wait_group0();
    }

    tc_fence_before_sync();
    __syncthreads();
    if (NSPLIT > 1) cluster_sync_all();  // no CTA leaves while its peer may still write statistics into it
    if (warp == 2) tmem_dealloc<512>(tmem_base);
}

template <int NSPLIT>
static int launch_resid_ln(const gd_linear_desc* d, const gd_ln_desc* ln, cudaStream_t stream) {
    CUtensorMap ta, tb, th, tx, tpf;
    int rc = make_tmap_2d(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d->A, d->M, d->K, d->lda, RL_BLOCK_K, RL_BLOCK_M,
                          CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tmap_2d(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d->W, d->N, d->K, d->ldw, RL_BLOCK_K, RL_BN,
                      CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tmap_2d(&th, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d->out_f32, d->M, d->N, d->ldo_f32, 32, 32,
                      CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tmap_2d(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ln->out_bf16, d->M, d->N, ln->ldo, 32, 32,
                      CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    rc = make_tmap_2d(&tpf, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d->out_f32, d->M, d->N, d->ldo_f32, RL_BN, RL_BLOCK_M,
                      CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    static bool attr_set[GD_MAX_DEVICES] = {};  // function attributes are per device
    const int dev_idx = current_device();
    if (!attr_set[dev_idx]) {
        GD_CUDA_CHECK(cudaFuncSetAttribute(gemm_resid_ln_kernel<NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           RL_SMEM_BYTES));
        attr_set[dev_idx] = true;
    }
    ResidLnParams p{};
    p.M = d->M, p.N = d->N, p.K = d->K, p.bias = d->bias;
    p.gamma0 = ln->gamma, p.beta0 = ln->beta;
    p.gamma1 = ln->gamma2 ? ln->gamma2 : ln->gamma, p.beta1 = ln->beta2 ? ln->beta2 : ln->beta;
    p.split_row = ln->gamma2 ? ln->split_row : d->M;
    p.eps = ln->eps;
    const int m_tiles = (d->M + RL_BLOCK_M - 1) / RL_BLOCK_M;
    const int max_clusters = sm_count() / NSPLIT;
    const int grid = (m_tiles < max_clusters ? m_tiles : max_clusters) * NSPLIT;
    GD_CUDA_CHECK(launch_k(gemm_resid_ln_kernel<NSPLIT>, grid, RL_THREADS, RL_SMEM_BYTES, stream, NSPLIT, ta, tb, th, tx, tpf, p));
    count_launch();
    GD_CUDA_CHECK(cudaGetLastError());
    return GD_OK;
}

}  // namespace gd

using namespace gd;

extern "C" int gd_linear_resid_ln(const gd_linear_desc* d, const gd_ln_desc* ln, void* stream) {
    gd::KindScope kind_scope("gemm");
    if (!d || !ln || !d->A || !d->W) return set_error(GD_ERR_INVALID, "gd_linear_resid_ln: null descriptor/operand");
    if (d->M <= 0 || d->K <= 0 || d->K % RL_BLOCK_K) return set_error(GD_ERR_INVALID, "gd_linear_resid_ln: K must be a positive multiple of 64");
    if (d->N != 256 && d->N != 512) return set_error(GD_ERR_INVALID, "gd_linear_resid_ln: N=%d must be the model width 256 or 512", d->N);
    if (!d->out_f32 || d->residual != d->out_f32 || d->ldr != d->ldo_f32 || d->ldo_f32 % 4 || d->ldo_f32 < d->N)
        return set_error(GD_ERR_INVALID, "gd_linear_resid_ln: residual must alias out_f32 (in-place residual stream), 16-byte rows");
    if (!d->bias || d->rowbias || d->act != GD_ACT_NONE || d->out_bf16)
        return set_error(GD_ERR_INVALID, "gd_linear_resid_ln: bias required; rowbias/activation/out_bf16 not supported here");
    if (!ln->gamma || !ln->beta || !ln->out_bf16 || ln->ldo % 8 || ln->ldo < d->N)
        return set_error(GD_ERR_INVALID, "gd_linear_resid_ln: LayerNorm gamma/beta/out_bf16 (16-byte rows) required");
    if (ln->gamma2 && (!ln->beta2 || ln->split_row < 0)) return set_error(GD_ERR_INVALID, "gd_linear_resid_ln: bad second parameter set");
    if (d->lda % 8 || d->ldw % 8 || d->lda < d->K || d->ldw < d->K)
        return set_error(GD_ERR_INVALID, "gd_linear_resid_ln: lda/ldw must be >= K and multiples of 8");
    if ((reinterpret_cast<uintptr_t>(d->A) | reinterpret_cast<uintptr_t>(d->W) | reinterpret_cast<uintptr_t>(d->out_f32) |
         reinterpret_cast<uintptr_t>(ln->out_bf16)) & 15)
        return set_error(GD_ERR_INVALID, "gd_linear_resid_ln: operands must be 16-byte aligned");
    int rc = check_device();
    if (rc) return rc;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    return d->N == 512 ? launch_resid_ln<2>(d, ln, s) : launch_resid_ln<1>(d, ln, s);
}
