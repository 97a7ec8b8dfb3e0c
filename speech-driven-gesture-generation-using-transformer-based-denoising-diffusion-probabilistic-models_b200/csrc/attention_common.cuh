// Declarations shared by the attention kernels (attention.cu: mma.sync kernels, attention_tc.cu: tcgen05 kernel).
#pragma once
#include "common.cuh"
#include "host_util.h"

namespace gd {

constexpr int ATT_ROW_BYTES = 128;  // 64 bf16 columns = one head (d_k = 64) or two (d_k = 32) per work-item row

struct AttnParams {
    const void* q[2];
    const void* k[2];
    const void* v[2];
    __nv_bfloat16* out[2];
    int q_rows[2], q_ld[2], kv_rows[2], kv_ld[2], out_ld[2];
    int q_stride[2];  // rows between consecutive clips of a query segment (== q_rows unless the segment is a halo view)
    const float *wq, *bq, *wk, *bk, *wv, *bv;
    int heads, Lq, Lk;
    int max_sms;  // host only: SMs the persistent grid is sized for (0 = all)
    float scale_log2;  // d_k^-1/2 * log2(e)
};

__device__ __forceinline__ float4 unpack_bf16x4(const uint2& u) {
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xffff0000u));
}

__device__ __forceinline__ float ex2_approx(float x) {  // MUFU.EX2; -inf -> 0, which is what masked keys need
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
    return y;
}

// row-major bf16 tensor map with a [box_rows x 64-column] box, no swizzle (attention.cu)
int make_rows_tmap(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
// tcgen05 attention for d_k = 64, software-pipelined over items (attention_tc.cu); returns GD_OK or an error
int launch_attention_tc(const AttnParams& p, int n_clips, cudaStream_t s);
size_t attention_tc_smem_bytes(const AttnParams& p);  // shared memory it needs for this shape (one CTA per SM)

}  // namespace gd
