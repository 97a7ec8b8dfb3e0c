// DDPM ancestral-update arithmetic shared by the standalone kernel and the fused GEMM epilogue.
// Mirrors gaussian_diffusion.py:287-292 (_predict_xstart_from_eps), :207-232
// (q_posterior_mean_variance) and :326-328 (p_sample) of the reference with the same rounding
// order: every product and sum is its own fp32 rounding, exactly like the chain of torch
// elementwise ops, hence the explicit *_rn intrinsics (they forbid FMA contraction).
#pragma once
#include "../../include/gd_b200.h"
#include <cuda_runtime.h>

namespace gd {

struct DdpmStepCoefs {
    float A, B, C1, C2, sig;  // sig already multiplied by the (t != 0) mask
};

__device__ __forceinline__ DdpmStepCoefs ddpm_load_coefs(const gd_ddpm_desc& u, int t) {
    DdpmStepCoefs c;
    c.A = __ldg(u.coef_A + t);
    c.B = __ldg(u.coef_B + t);
    c.C1 = __ldg(u.coef_C1 + t);
    c.C2 = __ldg(u.coef_C2 + t);
    c.sig = (t != 0) ? __ldg(u.sigma + t) : 0.0f;
    return c;
}

// One element of the update. `seed`, `m`, `f` describe the optional in-paint blend
// (generator.py:271-280): x0 <- (1-f)*m*seed + f*m*x0 + (1-m)*x0, evaluated left to right.
__device__ __forceinline__ float ddpm_update_elem(const DdpmStepCoefs& c, float x, float eps, float z, bool inpaint,
                                                  float seed, float m, float f, float clip, float* x0_out,
                                                  float* mean_out = nullptr, float* raw_x0_out = nullptr) {
    float x0 = __fsub_rn(__fmul_rn(c.A, x), __fmul_rn(c.B, eps));
    if (raw_x0_out) *raw_x0_out = x0;
    if (inpaint) {
        float a = __fmul_rn(__fmul_rn(__fsub_rn(1.0f, f), m), seed);
        float b = __fmul_rn(__fmul_rn(f, m), x0);
        float d = __fmul_rn(__fsub_rn(1.0f, m), x0);
        x0 = __fadd_rn(__fadd_rn(a, b), d);
    }
    if (clip > 0.0f) x0 = fminf(fmaxf(x0, -clip), clip);
    if (x0_out) *x0_out = x0;
    float mean = __fadd_rn(__fmul_rn(c.C1, x0), __fmul_rn(c.C2, x));
    if (mean_out) *mean_out = mean;
    return __fadd_rn(mean, __fmul_rn(c.sig, z));
}

}  // namespace gd
