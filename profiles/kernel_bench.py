"""Per-kernel micro-benchmarks through the C-ABI (CUDA events, rotating buffer sets larger than L2).

    python profiles/kernel_bench.py attention|ddpm|gemm|layernorm [--out gpurun_out/x.json]

Environment switches of the library (read once per process) select kernel variants, e.g. GD_ATTN=v1.
One JSON line per case: shape, microseconds per launch, algorithmic GB/s (and TFLOP/s where it applies).
"""
import argparse
import ctypes as C
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th  # noqa: E402
import gesture_b200  # noqa: E402,F401
from gesture_b200 import _lib as gd  # noqa: E402


def stream():
    return th.cuda.current_stream().cuda_stream


QUICK = False  # --quick: exactly one launch per case (ncu captures), no timing


def time_us(fns, iters=20, warm=3):
    """fns: list of callables (one per rotating buffer set); returns mean microseconds per call."""
    if QUICK:
        fns[0]()
        th.cuda.synchronize()
        return float("nan")
    for i in range(warm * len(fns)):
        fns[i % len(fns)]()
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters * len(fns)):
        fns[i % len(fns)]()
    e1.record()
    th.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (iters * len(fns))


def attention_case(N, H, dk, rows_q, rows_kv, shared, sets=3):
    d = H * dk
    lib = gd.load()
    taps = [th.randn(dk, 3, device="cuda") * 0.5 if i % 2 == 0 else th.randn(dk, device="cuda") * 0.1 for i in range(6)]
    fns, keep = [], []
    for _ in range(sets):
        segs_q = [th.randn(N * r, 3 * d, device="cuda").bfloat16() if r else None for r in rows_q]
        segs_kv = segs_q if shared else [th.randn(N * r, 3 * d, device="cuda").bfloat16() if r else None for r in rows_kv]
        outs = [th.zeros(N * r, d, device="cuda", dtype=th.bfloat16) if r else None for r in rows_q]
        a = gd.AttnDesc()
        for s in range(2):
            if rows_q[s]:
                a.q[s], a.q_rows[s], a.q_ld[s] = segs_q[s].data_ptr(), rows_q[s], 3 * d
                a.out[s], a.out_ld[s] = outs[s].data_ptr(), d
            if rows_kv[s]:
                a.k[s] = segs_kv[s].data_ptr() + d * 2
                a.v[s] = segs_kv[s].data_ptr() + 2 * d * 2
                a.kv_rows[s], a.kv_ld[s] = rows_kv[s], 3 * d
        a.conv_wq, a.conv_bq, a.conv_wk, a.conv_bk, a.conv_wv, a.conv_bv = [t.data_ptr() for t in taps]
        a.n_clips, a.heads, a.d_k, a.scale = N, H, dk, 1.0 / math.sqrt(dk)
        keep.append((segs_q, segs_kv, outs, a))
        fns.append(lambda a=a: gd.check(lib.gd_dconv_attention(C.byref(a), stream()), "attn"))
    us = time_us(fns)
    Lq, Lk = sum(rows_q), sum(rows_kv)
    nbytes = N * d * (2 * (Lq + 2 * Lk) + 2 * Lq)
    flops = N * (4 * Lq * Lk * d + 6 * d * (Lq + 2 * Lk))
    return {"kernel": "attention", "variant": os.environ.get("GD_ATTN", "v2"), "N": N, "H": H, "dk": dk, "rows_q": rows_q,
            "rows_kv": rows_kv, "us": round(us, 2), "GBs": round(nbytes / us * 1e-3, 1), "TFLOPs": round(flops / us * 1e-6, 2)}


def ddpm_case(N, C_, T, d, aux, sets=3):
    lib = gd.load()
    M = N * T
    tabs = [th.rand(1000, device="cuda") + 0.5 for _ in range(5)]
    step = th.full((1,), 500, dtype=th.int32, device="cuda")
    W = (th.randn(128, d, device="cuda") * 0.05).bfloat16()
    bias = th.randn(128, device="cuda")
    tape = th.randn(8, N, C_, T, device="cuda")  # a few tape slabs are enough for the access pattern
    step.fill_(3)
    fns, keep = [], []
    for _ in range(sets):
        A = th.randn(M, d, device="cuda").bfloat16()
        x = th.randn(N, C_, T, device="cuda")
        xa = th.zeros(M, 128, device="cuda", dtype=th.bfloat16)
        eps, x0 = th.zeros_like(x), th.zeros_like(x)
        dd = gd.LinearDesc()
        dd.A, dd.W, dd.M, dd.N, dd.K, dd.lda, dd.ldw, dd.bias = A.data_ptr(), W.data_ptr(), M, 128, d, d, d, bias.data_ptr()
        u = gd.DdpmDesc()
        u.x, u.noise_tape = x.data_ptr(), tape.data_ptr()
        u.coef_A, u.coef_B, u.coef_C1, u.coef_C2, u.sigma = [t.data_ptr() for t in tabs]
        u.step_ptr, u.n_clips, u.C, u.T = step.data_ptr(), N, C_, T
        if aux:
            u.eps_out, u.x0_out = eps.data_ptr(), x0.data_ptr()
        u.xa_bf16, u.ld_xa = xa.data_ptr(), 128
        keep.append((A, x, xa, eps, x0, dd, u))
        fns.append(lambda dd=dd, u=u: gd.check(lib.gd_linear_ddpm(C.byref(dd), C.byref(u), stream()), "ddpm"))
    us = time_us(fns)
    nbytes = 2 * M * d + 4 * N * C_ * T * (3 + (2 if aux else 0)) + 2 * M * 128
    return {"kernel": "gemm_ddpm", "N": N, "C": C_, "T": T, "d": d, "aux": aux, "us": round(us, 2),
            "GBs": round(nbytes / us * 1e-3, 1)}


def gemm_case(M, N, K, mode, sets=3):
    lib = gd.load()
    W = (th.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = th.randn(N, device="cuda")
    fns, keep = [], []
    for _ in range(sets):
        A = th.randn(M, K, device="cuda").bfloat16()
        d = gd.LinearDesc()
        d.A, d.W, d.M, d.N, d.K, d.lda, d.ldw, d.bias = A.data_ptr(), W.data_ptr(), M, N, K, K, K, bias.data_ptr()
        if mode == "bf16":
            out = th.empty(M, N, device="cuda", dtype=th.bfloat16)
            d.out_bf16, d.ldo_bf16 = out.data_ptr(), N
        elif mode == "relu2":
            out = th.empty(M, N, device="cuda", dtype=th.bfloat16)
            d.out_bf16, d.ldo_bf16, d.act = out.data_ptr(), N, gd.ACT_RELU2
        else:  # in-place fp32 residual
            out = th.zeros(M, N, device="cuda")
            d.out_f32, d.ldo_f32, d.residual, d.ldr = out.data_ptr(), N, out.data_ptr(), N
        keep.append((A, out, d))
        fns.append(lambda d=d: gd.check(lib.gd_linear_bf16(C.byref(d), stream()), "gemm"))
    us = time_us(fns)
    flops = 2.0 * M * N * K
    nbytes = 2 * M * K + 2 * N * K + (2 * M * N if mode != "resid" else 8 * M * N)
    return {"kernel": "gemm", "variant": os.environ.get("GD_GEMM", "default"), "M": M, "N": N, "K": K, "mode": mode,
            "us": round(us, 2), "TFLOPs": round(flops / us * 1e-6, 1), "GBs": round(nbytes / us * 1e-3, 1)}


def resid_ln_case(M, N, K, sets=3):
    lib = gd.load()
    W = (th.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias, gam, bet = th.randn(N, device="cuda"), th.rand(N, device="cuda") + 0.5, th.randn(N, device="cuda")
    fns, keep = [], []
    for _ in range(sets):
        A = th.randn(M, K, device="cuda").bfloat16()
        H = th.zeros(M, N, device="cuda")
        xn = th.empty(M, N, device="cuda", dtype=th.bfloat16)
        d = gd.LinearDesc()
        d.A, d.W, d.M, d.N, d.K, d.lda, d.ldw, d.bias = A.data_ptr(), W.data_ptr(), M, N, K, K, K, bias.data_ptr()
        d.out_f32, d.ldo_f32, d.residual, d.ldr = H.data_ptr(), N, H.data_ptr(), N
        ln = gd.LnDesc()
        ln.gamma, ln.beta, ln.out_bf16, ln.ldo, ln.eps = gam.data_ptr(), bet.data_ptr(), xn.data_ptr(), N, 1e-5
        keep.append((A, H, xn, d, ln))
        fns.append(lambda d=d, ln=ln: gd.check(lib.gd_linear_resid_ln(C.byref(d), C.byref(ln), stream()), "resid_ln"))
    us = time_us(fns)
    nbytes = 2 * M * K + 2 * N * K + 8 * M * N + 2 * M * N
    return {"kernel": "gemm_resid_ln", "M": M, "N": N, "K": K, "us": round(us, 2),
            "TFLOPs": round(2.0 * M * N * K / us * 1e-6, 1), "GBs": round(nbytes / us * 1e-3, 1)}


def layernorm_case(M, D, sets=3):
    lib = gd.load()
    g, b = th.randn(D, device="cuda"), th.randn(D, device="cuda")
    fns, keep = [], []
    for _ in range(sets):
        x = th.randn(M, D, device="cuda")
        o = th.empty(M, D, device="cuda", dtype=th.bfloat16)
        keep.append((x, o))
        fns.append(lambda x=x, o=o: gd.check(lib.gd_layernorm(x.data_ptr(), D, g.data_ptr(), b.data_ptr(), o.data_ptr(), D,
                                                             M, D, 1e-5, stream()), "ln"))
    us = time_us(fns)
    return {"kernel": "layernorm", "M": M, "D": D, "us": round(us, 2), "GBs": round(6 * M * D / us * 1e-3, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["attention", "ddpm", "gemm", "layernorm", "resid_ln"])
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true", help="one launch set per case (ncu captures)")
    args = ap.parse_args()
    global QUICK
    QUICK = args.quick
    res = []
    if args.what == "attention":
        cases = [  # tedexp-ours N=256: pose self, memory self, joint, last-layer joint; beat-ours N=1024: self, cross
            (256, 8, 64, (34, 0), (34, 0), True), (256, 8, 64, (104, 0), (104, 0), True),
            (256, 8, 64, (34, 104), (34, 104), True), (256, 8, 64, (34, 0), (34, 104), False),
            (1024, 8, 32, (40, 0), (40, 0), True), (1024, 8, 32, (40, 0), (32, 0), False),
            (64, 8, 32, (160, 0), (160, 0), True), (64, 8, 32, (160, 0), (127, 0), False)]
        for c in cases:
            res.append(attention_case(*c, sets=1 if args.quick else 3))
    elif args.what == "ddpm":
        for N, C_, T, d in [(256, 126, 34, 512), (1024, 123, 40, 256)]:
            for aux in (True, False):
                res.append(ddpm_case(N, C_, T, d, aux, sets=1 if args.quick else 3))
    elif args.what == "gemm":
        R, Mx = 35328, 8704
        cases = [(R, 1536, 512, "bf16"), (R, 512, 512, "resid"), (R, 2048, 512, "relu2"), (R, 512, 2048, "resid"),
                 (26624, 1536, 512, "bf16"), (26624, 2048, 512, "relu2"), (26624, 512, 2048, "resid"),
                 (Mx, 1536, 512, "bf16"), (Mx, 512, 512, "resid"), (Mx, 2048, 512, "relu2"), (Mx, 512, 2048, "resid"),
                 (40960, 768, 256, "bf16"), (40960, 256, 256, "resid"), (40960, 1024, 256, "relu2"),
                 (40960, 256, 1024, "resid")]
        for c in cases:
            res.append(gemm_case(*c, sets=1 if args.quick else 3))
    elif args.what == "resid_ln":
        for M, N, K in [(35328, 512, 512), (26624, 512, 512), (8704, 512, 512), (35328, 512, 2048), (26624, 512, 2048),
                        (8704, 512, 2048), (40960, 256, 256), (40960, 256, 1024)]:
            res.append(resid_ln_case(M, N, K, sets=1 if args.quick else 3))
    else:
        for M, D in [(35328, 512), (26624, 512), (8704, 512), (40960, 256)]:
            res.append(layernorm_case(M, D, sets=1 if args.quick else 3))
    for r in res:
        print(json.dumps(r))
    if args.out:
        with open(args.out, "a") as f:
            for r in res:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
