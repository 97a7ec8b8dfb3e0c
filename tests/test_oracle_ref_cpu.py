"""The oracle port (oracle/ddpm_oracle.py) against the UNMODIFIED reference staged under oracle/_ref/ by oracle/make_ref.py,
run live on CPU: the restatement is validated by the real thing, not only by committed vectors.  Skipped when the staged
copy is absent (it is git-ignored; `__graft_entry__.build()` stages it whenever /root/reference exists)."""
import os
import sys

import pytest
import torch as th

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_runner  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_runner.available(), reason="oracle/_ref not staged (python oracle/make_ref.py)")


def test_staged_reference_is_unmodified():
    assert ref_runner.verify() >= 30


@pytest.mark.parametrize("workload", ["beat-ours", "tedexp-ours"])
def test_oracle_port_matches_live_reference(workload):
    """4 ancestral steps of the reference's own p_sample_loop_progressive on its seed-0 weights vs the port, same tape:
    per-step eps and samples agree to fp32 round-off (and the full dict of p_sample is what the port's update implies)."""
    from oracle import ddpm_oracle as orc
    gen, model, diffusion, C, T, L = ref_runner.build(workload)
    N, n_steps = 2, 4
    g = th.Generator().manual_seed(5)
    wav = th.randn(N, L, generator=g)
    x_T = th.randn(N, C, T, generator=g)
    tape = [th.randn(N, C, T, generator=g) for _ in range(n_steps)]
    it, real = iter(tape), th.randn_like
    th.randn_like = lambda x: next(it)
    outs = []
    try:
        with th.no_grad():
            loop = diffusion.p_sample_loop_progressive(model, (N, C, T), noise=x_T, model_kwargs={"wav": wav}, device="cpu")
            for _ in range(n_steps):
                outs.append({k: v.clone() for k, v in next(loop).items()})
    finally:
        th.randn_like = real
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model_type = "s2g_v2" if workload == "beat-ours" else "default"
    tabs = orc.spaced_diffusion_tables("linear", 1000, "")
    rec = []
    orc.sample_chain(sd, model_type, 8, tabs, x_T, wav, th.stack(tape), steps=n_steps, reencode_every_step=True,
                     record=lambda i, x, eps, xn: rec.append((i, eps.clone(), xn.clone())))
    assert [r[0] for r in rec] == [999, 998, 997, 996]
    for (i, eps, xn), out in zip(rec, outs):
        assert set(out) == {"sample", "mean", "variance", "log_variance", "eps", "pred_x_start", "raw_x_start"}
        rel = lambda a, b: ((a - b).norm() / b.norm()).item()  # noqa: E731
        assert rel(eps, out["eps"]) < 1e-4, (workload, i, rel(eps, out["eps"]))
        assert rel(xn, out["sample"]) < 1e-4, (workload, i, rel(xn, out["sample"]))
