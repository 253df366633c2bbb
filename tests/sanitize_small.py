"""Small invocation of every kernel family for compute-sanitizer (debug tool)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from tests.util import CONFIGS, make_case
from jaxili_b200.engine import MafEngine
dev = torch.device("cuda:0")
t = lambda a: torch.tensor(np.ascontiguousarray(a), device=dev)
def run(cfg, B, act="relu", wide=False, core="tf32"):
    if wide: os.environ["MAFB200_FORCE_WIDE"] = "1"
    else: os.environ.pop("MAFB200_FORCE_WIDE", None)
    spec, flat, std, x, y = make_case(cfg, B, seed=1, activation=act)
    eng = MafEngine(spec.n_in, spec.n_cond, spec.n_layers, spec.layers, act, True, 42, std.shift, std.scale, std.mean, std.std, device=dev)
    if wide: eng.set_wide_core(core)
    p, xt, yt = t(flat), t(x), t(y)
    m = torch.zeros_like(p); v = torch.zeros_like(p); step = torch.zeros(1, dtype=torch.int32, device=dev)
    loss = torch.zeros(1, device=dev); grad = torch.zeros_like(p)
    opt = eng.make_opt_desc(lr=5e-4, warmup=0.1, decay_steps=1e9)
    eng.log_prob(p, xt, yt)
    eng.forward(p, xt, yt)
    for _ in range(2):
        eng.loss_grad(p, xt, yt, loss, grad)
        eng.clip_adam(p, grad, m, v, step, opt)
    if eng.query()['smem_bytes'] > 0:
        u = torch.randn(min(B, 100), spec.n_in, device=dev)
        eng.sample_from_noise(p, u, yt[:1])
        eng.sample_philox(p, 77, yt[:2], seed=3)
        eng.inverse(p, u, yt[:u.shape[0]].contiguous())
    torch.cuda.synchronize()
    print("ok", cfg, B, act, wide, core, loss.item(), flush=True)
run(CONFIGS["c2"], 100)
run(CONFIGS["c2"], 4096)
run(CONFIGS["c1"], 7)
run(CONFIGS["ref_test"], 33, act="silu")
run(dict(n_in=8, n_cond=4, n_layers=2, layers=[64, 32]), 300, wide=True, core="ffma")
run(dict(n_in=8, n_cond=4, n_layers=2, layers=[64, 32]), 300, wide=True, core="tf32")
run(dict(n_in=32, n_cond=128, n_layers=2, layers=[512, 512]), 300, wide=False, core="tf32")
