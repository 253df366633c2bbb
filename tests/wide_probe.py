"""Error / speed probe of the wide path cores (debug tool)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import maf_numpy as onp
from tests.util import make_case, rel_err, grad_err
from jaxili_b200.engine import MafEngine
dev = torch.device("cuda:0")
t = lambda a: torch.tensor(np.ascontiguousarray(a), device=dev)
os.environ["MAFB200_FORCE_WIDE"] = "1"
for cfg, B in [(dict(n_in=8, n_cond=4, n_layers=3, layers=[64, 32]), 300),
               (dict(n_in=32, n_cond=128, n_layers=2, layers=[512, 512]), 200),
               (dict(n_in=32, n_cond=128, n_layers=3, layers=[512, 512]), 4096)]:
    spec, flat, std, x, y = make_case(cfg, B, seed=11, scale=0.05 if cfg["n_in"] == 32 else 0.12)
    eng = MafEngine(spec.n_in, spec.n_cond, spec.n_layers, spec.layers, "relu", True, 42, std.shift, std.scale, std.mean, std.std, device=dev)
    p, xt, yt = t(flat), t(x), t(y)
    f64, x64, y64 = flat.astype(np.float64), x.astype(np.float64), y.astype(np.float64)
    ref = onp.nde_log_prob(spec, f64, std, x64, y64)
    l_ref, g_ref = onp.nll_loss_and_grad(spec, f64, std, x64, y64)
    for core in ("ffma", "tf32"):
        eng.set_wide_core(core)
        lp = eng.log_prob(p, xt, yt).cpu().numpy()
        loss, grad = torch.zeros(1, device=dev), torch.zeros_like(p)
        eng.loss_grad(p, xt, yt, loss, grad)
        torch.cuda.synchronize()
        g = grad.cpu().numpy()
        print(cfg["layers"], B, core, "logp rel", rel_err(lp, ref), "max abs", np.abs(lp - ref).max(), "loss rel", rel_err(loss.item(), l_ref),
              "grad err", grad_err(g, g_ref), "masked zero", bool((g[spec.flat_mask() == 0] == 0).all()), flush=True)
