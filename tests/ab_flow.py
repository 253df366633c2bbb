"""A/B timing of the fused flow kernels (debug tool, not a pytest file): log_prob and loss_grad, chain vs phased
(run twice with MAFB200_FLOW=chain|phased)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, ".")
from tests.util import CONFIGS, make_case
from jaxili_b200.engine import MafEngine
dev = torch.device("cuda:0")
t = lambda a: torch.tensor(np.ascontiguousarray(a), device=dev)
for name, B in (("c2", 4096), ("c2", 65536), ("c3", 16384), ("c1", 128)):
    spec, flat, std, x, y = make_case(CONFIGS[name], B, seed=1)
    eng = MafEngine(spec.n_in, spec.n_cond, spec.n_layers, spec.layers, "relu", True, 42, std.shift, std.scale, std.mean, std.std, device=dev)
    p, xt, yt = t(flat), t(x), t(y)
    loss = torch.zeros(1, device=dev); grad = torch.zeros_like(p)
    eng.pack(p)
    def timeit(fn, n=30):
        for _ in range(5): fn()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / n
    tl = timeit(lambda: eng.log_prob(p, xt, yt, packed_current=True))
    tg = timeit(lambda: eng.loss_grad(p, xt, yt, loss, grad, packed_current=True))
    print(f"{os.environ.get('MAFB200_FLOW','chain'):7s} {name} B={B:6d}: log_prob {tl:8.1f} us   loss_grad(+reduce) {tg:8.1f} us")
