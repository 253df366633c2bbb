"""Compare plan variants (MAFB200_OCC=1|2) on train-step / log_prob time (debug tool)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, ".")
from tests.util import CONFIGS, make_case
from jaxili_b200.engine import MafEngine
dev = torch.device("cuda:0")
t = lambda a: torch.tensor(np.ascontiguousarray(a), device=dev)
for name, B in [("c2", 4096), ("c3", 16384), ("c1", 128), ("c2", 65536)]:
    spec, flat, std, x, y = make_case(CONFIGS[name], B, seed=1)
    for occ in ("1", "2"):
        os.environ["MAFB200_OCC"] = occ
        eng = MafEngine(spec.n_in, spec.n_cond, spec.n_layers, spec.layers, "relu", True, 42, std.shift, std.scale, std.mean, std.std, device=dev)
        p, xt, yt = t(flat), t(x), t(y)
        m = torch.zeros_like(p); v = torch.zeros_like(p); step = torch.zeros(1, dtype=torch.int32, device=dev)
        loss = torch.zeros(1, device=dev); grad = torch.zeros_like(p)
        opt = eng.make_opt_desc(lr=5e-4, warmup=0.1, decay_steps=1e9)
        eng.pack(p)
        def one():
            eng.loss_grad(p, xt, yt, loss, grad, packed_current=True); eng.clip_adam(p, grad, m, v, step, opt, norm_current=True)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            one(); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s): one()
        torch.cuda.synchronize()
        for _ in range(10): g.replay()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(300): g.replay()
        e1.record(); torch.cuda.synchronize()
        tt = e0.elapsed_time(e1) / 300 * 1e3
        for _ in range(5): eng.log_prob(p, xt, yt, packed_current=True)
        e0.record()
        for _ in range(100): eng.log_prob(p, xt, yt, packed_current=True)
        e1.record(); torch.cuda.synchronize()
        print(f"{name} B={B} occ={occ}: train step {tt:.1f} us ({B/tt:.1f} Msamples/s)  log_prob {e0.elapsed_time(e1)/100*1e3:.1f} us  loss {loss.item():.4f}", flush=True)

# wide config c5 (one step)
cfg = dict(n_in=32, n_cond=128, n_layers=10, layers=[512, 512])
for B in (8192, 65536):
    spec, flat, std, x, y = make_case(cfg, B, seed=1, scale=0.02)
    eng = MafEngine(32, 128, 10, [512, 512], "relu", True, 42, std.shift, std.scale, std.mean, std.std, device=dev)
    p, xt, yt = t(flat), t(x), t(y)
    m = torch.zeros_like(p); v = torch.zeros_like(p); step = torch.zeros(1, dtype=torch.int32, device=dev)
    loss = torch.zeros(1, device=dev); grad = torch.zeros_like(p)
    opt = eng.make_opt_desc(lr=5e-4, warmup=0.1, decay_steps=1e9)
    eng.pack(p)
    def one():
        eng.loss_grad(p, xt, yt, loss, grad, packed_current=True); eng.clip_adam(p, grad, m, v, step, opt, norm_current=True)
    for _ in range(2): one()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): one()
    e1.record(); torch.cuda.synchronize()
    tt = e0.elapsed_time(e1) / 3
    print(f"c5 wide B={B}: train step {tt:.2f} ms ({B/tt/1e3:.2f} Msamples/s, {22609920*B/tt/1e9:.1f} TFLOP/s) loss {loss.item():.4f}", flush=True)
