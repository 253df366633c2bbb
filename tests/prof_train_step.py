"""Profiling driver (not a pytest file): a few fused training steps (flow kernel + step tail) of one config.
usage: python tests/prof_train_step.py [c2|c3|c1] [batch] [steps]"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from tests.util import CONFIGS, make_case
from jaxili_b200.engine import MafEngine

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
spec, flat, std, x, y = make_case(CONFIGS[name], B, seed=1)
dev = torch.device("cuda:0")
t = lambda a: torch.tensor(np.ascontiguousarray(a), device=dev)
eng = MafEngine(spec.n_in, spec.n_cond, spec.n_layers, spec.layers, "relu", True, 42, std.shift, std.scale, std.mean, std.std, device=dev)
p, xt, yt = t(flat), t(x), t(y)
m = torch.zeros_like(p); v = torch.zeros_like(p); step = torch.zeros(1, dtype=torch.int32, device=dev)
loss = torch.zeros(1, device=dev); grad = torch.zeros_like(p)
opt = eng.make_opt_desc(lr=5e-4, warmup=0.1, decay_steps=1e9)
eng.pack(p)
for _ in range(steps):
    eng.train_step(p, xt, yt, loss, grad, m, v, step, opt, packed_current=True)
torch.cuda.synchronize()
print("ok", loss.item(), int(step.item()))
