"""One c5 (wide) training step for profiling (debug tool)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from tests.util import make_case
from jaxili_b200.engine import MafEngine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
cfg = dict(n_in=32, n_cond=128, n_layers=10, layers=[512, 512])
spec, flat, std, x, y = make_case(cfg, B, seed=1, scale=0.02)
dev = torch.device("cuda:0")
t = lambda a: torch.tensor(np.ascontiguousarray(a), device=dev)
eng = MafEngine(32, 128, 10, [512, 512], "relu", True, 42, std.shift, std.scale, std.mean, std.std, device=dev)
p, xt, yt = t(flat), t(x), t(y)
m = torch.zeros_like(p); v = torch.zeros_like(p); step = torch.zeros(1, dtype=torch.int32, device=dev)
loss = torch.zeros(1, device=dev); grad = torch.zeros_like(p)
opt = eng.make_opt_desc(lr=5e-4, warmup=0.1, decay_steps=1e9)
eng.pack(p)
for _ in range(2):
    eng.loss_grad(p, xt, yt, loss, grad, packed_current=True); eng.clip_adam(p, grad, m, v, step, opt, norm_current=True)
torch.cuda.synchronize()
print("ok", loss.item())
