"""Profiling driver (not a pytest file): a few sampling calls of the c4 model."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from tests.util import CONFIGS, make_case
from jaxili_b200.engine import MafEngine
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
spec, flat, std, x, y = make_case(CONFIGS["c2"], 4, seed=1)
dev = torch.device("cuda:0")
t = lambda a: torch.tensor(np.ascontiguousarray(a), device=dev)
eng = MafEngine(spec.n_in, spec.n_cond, spec.n_layers, spec.layers, "relu", True, 42, std.shift, std.scale, std.mean, std.std, device=dev)
p, yt = t(flat), t(y)
for i in range(4):
    out = eng.sample_philox(p, N, yt[:1], seed=1, packed_current=i > 0)
torch.cuda.synchronize()
print("ok", out.mean().item())
