"""Leapfrog throughput of the vectorised-chains HMC (debug tool): chains x steps per second."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from tests.util import CONFIGS, make_case
from jaxili_b200.engine import MafEngine
spec, flat, std, x, y = make_case(CONFIGS["c3"], 8, seed=1)
dev = torch.device("cuda:0")
eng = MafEngine(spec.n_in, spec.n_cond, spec.n_layers, spec.layers, "relu", True, 42, std.shift, std.scale, std.mean, std.std, device=dev)
p = torch.tensor(flat, device=dev)
C = spec.n_cond
for n in (1, 256, 4096, 65536):
    f32 = dict(dtype=torch.float32, device=dev)
    S = {"eps": torch.full((1,), 0.01, **f32), "inv_mass": torch.ones(C, **f32),
         "prior_kind": torch.ones(C, dtype=torch.int32, device=dev), "lo": torch.full((C,), -3.0, **f32),
         "hi": torch.full((C,), 3.0, **f32), "z": torch.zeros(n, C, **f32), "p": torch.randn(n, C, **f32),
         "g": torch.zeros(n, C, **f32), "logdens": torch.zeros(n, **f32), "theta": torch.zeros(n, C, **f32),
         "scratch": torch.zeros(n * (C + 1), **f32)}
    xr = torch.tensor(x[:1], device=dev).expand(n, -1).contiguous()
    eng.hmc_trajectory(p, xr, S, 0)
    eng.hmc_trajectory(p, xr, S, 20, packed_current=True)
    torch.cuda.synchronize()
    steps = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.hmc_trajectory(p, xr, S, steps, packed_current=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"chains {n:6d}: {1e3 * ms / steps:8.1f} us per leapfrog step of all chains, {n * steps / (ms * 1e-3) / 1e6:9.2f} M gradient evaluations/s, finite={bool(torch.isfinite(S['logdens']).all())}")
