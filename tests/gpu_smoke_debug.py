"""Quick GPU bring-up script (not a pytest file): prints errors per entry point and rough timings."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import maf_numpy as onp
from tests.util import CONFIGS, make_case, rel_err, grad_err
from jaxili_b200.engine import MafEngine, ffma_peak_tflops

def t(a): return torch.tensor(np.ascontiguousarray(a), device="cuda:0")

for name, B in [("c2", 64), ("c2", 4096), ("c1", 128), ("c3", 16384), ("ref_test", 10)]:
    spec, flat, std, x, y = make_case(CONFIGS[name], B, seed=1)
    eng = MafEngine(spec.n_in, spec.n_cond, spec.n_layers, spec.layers, "relu", True, 42, std.shift, std.scale, std.mean, std.std, device=torch.device("cuda:0"))
    print(name, B, eng.query(), flush=True)
    lp = eng.log_prob(t(flat), t(x), t(y)); torch.cuda.synchronize()
    ref = onp.nde_log_prob(spec, flat.astype(np.float64), std, x.astype(np.float64), y.astype(np.float64))
    print("  log_prob err", rel_err(lp.cpu().numpy(), ref), flush=True)
    loss = torch.zeros(1, device="cuda:0"); grad = torch.zeros(spec.n_params, device="cuda:0")
    eng.loss_grad(t(flat), t(x), t(y), loss, grad); torch.cuda.synchronize()
    l_ref, g_ref = onp.nll_loss_and_grad(spec, flat.astype(np.float64), std, x.astype(np.float64), y.astype(np.float64))
    print("  loss err", rel_err(loss.item(), l_ref), "grad err", grad_err(grad.cpu().numpy(), g_ref), flush=True)
    u = np.random.default_rng(0).standard_normal((min(B, 512), spec.n_in)).astype(np.float32)
    out = eng.sample_from_noise(t(flat), t(u), t(y[:1])); torch.cuda.synchronize()
    sref = onp.nde_sample_from_noise(spec, flat.astype(np.float64), std, u.astype(np.float64), y[0])
    print("  sample err", rel_err(out.cpu().numpy(), sref), flush=True)

print("ffma peak TFLOP/s", ffma_peak_tflops(0, 4096))
# timing c2 train step
spec, flat, std, x, y = make_case(CONFIGS["c2"], 4096, seed=1)
eng = MafEngine(10, 10, 5, [50, 50], "relu", True, 42, std.shift, std.scale, std.mean, std.std, device=torch.device("cuda:0"))
p, xt, yt = t(flat), t(x), t(y)
m = torch.zeros_like(p); v = torch.zeros_like(p); step = torch.zeros(1, dtype=torch.int32, device="cuda:0")
loss = torch.zeros(1, device="cuda:0"); grad = torch.zeros_like(p)
opt = eng.make_opt_desc(lr=5e-4, warmup=0.1, decay_steps=1e9)
def one(): 
    eng.loss_grad(p, xt, yt, loss, grad, packed_current=True); eng.clip_adam(p, grad, m, v, step, opt, norm_current=True)
eng.pack(p)
for _ in range(5): one()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200): one()
e1.record(); torch.cuda.synchronize()
print("c2 eager step us", e0.elapsed_time(e1) / 200 * 1e3, "loss", loss.item())
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    one(); 
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        one()
torch.cuda.synchronize()
for _ in range(5): g.replay()
e0.record()
for _ in range(500): g.replay()
e1.record(); torch.cuda.synchronize()
print("c2 graph step us", e0.elapsed_time(e1) / 500 * 1e3, "loss", loss.item())
# kernels separately
for nm, fn in [("loss_grad", lambda: eng.loss_grad(p, xt, yt, loss, grad, packed_current=True)), ("clip_adam", lambda: eng.clip_adam(p, grad, m, v, step, opt, norm_current=True)), ("log_prob", lambda: eng.log_prob(p, xt, yt, packed_current=True))]:
    for _ in range(3): fn()
    e0.record()
    for _ in range(200): fn()
    e1.record(); torch.cuda.synchronize()
    print(nm, "us", e0.elapsed_time(e1) / 200 * 1e3)
# sampling throughput
N = 1 << 20
yo = t(y[:64])
out = torch.empty((N, 10), device="cuda:0")
eng.sample_philox(p, N, yo[:1], seed=1, out=out)
torch.cuda.synchronize()
e0.record()
for _ in range(5): eng.sample_philox(p, N, yo[:1], seed=1, out=out, packed_current=True)
e1.record(); torch.cuda.synchronize()
print("sample 1M rows ms", e0.elapsed_time(e1) / 5, "Msamples/s", N / (e0.elapsed_time(e1) / 5 * 1e-3) / 1e6)
