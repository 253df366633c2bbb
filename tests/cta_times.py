"""Per-CTA entry / exit times (%globaltimer) of the chain kernel (debug tool, not a pytest file).
Needs a library built with MAFB_CTA_TIMES=1."""
import sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from tests.util import CONFIGS, make_case
from jaxili_b200.engine import MafEngine
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
spec, flat, std, x, y = make_case(CONFIGS[name], B, seed=1)
dev = torch.device("cuda:0")
t = lambda a: torch.tensor(np.ascontiguousarray(a), device=dev)
eng = MafEngine(spec.n_in, spec.n_cond, spec.n_layers, spec.layers, "relu", True, 42, std.shift, std.scale, std.mean, std.std, device=dev)
p, xt, yt = t(flat), t(x), t(y)
loss = torch.zeros(1, device=dev); grad = torch.zeros_like(p)
eng.pack(p)
trace = torch.zeros(1001, dtype=torch.int64, device=dev)
for _ in range(5): eng.loss_grad(p, xt, yt, loss, grad, packed_current=True)
eng.lib.mafb200_debug_trace(eng._h, C.c_void_p(trace.data_ptr()))
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
eng.loss_grad(p, xt, yt, loss, grad, packed_current=True)
e1.record()
torch.cuda.synchronize()
tr = trace.cpu().numpy()
n = 147
t_in = tr[0:2 * n:2]; t_grad = tr[1:2 * n:2]; t_chain = tr[400:400 + n]
base = t_in.min()
print("events (flow + reduce kernels): %.1f us" % (e0.elapsed_time(e1) * 1e3))
print("CTA entry  : min %.1f max %.1f us" % (0, (t_in.max() - base) / 1e3))
print("chain exit : min %.1f mean %.1f max %.1f us" % ((t_chain.min() - base) / 1e3, (t_chain.mean() - base) / 1e3, (t_chain.max() - base) / 1e3))
print("grad exit  : min %.1f mean %.1f max %.1f us" % ((t_grad.min() - base) / 1e3, (t_grad.mean() - base) / 1e3, (t_grad.max() - base) / 1e3))
d = (t_grad - t_in) / 1e3
print("per-CTA duration: min %.1f mean %.1f max %.1f us" % (d.min(), d.mean(), d.max()))
