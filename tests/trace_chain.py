"""Timeline of CTA 0 of the chain kernel: chain warp 0 and gradient warp 0 (debug tool, not a pytest file).
Needs a library built with MAFB_TRACE=1.  usage: python tests/trace_chain.py [c2|c3|c1] [batch] [train|fwd]"""
import sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from tests.util import CONFIGS, make_case
from jaxili_b200.engine import MafEngine
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
mode = sys.argv[3] if len(sys.argv) > 3 else "train"
spec, flat, std, x, y = make_case(CONFIGS[name], B, seed=1)
dev = torch.device("cuda:0")
t = lambda a: torch.tensor(np.ascontiguousarray(a), device=dev)
eng = MafEngine(spec.n_in, spec.n_cond, spec.n_layers, spec.layers, "relu", True, 42, std.shift, std.scale, std.mean, std.std, device=dev)
p, xt, yt = t(flat), t(x), t(y)
loss = torch.zeros(1, device=dev); grad = torch.zeros_like(p)
eng.pack(p)
trace = torch.zeros(1001, dtype=torch.int64, device=dev)
run = (lambda: eng.loss_grad(p, xt, yt, loss, grad, packed_current=True)) if mode == "train" else (lambda: eng.log_prob(p, xt, yt, packed_current=True))
for _ in range(3): run()
eng.lib.mafb200_debug_trace(eng._h, C.c_void_p(trace.data_ptr()))
run()
torch.cuda.synchronize()
tr = trace.cpu().numpy()
def name_of(i):
    if i < 10: return {0: "start", 1: "tile start", 2: "stage landed", 3: "transposed", 4: "finalize done"}.get(i, str(i))
    if 90 <= i < 100: return f"W fwd acquired k={i-90}"
    if i == 99: return "end"
    if 100 <= i < 190: return f"  fwd k={(i-100)//10} lin {(i-100)%10} done"
    if 190 <= i < 200: return f"W bwd acquired k={i-190}"
    if 200 <= i < 300: return f"  bwd k={(i-200)//10} lin {(i-200)%10}: at empty-sync"
    if 300 <= i < 400: return f"  bwd k={(i-300)//10} lin {(i-300)%10}: past empty-sync"
    if 400 <= i < 500: return f"  bwd k={(i-400)//10} lin {(i-400)%10}: dX done"
    if 500 <= i < 600: return f"grad k={(i-500)//10} lin {(i-500)%10}: past full-sync"
    if 600 <= i < 700: return f"grad k={(i-600)//10} lin {(i-600)%10}: dW done"
    if 800 <= i < 900: return f"grad k=3 lin {i-800}: unit done"
    if 700 <= i < 800: return f"     k=1 lin {(i-700)//10}: " + {1: "setup done", 2: "mac done", 3: "reduce g0 done", 4: "reduce g1 done"}.get((i-700)%10, "?")
    return str(i)
ev = []
for base, who in ((0, "chain"), (500, "grad ")):
    n = int(tr[base + 498])
    for i in range(n):
        ev.append((int(tr[base + 2 * i + 1]), who, int(tr[base + 2 * i])))
ev.sort()
t0 = ev[0][0]
last = {"chain": t0, "grad ": t0}
for clk, who, i in ev:
    print(f"{who} t={clk - t0:8d} +{clk - last[who]:6d}  {name_of(i)}")
    last[who] = clk
