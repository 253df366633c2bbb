"""Phase timeline of CTA 0 (debug tool, not a pytest file)."""
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
for _ in range(3):
    if mode == "train": eng.loss_grad(p, xt, yt, loss, grad, packed_current=True)
    else: eng.log_prob(p, xt, yt, packed_current=True)
eng.lib.mafb200_debug_trace(eng._h, C.c_void_p(trace.data_ptr()))
if mode == "train": eng.loss_grad(p, xt, yt, loss, grad, packed_current=True)
else: eng.log_prob(p, xt, yt, packed_current=True)
torch.cuda.synchronize()
tr = trace.cpu().numpy(); n = int(tr[1000])
ids = tr[0:2*n:2]; clk = tr[1:2*n:2]
names = {50:"   acq: before wait",51:"   acq: after wait",40:"  X start",41:"  X done",42:"  B start",43:"  B done",30:" mm start",31:" mm done",32:" scatter done",33:" epilogue done",0:"start",1:"stage ready",2:"transpose done",3:"W fwd acquired",4:"W bwd acquired",5:"bwd elementwise",6:"finalize",19:"fwd affine",99:"end"}
prev = clk[0]
for i in range(n):
    nm = names.get(int(ids[i]), ("fwd lin %d" % (ids[i]-10)) if 10 <= ids[i] < 19 else ("bwd lin %d" % (ids[i]-20)))
    print(f"{nm:18s} +{clk[i]-prev:7d}  t={clk[i]-clk[0]:8d}")
    prev = clk[i]
