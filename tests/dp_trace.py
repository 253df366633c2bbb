"""Where the data-parallel step tail spends its time (debug tool; run under torchrun, N >= 1)."""
import os, sys, ctypes as C
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxili_b200 import nn, parallel
from jaxili_b200.loss import loss_nll_npe
from jaxili_b200.model import ConditionalMAF, Identity, NDE_w_Standardization, ScalarAffine, Sequential, Standardizer
from jaxili_b200.train import TrainerModule
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.default_rng(rank)
B = 4096
theta = rng.uniform(-1, 1, size=(B, 10)).astype(np.float32)
x = (theta + np.sqrt(0.1) * rng.standard_normal((B, 10))).astype(np.float32)
maf = ConditionalMAF(10, 10, 5, [50, 50], activation=nn.relu, use_reverse=True, seed=42)
hp = {"nde": maf, "embedding_net": Sequential([Standardizer(x.mean(0), x.std(0)), Identity()]),
      "transformation": ScalarAffine(shift=theta.mean(0), scale=theta.std(0))}
tr = TrainerModule(NDE_w_Standardization, hp, {"lr": 1e-3, "warmup": 0.1}, loss_nll_npe,
                   (np.zeros((1, 10), np.float32), np.zeros((1, 10), np.float32)),
                   logger_params={"base_log_dir": f"/tmp/dp_trace_{rank}"}, enable_progress_bar=False)
tr.init_optimizer(100, 1000)
eng = tr.model.engine()
dev = torch.device("cuda", local)
td, xd = torch.tensor(theta, device=dev), torch.tensor(x, device=dev)
trace = torch.zeros(1001, dtype=torch.int64, device=dev)
eng.lib.mafb200_debug_trace(eng._h, C.c_void_p(trace.data_ptr()))     # before capture: the graph bakes the pointer
step = tr.make_graph_step(td, xd, B)
for _ in range(20):
    step()
torch.cuda.synchronize(); dist.barrier()
# steady-state graph replay: the stamps of the last step survive; sample a few by syncing every 50 steps
g_rows = []
for rep in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        step()
    e1.record(); torch.cuda.synchronize()
    t = trace[900:908].cpu().numpy()
    g_rows.append([e0.elapsed_time(e1) * 1e3 / 50] + [(t[k] - t[0]) / 1e3 for k in range(1, 7)] + [t[7] / 1e3])
    if os.environ.get("MAFB_CTA_TIMES") and rank == 0 and rep == 5:
        # library built with MAFB_CTA_TIMES=1: the last flow launch's per-CTA stamps, relative to the tail's start
        tr_ = trace.cpu().numpy()
        n = 147
        t_in, t_grad, t_chain = tr_[0:2 * n:2], tr_[1:2 * n:2], tr_[400:400 + n]
        f = lambda a: "min %.2f max %.2f" % ((a.min() - t[0]) / 1e3, (a.max() - t[0]) / 1e3)
        t_tail = tr_[600:600 + 181]
        print("tail block exit", f(t_tail), "| previous tail's exits are overwritten: compare with column 'end'")
        print("flow CTA entry", f(t_in), "| chain exit", f(t_chain), "| grad exit", f(t_grad), " (us, tail start = 0)")
if rank == 0:
    print("GRAPH mode: step_us | stage1 done, flags raised, peers seen, summed, barrier B, end | previous tail end -> tail start")
    for r in g_rows:
        print("G " + " ".join(f"{v:8.2f}" for v in r))
dist.barrier()
# the graph baked a null trace pointer: run eager steps for the stamps
st = tr.state
flat = st.params.flat
P = flat.numel()
grad, loss = tr._gbuf[:P], tr._gbuf[-4:-3]
peer = world > 1
rows = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(12):
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    eng.train_step(flat, td, xd, loss, grad, st.opt_state["m"], st.opt_state["v"], st.step, st.tx,
                   inv_global_b=1.0 / (B * world), packed_current=True, dp_slot=peer)
    e1.record(); torch.cuda.synchronize()
    t = trace[900:907].cpu().numpy()
    rows.append([e0.elapsed_time(e1) * 1e3] + [(t[k] - t[0]) / 1e3 for k in range(1, 7)])
if rank == 0:
    print("step_us | tail: stage1 done, barrier A, peers seen, summed, barrier B, end  (us since tail start)")
    for r in rows[2:]:
        print(" ".join(f"{v:8.2f}" for v in r))
dist.destroy_process_group()
