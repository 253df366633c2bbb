"""Multi-GPU data-parallel check (run with torchrun, one rank per GPU; not a pytest file):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dp_check.py
Every rank trains on its shard through TrainerModule.train_step (NCCL all-reduce of the flat
gradient); the replicas must stay bit-identical and match the oracle on the global batch."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxili_b200 import nn, parallel
from jaxili_b200.loss import loss_nll_npe
from jaxili_b200.model import ConditionalMAF, Identity, NDE_w_Standardization, ScalarAffine, Sequential, Standardizer
from jaxili_b200.train import TrainerModule
from oracle import maf_numpy as onp

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.default_rng(0)
B = 256 * world
theta = rng.uniform(-1, 1, size=(4 * B, 10)).astype(np.float32)
x = (theta + np.sqrt(0.1) * rng.standard_normal((4 * B, 10))).astype(np.float32)
sh, sc, me, sd = theta.mean(0), theta.std(0), x.mean(0), x.std(0)
maf = ConditionalMAF(10, 10, 5, [50, 50], activation=nn.relu, use_reverse=True, seed=42)
hp = {"nde": maf, "embedding_net": Sequential([Standardizer(me, sd), Identity()]), "transformation": ScalarAffine(shift=sh, scale=sc)}
tr = TrainerModule(NDE_w_Standardization, hp, {"lr": 1e-3, "warmup": 0.1, "gradient_clip": 0.05}, loss_nll_npe,
                   (np.zeros((1, 10), np.float32), np.zeros((1, 10), np.float32)),
                   logger_params={"base_log_dir": f"/tmp/dp_check_{rank}"}, enable_progress_bar=False)
tr.init_optimizer(100, 4)
spec = onp.MafSpec(10, 10, 5, [50, 50], seed=42)
std = onp.Standardization(sh, sc, me, sd)
ref = tr.state.params.flat.cpu().numpy().astype(np.float64)
st = onp.AdamState(ref.size, np.float64)
for i in range(4):
    gb = slice(i * B, (i + 1) * B)
    lo, hi = parallel.shard_rows(B, rank, world)
    tr.state, m = tr.train_step(tr.state, [theta[gb][lo:hi], x[gb][lo:hi]])
    l_ref, g_ref = onp.nll_loss_and_grad(spec, ref, std, theta[gb].astype(np.float64), x[gb].astype(np.float64))
    assert abs(m["loss"].item() - l_ref) < 2e-5 * max(1, abs(l_ref)), (m["loss"].item(), l_ref)
    ref, _, _ = onp.clip_adam_step(ref, g_ref, st, 1e-3, 0.1, float(int(100 // 2 * 4)), clip=0.05)
flat = tr.state.params.flat
assert np.max(np.abs(flat.cpu().numpy() - ref)) < 5e-6
gathered = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
assert all(torch.equal(g, gathered[0]) for g in gathered), "replicas diverged"
if rank == 0:
    print(f"dp_check ok: world={world}, replicas bit-identical, max |p - oracle| = {np.max(np.abs(flat.cpu().numpy() - ref)):.2e}")
dist.destroy_process_group()
