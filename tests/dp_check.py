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
# the streamed epoch (native pipeline, fused tail with the peer exchange inside) continues the same trajectory
lo, hi = parallel.shard_rows(B, rank, world)
loader = [[theta[i * B:(i + 1) * B][lo:hi], x[i * B:(i + 1) * B][lo:hi]] for i in range(4)]
# FP32 envelope: the same restatement run in float32 from the current state (oracle (2))
ref32 = ref.astype(np.float32)
st32 = onp.AdamState(ref.size, np.float32)
st32.m, st32.v, st32.count = st.m.astype(np.float32), st.v.astype(np.float32), st.count
for ep in range(2):
    met = tr.train_epoch(loader)
    l_mean = 0.0
    for i in range(4):
        gb = slice(i * B, (i + 1) * B)
        l_ref, g_ref = onp.nll_loss_and_grad(spec, ref, std, theta[gb].astype(np.float64), x[gb].astype(np.float64))
        l_mean += l_ref / 4
        _, g32 = onp.nll_loss_and_grad(spec, ref32, std, theta[gb], x[gb])
        ref32, _, _ = onp.clip_adam_step(ref32, g32.astype(np.float32), st32, 1e-3, 0.1, float(int(100 // 2 * 4)), clip=0.05)
        ref, _, _ = onp.clip_adam_step(ref, g_ref, st, 1e-3, 0.1, float(int(100 // 2 * 4)), clip=0.05)
    assert abs(met["train/loss"] - l_mean) < 2e-5 * max(1, abs(l_mean)), (met["train/loss"], l_mean)
streamed = tr._pipeline is not None
flat = tr.state.params.flat
assert int(tr.state.step.item()) == 12
dev = np.abs(flat.cpu().numpy() - ref)
# Adam normalises by sqrt(v): entries whose gradient is at FP32 noise level take lr-sized steps of either sign, so
# over 12 steps the bound is the FP32 envelope (float32 restatement vs float64 restatement), not a constant
env = np.abs(ref32.astype(np.float64) - ref)
assert dev.max() < 5e-3 and np.quantile(dev, 0.99) < 5e-6 + 3 * np.quantile(env, 0.99), (dev.max(), np.quantile(dev, 0.99), np.quantile(env, 0.99), env.max())
gathered = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
assert all(torch.equal(g, gathered[0]) for g in gathered), "replicas diverged"
if rank == 0:
    print(f"dp_check ok: world={world}, streamed_epoch={streamed}, peer_error={tr.model.engine().dp_error() if world > 1 and streamed else None}, replicas bit-identical, max |p - oracle| = {np.max(np.abs(flat.cpu().numpy() - ref)):.2e}")
dist.destroy_process_group()
