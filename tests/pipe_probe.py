"""Where does the streamed step's time go?  (debug tool, not a test)"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench as B
from jaxili_b200 import nn
from jaxili_b200.loss import loss_nll_npe
from jaxili_b200.model import ConditionalMAF, Identity, NDE_w_Standardization, ScalarAffine, Sequential, Standardizer
from jaxili_b200.train import TrainerModule, HostBatchPipeline

cfg = dict(n_in=10, n_cond=10, n_layers=5, layers=[50, 50])
rng = np.random.default_rng(0)
rows = 4096
theta = rng.uniform(-1, 1, (16 * rows, 10)).astype(np.float32)
x = (theta + 0.3 * rng.standard_normal(theta.shape)).astype(np.float32)
maf = ConditionalMAF(activation=nn.relu, use_reverse=True, seed=42, **cfg)
hp = {"nde": maf, "embedding_net": Sequential([Standardizer(x.mean(0), x.std(0)), Identity()]),
      "transformation": ScalarAffine(shift=theta.mean(0), scale=theta.std(0))}
tr = TrainerModule(NDE_w_Standardization, hp, {"lr": 5e-4, "warmup": 0.1}, loss_nll_npe,
                   (np.zeros((1, 10), np.float32), np.zeros((1, 10), np.float32)),
                   logger_params={"base_log_dir": "/tmp/pp"}, enable_progress_bar=False)
tr.init_optimizer(100, 1000)
hb = [[torch.from_numpy(theta[i * rows:(i + 1) * rows]).pin_memory(), torch.from_numpy(x[i * rows:(i + 1) * rows]).pin_memory()] for i in range(16)]
pipe = tr.host_pipeline(rows)
N = 1000
for i in range(50):
    pipe.submit(hb[i % 16])
pipe.drain()
t0 = time.perf_counter()
for i in range(N):
    pipe.submit(hb[i % 16])
t1 = time.perf_counter()
pipe.drain()
t2 = time.perf_counter()
print(f"pipe.submit host {1e6 * (t1 - t0) / N:6.1f} us/submit   total {1e6 * (t2 - t0) / N:6.1f} us/step")
for n in (300, 300, 1000, 3000):
    loader = [hb[i % 16] for i in range(n)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tr.train_epoch(loader)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"train_epoch {n} steps: {1e6 * (t1 - t0) / n:6.1f} us/step")
