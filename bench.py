#!/usr/bin/env python
"""bench.py -- the hot-path benchmark (driver contract).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2|c3|c1|c4]

A "step" is one pass of the hot path over one batch of synthetic simulations:
  training workloads (c1, c2, c3): fused forward + backward + gradient reduction + global-norm clip
      + warmup-cosine Adam update on one batch (jaxili/train.py:330-335 with the chain of 246-300);
  sampling workload (c4): one posterior.sample call (jaxili/posterior/direct_posterior.py:45-73).
Default workload = BASELINE.json configs[1]: NPE gaussian-linear, n_in=10, n_cond=10, 5 x [50,50],
batch 4096 per GPU.  With N GPUs (torchrun, one rank per GPU) every rank trains on its own batch
shard (weak scaling, global batch = N x 4096) and the flat gradient is all-reduced once per step.

One JSON line is printed by rank 0:
  value     whole-job samples/s with the dataset resident in HBM (CUDA-graph replay, device-side
            batch selection, CUDA events on the launch stream, max over ranks)
  e2e       the same metric through the reference-facing API (TrainerModule.train_step) with HOST
            batches in pinned memory: H2D copy of every batch and D2H read of the loss inside the
            timed region
  roofline  FP32-FFMA roofline of the fused flow kernel: algorithmic FLOPs / CUDA-event time of the
            kernel alone, against the FFMA peak measured live on this GPU (MEASURED_PEAKS.json has
            no FP32 figure; its HBM number bounds nothing here: 80 B/sample)
  cpu_baseline  the CPU oracle port (torch, all host threads) on a bounded sample of the same workload

``--impl reference`` times the reference's CPU path instead: jaxili itself (JAX) cannot be installed
in this image, so the arm runs the line-by-line CPU restatement under oracle/ (kind "port").
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
# Training steps captured per CUDA-graph launch of the resident-data loop (narrow workloads): the largest of these
# that divides --steps, unless MAFB200_BENCH_GRAPH_STEPS says otherwise.  Measured on B200 (c2): 54.4 -> 52.5 us per
# step at N=1, 66.0 -> 61.1 us at N=8 with 25 steps per launch -- the boundary between two graph launches costs
# ~2 us more than a kernel-to-kernel edge inside a graph, and the ranks' launch jitter ends up in the exchange.
SPL_CHOICES = (25, 20, 10, 8, 5, 4, 2, 1)


def graph_steps(K):
    e = os.environ.get("MAFB200_BENCH_GRAPH_STEPS")
    if e:
        return max(1, int(e))
    return next(s for s in SPL_CHOICES if K % s == 0)
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (model config, per-GPU batch, description)
    "c1": (dict(n_in=2, n_cond=2, n_layers=5, layers=[50, 50]), 128,
           "NPE two-moons-shaped D=2,C=2, 5x[50,50] ReLU reverse, batch 128"),
    "c2": (dict(n_in=10, n_cond=10, n_layers=5, layers=[50, 50]), 4096,
           "NPE gaussian-linear D=10,C=10, 5x[50,50] ReLU reverse, batch 4096 per GPU"),
    "c3": (dict(n_in=8, n_cond=5, n_layers=5, layers=[50, 50]), 16384,
           "NLE SLCP-shaped D=8 (x), C=5 (theta), 5x[50,50], batch 16384 per GPU"),
    "c4": (dict(n_in=10, n_cond=10, n_layers=5, layers=[50, 50]), 1 << 20,
           "posterior sampling D=10,C=10, 5x[50,50]: 2^20 samples per observation per step"),
    "c5": (dict(n_in=32, n_cond=128, n_layers=10, layers=[512, 512]), 65536,
           "wide NPE D=32,C=128, 10x[512,512] ReLU reverse, batch 65536 per GPU, tcgen05 TF32 GEMM core"),
}


def macs_per_pass(cfg):
    dims = [cfg["n_in"] + cfg["n_cond"], *cfg["layers"], 2 * cfg["n_in"]]
    return sum(a * b for a, b in zip(dims[:-1], dims[1:]))


def train_flops_per_sample(cfg):
    """dense, 2 FLOP/MAC, masks not discounted: fwd + dX + dW = 6*M*L."""
    return 6 * macs_per_pass(cfg) * cfg["n_layers"]


def sample_flops_per_sample(cfg):
    """the reference's D full MADE passes per layer: 2*M*L*D."""
    return 2 * macs_per_pass(cfg) * cfg["n_layers"] * cfg["n_in"]


def synth_data(name, cfg, n_rows, seed):
    """Synthetic simulator draws of the named shape, FP32."""
    rng = np.random.default_rng(seed)
    D, C = cfg["n_in"], cfg["n_cond"]
    if name == "c5":              # theta ~ U(-1,1)^32, x = A theta + 0.3 eps, fixed A ~ N(0, 1/32)
        theta = rng.uniform(-1, 1, size=(n_rows, D)).astype(np.float32)
        A = np.random.default_rng(5).standard_normal((C, D)).astype(np.float32) / np.sqrt(D)
        x = (theta @ A.T + 0.3 * rng.standard_normal((n_rows, C))).astype(np.float32)
        return theta, x
    if name in ("c2", "c4"):      # gaussian-linear: theta ~ U(-1,1)^10, x = theta + sqrt(0.1) eps
        theta = rng.uniform(-1, 1, size=(n_rows, D)).astype(np.float32)
        x = (theta + np.sqrt(0.1) * rng.standard_normal((n_rows, C))).astype(np.float32)
        return theta, x
    if name == "c1":              # two moons
        theta = rng.uniform(-1, 1, size=(n_rows, 2)).astype(np.float32)
        a = rng.uniform(-np.pi / 2, np.pi / 2, n_rows)
        r = rng.normal(0.1, 0.01, n_rows)
        x = np.stack([r * np.cos(a) + 0.25 - np.abs(theta[:, 0] + theta[:, 1]) / np.sqrt(2),
                      r * np.sin(a) + (-theta[:, 0] + theta[:, 1]) / np.sqrt(2)], axis=1).astype(np.float32)
        return theta, x
    if name == "c3":              # SLCP-shaped: theta in R^5, x in R^8 (four 2-D Gaussian draws)
        th = rng.uniform(-3, 3, size=(n_rows, 5))
        s1, s2, rho = th[:, 2] ** 2, th[:, 3] ** 2, np.tanh(th[:, 4])
        z = rng.standard_normal((n_rows, 4, 2))
        x0 = th[:, None, 0] + s1[:, None] * z[:, :, 0]
        x1 = th[:, None, 1] + s2[:, None] * (rho[:, None] * z[:, :, 0] + np.sqrt(1 - rho[:, None] ** 2) * z[:, :, 1])
        x = np.stack([x0, x1], axis=2).reshape(n_rows, 8)
        return th.astype(np.float32), x.astype(np.float32)
    raise ValueError(name)


def density_and_cond(name, theta, x):
    """(density variable, conditioner): NPE learns p(theta|x), NLE p(x|theta)."""
    return (x, theta) if name == "c3" else (theta, x)


def zscore(a):
    return a.mean(0, dtype=np.float32), a.std(0, dtype=np.float32)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region.

    nvidia-smi needs ~0.2 s to come up, so the sampler is launched before the warm-up (``launch``), the timed
    region is bracketed with ``start`` / ``mark_end`` and ``stop`` reports the samples that arrived inside the
    bracket.  A timed region shorter than the 20 ms sampling period may catch none: then the samples taken while
    the GPU was under load during the rest of the run (warm-up, end-to-end and roofline legs) are reported and
    ``window`` says so."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.lines, self.proc = gpu_index, [], None
        self.t0 = self.t1 = None

    def launch(self):
        if self.proc is not None:
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def start(self):
        self.launch()
        self.t0 = time.perf_counter()

    def mark_end(self):
        if self.t1 is None:
            self.t1 = time.perf_counter()

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self):
        self.mark_end()
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()

        def parse(lines):
            sm, mx, reasons = [], [], set()
            for _, ln in lines:
                f = [v.strip() for v in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            return sm, mx, reasons

        inside = [l for l in self.lines if self.t0 is not None and self.t0 <= l[0] <= self.t1 + 0.02]
        sm, mx, reasons = parse(inside)
        window = "timed region"
        if len(sm) < 2:
            sm, mx, reasons = parse(self.lines)
            window = "whole bench run under load (timed region shorter than the sampling period)"
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


# --------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_train_throughput(name, cfg, batch, budget_s, threads, steps=None, warmup=3):
    """Train steps of the CPU restatement (oracle/maf_torch.py) on `threads` host threads: `warmup` untimed
    steps, then `steps` timed ones (all that fit `budget_s` seconds if None or too many)."""
    import torch
    from oracle import maf_numpy as onp
    from oracle.maf_torch import TorchMaf, TorchTrainer
    torch.set_num_threads(threads)
    theta, x = synth_data(name, cfg, batch * 4, seed=11)
    dv, cv = density_and_cond(name, theta, x)
    spec = onp.MafSpec(activation="relu", use_reverse=True, seed=42, **cfg)
    flat = spec.init_params(np.random.default_rng(42), 0.01, np.float32)
    std = onp.Standardization(*zscore(dv), *zscore(cv))
    tr = TorchTrainer(TorchMaf(spec, std), flat, lr=5e-4, warmup=0.1, decay_steps=1e9)
    xs = [torch.tensor(dv[i * batch:(i + 1) * batch]) for i in range(4)]
    ys = [torch.tensor(cv[i * batch:(i + 1) * batch]) for i in range(4)]
    for i in range(max(1, warmup)):
        tr.train_step(xs[i % 4], ys[i % 4])
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < budget_s and (steps is None or n < steps):
        tr.train_step(xs[n % 4], ys[n % 4])
        n += 1
    dt = time.perf_counter() - t0
    return n * batch / dt, n, dt


def cpu_sample_throughput(cfg, rows, threads):
    import torch
    from oracle import maf_numpy as onp
    from oracle.maf_torch import TorchMaf
    torch.set_num_threads(threads)
    spec = onp.MafSpec(activation="relu", use_reverse=True, seed=42, **cfg)
    flat = spec.init_params(np.random.default_rng(42), 0.01, np.float32)
    std = onp.Standardization.identity(cfg["n_in"], cfg["n_cond"])
    tm = TorchMaf(spec, std)
    u = np.random.default_rng(1).standard_normal((rows, cfg["n_in"])).astype(np.float32)
    y = np.zeros(cfg["n_cond"], np.float32)
    tm.sample_from_noise(flat, u[:1024], y)
    t0 = time.perf_counter()
    tm.sample_from_noise(flat, u, y)
    dt = time.perf_counter() - t0
    return rows / dt, dt


def run_reference_arm(args, name, cfg, batch):
    """--impl reference: the reference's own CPU implementation of the path.  jaxili (JAX) is not
    installable here, so this is the CPU restatement (oracle/), on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    t_all = time.perf_counter()
    if name == "c4":
        rows = 20000
        vals = []
        for _ in range(max(1, min(args.steps, 3))):
            v, dt = cpu_sample_throughput(cfg, rows, threads)
            vals.append(v)
        value = float(np.mean(vals))
        sample = f"{rows} samples per call, 1 observation, {len(vals)} calls"
        metric, steps = "posterior samples/sec", len(vals)
        ms = rows / value * 1e3
    else:
        # bounded: each step is one full batch; W warm-up steps, then K timed steps -- capped at the steps that
        # fit 20 s of CPU work so that the arm ends within a minute whatever K is
        budget = 20.0
        warm = max(1, min(args.warmup, 20))
        value, n, dt = cpu_train_throughput(name, cfg, batch, budget, threads, steps=args.steps, warmup=warm)
        sample = f"{n} train steps of batch {batch} in {dt:.1f} s" + \
                 (f" ({args.steps} requested, capped by the {budget:.0f} s budget)" if n < args.steps else "")
        metric, steps = "MAF train samples/sec (fwd+bwd+clip+Adam)", n
        ms = batch / value * 1e3
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": (max(1, min(args.warmup, 20)) if name != "c4" else 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(name, batch, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "run": {"batch_timed": batch, "ranks_timed": 1},
        "note": "jaxili (JAX/flax/optax) is not installable in this image; CPU restatement under oracle/ timed "
                f"on {threads} host threads, one batch of the per-GPU size per step (rank 0 only); "
                f"wall {time.perf_counter() - t_all:.1f} s",
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU arm
def workload_config(name, batch, world):
    """The keys both arms print under "config": what the workload IS (identical in the GPU and the reference
    arm); how a run executed it goes under "run"."""
    cfg, _, desc = WORKLOADS[name]
    if name == "c4":
        return {"workload": desc, "observations": 64, "samples_per_step": batch, "parallelism": f"obs-shard x{world}"}
    return {"workload": desc, "per_gpu_batch": batch, "global_batch": batch * world,
            "optimizer": "clip_by_global_norm(5) + adam(warmup_cosine(5e-4))", "parallelism": f"dp{world}"}


def run_workload(ctx, name, batch, K, W, full=True):
    """One workload on this rank's GPU (all ranks call it together).  full=False: the secondary records of the
    default run (resident value + end-to-end + roofline on fewer steps, no blocking-per-step leg)."""
    import torch
    import torch.distributed as dist
    from jaxili_b200 import nn as bnn
    from jaxili_b200.engine import fp32_peak_tflops, tf32_peak_tflops
    from jaxili_b200.loss import loss_nll_nle, loss_nll_npe
    from jaxili_b200.model import ConditionalMAF, Identity, NDE_w_Standardization, ScalarAffine, Sequential, Standardizer
    from jaxili_b200.train import TrainerModule
    world, rank, local, dev = ctx["world"], ctx["rank"], ctx["local"], ctx["dev"]
    barrier_sync, max_over_ranks = ctx["barrier_sync"], ctx["max_over_ranks"]
    cfg, _, desc = WORKLOADS[name]
    clocks = ClockSampler(local)
    clocks.launch()
    out = {}

    if name == "c4":
        # ------------------------------------------------------------- posterior sampling
        n_obs_total = 64
        rows = batch                                   # 2^20 samples per observation per step
        theta, x = synth_data("c4", cfg, 4096, seed=3)
        maf = ConditionalMAF(activation=bnn.relu, use_reverse=True, seed=42, device=dev, **cfg)
        model = NDE_w_Standardization(maf, Sequential([Standardizer(*zscore(x)), Identity()]),
                                      ScalarAffine(*zscore(theta)))
        variables = model.init(42)
        eng = model.engine()
        flat = model.flat_params(variables)
        obs = torch.tensor(x[:n_obs_total], device=dev)
        my_obs = list(range(rank, n_obs_total, world))  # observations are independent: shard, no collective
        outbuf = torch.empty((rows, cfg["n_in"]), device=dev)
        eng.sample_philox(flat, rows, obs[0:1], seed=4, out=outbuf)            # packs the sampler image
        for i in range(W):
            eng.sample_philox(flat, rows, obs[my_obs[i % len(my_obs)]][None], seed=4, row_offset=i * rows,
                              out=outbuf, packed_current=True)
        barrier_sync()
        clocks.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            o = my_obs[i % len(my_obs)]
            eng.sample_philox(flat, rows, obs[o:o + 1], seed=4, row_offset=(o << 20), out=outbuf, packed_current=True)
        e1.record()
        barrier_sync()
        clocks.mark_end()
        ms = max_over_ranks(e0.elapsed_time(e1))
        value = K * rows * world / (ms * 1e-3)
        # e2e: DirectPosterior.sample_to_host -- host observation in, samples delivered to pinned host memory
        # (chunked: the device->host copy of chunk k overlaps the inversion kernel of chunk k+1)
        from jaxili_b200.posterior import DirectPosterior
        from jaxili_b200.train import TrainState
        post = DirectPosterior(model, TrainState(step=None, apply_fn=None, params=variables["params"], tx=None,
                                                 opt_state=None), x=None)
        host_out = torch.empty((rows, cfg["n_in"]), dtype=torch.float32).pin_memory()
        Ke = max(3, min(K, 20))
        post.sample_to_host(rows, key=4, x=x[0:1], out=host_out)
        barrier_sync()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        t0.record()
        for i in range(Ke):
            o = my_obs[i % len(my_obs)]
            post.sample_to_host(rows, key=4 + i, x=x[o:o + 1], out=host_out)     # returns with the samples on the host
        t1.record()
        barrier_sync()
        ms_e = max_over_ranks(max(t0.elapsed_time(t1), (time.perf_counter() - w0) * 1e3))
        clk = clocks.stop()
        peak, peak_parts = fp32_peak_tflops(local)
        fl = sample_flops_per_sample(cfg)
        achieved = fl * rows / (ms / K * 1e-3) / 1e12
        out = {
            "metric": "posterior samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(name, batch, world),
            "run": {"l2": "outputs (42 MB/step) stream to HBM; compute-bound; Philox noise drawn in-kernel",
                    "observations_per_rank": len(my_obs)},
            "e2e": {"value": Ke * rows * world / (ms_e * 1e-3), "unit": "samples/s",
                    "h2d_bytes_per_step": cfg["n_cond"] * 4, "d2h_bytes_per_step": rows * cfg["n_in"] * 4},
            "gpu_launches": K,
            "roofline": {"bound": "fp32_ffma", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": None,
                         "note": "algorithmic FLOPs = reference's dense D-pass count (2*M*L*D per sample); the kernel "
                                 "executes the degree-sorted incremental evaluation (~1/18 of those MACs), so frac "
                                 "above 1 is an algorithmic gain, not pipe utilisation; peak = FFMA measured live"},
            "clocks": clk,
        }
    else:
        # ------------------------------------------------------------- training step
        n_batches = max(8, (168 * 2**20) // (batch * (cfg["n_in"] + cfg["n_cond"]) * 4))  # > 126 MB L2
        n_batches = min(n_batches, 512)
        wide = name == "c5"
        if wide:
            n_batches = 8                                  # 8 x 65536 x 640 B = 336 MB > L2
        rows = n_batches * batch
        ds_bytes = rows * (cfg["n_in"] + cfg["n_cond"]) * 4
        l2_note = (f"resident dataset {ds_bytes / 2**20:.0f} MiB > 126 MB L2, batches walked by the on-device step counter"
                   if ds_bytes > 126e6 else
                   f"resident dataset {ds_bytes / 2**20:.1f} MiB fits the 126 MB L2 and is NOT flushed between steps: "
                   f"the step moves {batch * (cfg['n_in'] + cfg['n_cond']) * 4} input bytes and is latency-bound, "
                   "L2 residency of the inputs is immaterial")
        theta, x = synth_data(name, cfg, rows, seed=100 + rank)
        dv, cv = density_and_cond(name, theta, x)
        maf = ConditionalMAF(activation=bnn.relu, use_reverse=True, seed=42, device=dev, **cfg)
        # NDE_w_Standardization exactly as NPE/NLE build it (npe.py:383-413 / nle.py:323-353)
        model_hparams = {"nde": maf, "embedding_net": Sequential([Standardizer(*zscore(cv)), Identity()]),
                         "transformation": ScalarAffine(*zscore(dv))}
        loss_fn = loss_nll_nle if name == "c3" else loss_nll_npe
        exmp = (np.zeros((1, theta.shape[1]), np.float32), np.zeros((1, x.shape[1]), np.float32))
        trainer = TrainerModule(NDE_w_Standardization, model_hparams,
                                {"lr": 5e-4, "optimizer_name": "adam", "gradient_clip": 5.0, "warmup": 0.1,
                                 "weight_decay": 0.0},
                                loss_fn, exmp, seed=42, logger_params={"base_log_dir": "/tmp/jaxili_b200_bench"},
                                enable_progress_bar=False, nde_class="NLE" if name == "c3" else "NPE")
        trainer.init_optimizer(num_epochs=2**31 - 1, num_steps_per_epoch=n_batches)
        eng = trainer.model.engine()
        theta_d, x_d = torch.tensor(theta, device=dev), torch.tensor(x, device=dev)
        flops_step = train_flops_per_sample(cfg) * batch

        # (1) kernel-resident throughput
        if wide:
            # layer-by-layer path: ~150 launches per step, batches are device-resident slices
            state = trainer.state
            xd, yd = trainer._batch_xy([theta_d, x_d])
            flat = state.params.flat
            Pn = flat.numel()
            gbuf = trainer._gbuf
            grad, loss = gbuf[:Pn], gbuf[-4:-3]
            eng.pack(flat)
            inv = 1.0 / (batch * world)
            it_box = [0]

            def step_fn():
                i = it_box[0] % n_batches
                it_box[0] += 1
                eng.loss_grad(flat, xd[i * batch:(i + 1) * batch], yd[i * batch:(i + 1) * batch], loss, grad,
                              inv_global_b=inv, packed_current=True)
                if world > 1:
                    dist.all_reduce(gbuf)
                eng.clip_adam(flat, grad, state.opt_state["m"], state.opt_state["v"], state.step, state.tx,
                              norm_current=world == 1)
            launches_per_step = 10 * (3 + 1 + 1 + 3 * 3 + 1) + 6
        elif world == 1:
            step_fn = trainer.make_graph_step(theta_d, x_d, batch, steps_per_launch=graph_steps(K))
            launches_per_step = 2          # flow kernel, cooperative tail (reduce + norm + clip + Adam + re-pack)
        else:
            state = trainer.state
            xd, yd = trainer._batch_xy([theta_d, x_d])
            flat = state.params.flat
            Pn = flat.numel()
            gbuf = trainer._gbuf
            grad, loss = gbuf[:Pn], gbuf[-4:-3]
            eng.pack(flat)
            inv = 1.0 / (batch * world)

            from jaxili_b200 import parallel
            peer = (not os.environ.get("MAFB200_NO_PEER")) and parallel.setup_peer_allreduce(eng, state.step)
            if peer:
                # one-shot all-reduce over NVLink peer memory fused with the norm partials; no NCCL call
                # is left in the step, so the whole step is one CUDA graph like the single-GPU case
                step_fn = trainer.make_graph_step(theta_d, x_d, batch, steps_per_launch=graph_steps(K))
            else:
                def step_fn():
                    eng.loss_grad(flat, xd, yd, loss, grad, batch_rows=batch, inv_global_b=inv,
                                  batch_index=state.step, n_batches=n_batches, packed_current=True)
                    dist.all_reduce(gbuf)
                    eng.clip_adam(flat, grad, state.opt_state["m"], state.opt_state["v"], state.step, state.tx)
            # peer: flow kernel + cooperative tail (reduce, exchange, clip, Adam); NCCL: flow, reduce, NCCL's, clip_adam
            launches_per_step = 2 if peer else 4
        spl = getattr(step_fn, "steps", 1)           # steps per call (one CUDA graph may hold several)
        assert K % spl == 0, "--steps must be a multiple of the steps per graph launch"
        for _ in range((W + spl - 1) // spl):
            step_fn()
        barrier_sync()
        clocks.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K // spl):
            step_fn()
        e1.record()
        barrier_sync()
        clocks.mark_end()
        ms = max_over_ranks(e0.elapsed_time(e1))
        value = K * batch * world / (ms * 1e-3)
        final_loss = float(trainer._gbuf[-4].item())

        e2e = e2e_blocking = roof = None
        if full is not None:        # (full is None: strong-scaling leg, the resident value only)
            # (2) end to end through TrainerModule.train_epoch over HOST batches in pinned memory: every step's
            #     inputs are copied host->device inside the timed region (copy stream, overlapping the previous
            #     step) and every step's loss is copied device->host into a pinned ring right behind the step; the
            #     host reads the ring once per epoch, as the reference does (jaxili/train.py:446-449)
            # steady state: one epoch of at least 300 host batches for the narrow workloads (an epoch of 20 steps is
            # dominated by its fill and drain: 66.6 vs 76 M samples/s at c2); the count is reported as e2e.steps
            Ke = max(W, min(K, 10)) if wide else max(W, min(max(K, 300), 2000))
            n_host = min(n_batches, 16)
            host_batches = [[torch.from_numpy(np.ascontiguousarray(theta[i * batch:(i + 1) * batch])).pin_memory(),
                             torch.from_numpy(np.ascontiguousarray(x[i * batch:(i + 1) * batch])).pin_memory()]
                            for i in range(n_host)]
            trainer.train_epoch([host_batches[i % n_host] for i in range(max(W, 3))])
            loader = [host_batches[i % n_host] for i in range(Ke)]
            barrier_sync()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0 = time.perf_counter()
            t0.record()
            em = trainer.train_epoch(loader)      # returns after the last loss has reached the host
            t1.record()
            torch.cuda.synchronize()
            wall_ms = (time.perf_counter() - w0) * 1e3
            ms_e = max_over_ranks(max(t0.elapsed_time(t1), wall_ms))
            e2e = Ke * batch * world / (ms_e * 1e-3)
            assert np.isfinite(em["train/loss"])
            # the same through the blocking per-step call (train_step + .item() every step), for reference
            for i in range(W if full else 0):
                trainer.state, m = trainer.train_step(trainer.state, host_batches[i % n_host])
            barrier_sync()
            Kb = min(Ke, 100) if full else 0
            t0.record()
            for i in range(Kb):
                trainer.state, m = trainer.train_step(trainer.state, host_batches[i % n_host])
                _ = m["loss"].item()
            t1.record()
            barrier_sync()
            e2e_blocking = Kb * batch * world / (max_over_ranks(t0.elapsed_time(t1)) * 1e-3) if Kb else None

            # (3) roofline of the dominant kernel
            xd, yd = trainer._batch_xy([theta_d, x_d])
            flat = trainer.state.params.flat
            Pn = flat.numel()
            grad, loss = trainer._gbuf[:Pn], trainer._gbuf[-4:-3]
            eng.pack(flat)
            if wide:
                # the wide step is a chain of ~90 tcgen05 TF32 GEMM launches (+ elementwise kernels); the
                # roofline line is the chain's tensor-pipe rate: CUDA events around loss_grad
                r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                Kr = min(K, 10)
                r0.record()
                for i in range(Kr):
                    j = i % n_batches
                    eng.loss_grad(flat, xd[j * batch:(j + 1) * batch], yd[j * batch:(j + 1) * batch], loss, grad,
                                  packed_current=True)
                r1.record()
                torch.cuda.synchronize(dev)
                kernel_ms = r0.elapsed_time(r1) / Kr
                # TF32 tcgen05 peak measured live: back-to-back M=128,N=256,K=8 MMAs on resident operands, looped for
                # ~1.5 s so that the figure is the SUSTAINED one (a wide step is a seconds-long stream of GEMMs)
                t_end = time.perf_counter() + 1.5
                sustained = []
                burst = tf32_peak_tflops(local, 20000)
                while time.perf_counter() < t_end:
                    sustained.append(tf32_peak_tflops(local, 100000))
                peak_tc = float(np.median(sustained[len(sustained) // 2:])) if sustained else burst
                src = (f"tcgen05 kind::tf32 micro-benchmark run live (mafb200_tf32_peak): sustained {peak_tc:.0f} after "
                       f"1.5 s of back-to-back MMAs, burst {burst:.0f} TFLOP/s; MEASURED_PEAKS.json has no TF32 entry "
                       "(nominal 1100)")
                achieved = flops_step / (kernel_ms * 1e-3) / 1e12
                roof = {"bound": "tensor", "achieved": achieved, "peak": peak_tc, "unit": "TFLOP/s",
                        "frac": achieved / peak_tc, "traffic": None, "kernel": "wide_tc_gemm_kernel chain (loss_grad)",
                        "kernel_us": kernel_ms * 1e3, "flops_per_launch": flops_step, "peak_source": src,
                        "dtype_note": "TF32 operands (tensor core truncation), FP32 accumulate; tolerance in DESIGN.md"}
            for _ in range(0 if wide else W):
                eng.loss_grad(flat, xd, yd, loss, grad, batch_rows=batch, batch_index=trainer.state.step,
                              n_batches=n_batches, packed_current=True)
            torch.cuda.synchronize(dev)
            if not wide:
                eng.flow_timing(True)
                Kr = min(K, 300)
                for i in range(Kr):
                    trainer.state.step.add_(1)     # walk through the dataset (> L2) as the training loop does
                    eng.loss_grad(flat, xd, yd, loss, grad, batch_rows=batch, batch_index=trainer.state.step,
                                  n_batches=n_batches, packed_current=True)
                tot_ms, n_l = eng.flow_elapsed()
                eng.flow_timing(False)
                kernel_ms_bracketed = tot_ms / max(n_l, 1)      # one event pair per launch (eager launches)
                # the figure used: back-to-back launches of the flow kernel alone between ONE event pair, i.e. the
                # kernel's launch period in steady state, as it runs inside the step's graph.  The per-launch event
                # brackets above add their own stream operations around every kernel (52 vs ~47 us at c2).
                torch.cuda.synchronize(dev)
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0.record()
                for i in range(Kr):
                    eng.loss_grad(flat, xd, yd, loss, grad, batch_rows=batch, batch_index=trainer.state.step,
                                  n_batches=n_batches, packed_current=True, flow_only=True)
                f1.record()
                torch.cuda.synchronize(dev)
                kernel_ms = f0.elapsed_time(f1) / Kr
                peak, peak_parts = fp32_peak_tflops(local)
                achieved = flops_step / (kernel_ms * 1e-3) / 1e12
                static = {}
                try:
                    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                        static = json.load(f).get(name) or {}
                except Exception:
                    pass
                roof = {"bound": "fp32_ffma", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                        "frac": achieved / peak, "traffic": static.get("dram_bytes_per_launch"),
                        "traffic_source": static.get("source", "not captured for this workload"),
                        "kernel": "maf_chain_kernel<BWD> (csrc/chain_kernel.cuh)",
                        "kernel_us": kernel_ms * 1e3, "kernel_us_event_bracketed": kernel_ms_bracketed * 1e3,
                        "kernel_us_method": "back-to-back launches of the flow kernel alone between one CUDA event pair "
                                            "(launch period, same batch every launch: the input is 0.3 MB of the 14 MB a launch moves); "
                                            "kernel_us_event_bracketed: one event pair around every launch, walking through the dataset",
                        "flops_per_launch": flops_step,
                        "flops_note": "dense reference-equivalent count 6*M*L per row (SURVEY 8d), masks not discounted; "
                                      "the kernel multiplies the masked zeros too (no tile skipping); it skips the input "
                                      "gradient of layer 0 and computes only the data columns of each layer's input gradient",
                        "executed_fma_pipe_pct": static.get("fma_pipe_active_pct"),
                        "executed_fma_pipe_source": static.get("source"),
                        "peak_source": "FP32 FMA-pipe micro-benchmarks run live (MEASURED_PEAKS.json has no FP32 entry): "
                                       f"FFMA {peak_parts['ffma']:.1f}, packed FFMA2 {peak_parts['ffma2']:.1f} TFLOP/s, "
                                       "the larger is the denominator; nominal 74.5 TFLOP/s"}
        clk = clocks.stop()
        out = {
            "metric": "MAF train samples/sec (fwd+bwd+clip+Adam)", "value": value, "unit": "samples/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(name, batch, world),
            "run": {"dataset_rows_per_gpu": rows, "l2": l2_note, "final_loss": final_loss,
                    "steps_per_graph_launch": spl,
                    "allreduce": ("none" if world == 1 else ("nccl" if wide or not peer else "peer-memory one-shot (own kernel)"))},
            "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": batch * (cfg["n_in"] + cfg["n_cond"]) * 4,
                    "d2h_bytes_per_step": 4, "steps": Ke if full is not None else 0,
                    "api": "TrainerModule.train_epoch(pinned host batches)",
                    "blocking_per_step_value": e2e_blocking},
            "gpu_launches": launches_per_step * K,
            "roofline": roof,
            "clocks": clk,
        }
        if wide:
            out["dtype"] = "tf32"
            out["run"]["l2"] = "8 resident batches x 40 MiB > 126 MB L2, activations (GBs per step) stream through HBM"
        if e2e is None:
            out.pop("e2e")
        if roof is None:
            out.pop("roofline")

    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the secondary records (c3, c4, c5, strong scaling) of the default run")
    args = ap.parse_args()
    name = args.workload
    cfg, batch, desc = WORKLOADS[name]
    if args.impl == "reference":
        return run_reference_arm(args, name, cfg, batch)

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: jaxili_b200 has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier_sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = dict(world=world, rank=rank, local=local, dev=dev, barrier_sync=barrier_sync, max_over_ranks=max_over_ranks)
    W, K = max(args.warmup, 3), args.steps
    if name == "c5" and K > 200:
        K = 30          # a wide step is ~10 ms and 1.5 TFLOP: keep the default run within a minute
    if name == "c4" and K > 200:
        K = 50
    out = run_workload(ctx, name, batch, K, W, full=True)

    # ---- secondary records of the default run: the other BASELINE.json workloads at this N (fewer steps), and
    #      the strong-scaling figure (global batch fixed, B/N rows per GPU) for the training workloads
    if name == "c2" and not args.no_extra:
        extra = {}
        plan = [("c3", WORKLOADS["c3"][1], min(K, 1000)), ("c4", WORKLOADS["c4"][1], min(K, 30)),
                ("c5", WORKLOADS["c5"][1], min(K, 20))]
        for nm, b, k in plan:
            try:
                r = run_workload(ctx, nm, b, k, W, full=False)
                keep = ("metric", "value", "unit", "steps", "ms_per_step", "dtype", "scaling", "config", "run", "e2e",
                        "gpu_launches", "roofline", "clocks")
                extra[nm] = {kk: r[kk] for kk in keep if kk in r}
            except Exception as e:      # a secondary record must never take the headline line down
                extra[nm] = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.empty_cache()
        out["workloads"] = extra
        strong = {}
        for nm in ("c2", "c3", "c5"):
            gb = WORKLOADS[nm][1]
            if world == 1:
                src = out if nm == "c2" else extra.get(nm, {})
                if "value" in src:
                    strong[nm] = {"global_batch": gb, "per_gpu_batch": gb, "value": src["value"],
                                  "ms_per_step": src["ms_per_step"], "unit": "samples/s"}
                continue
            try:
                r = run_workload(ctx, nm, gb // world, min(K, 1000 if nm != "c5" else 20), W, full=None)
                strong[nm] = {"global_batch": gb, "per_gpu_batch": gb // world, "value": r["value"],
                              "ms_per_step": r["ms_per_step"], "unit": "samples/s"}
            except Exception as e:
                strong[nm] = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.empty_cache()
        out["strong"] = {"note": "global batch fixed, B/N contiguous rows per GPU (SURVEY 8e); value = global batch / "
                                 "max-over-ranks step time", "workloads": strong}

    if rank == 0 and not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        if name == "c4":
            v, dt = cpu_sample_throughput(cfg, 20000, threads)
            out["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": threads, "kind": "port",
                                   "sample": f"20000 samples of one observation in {dt:.1f} s (oracle/maf_torch.py)"}
        else:
            cb = batch if name != "c5" else 4096
            v, n, dt = cpu_train_throughput(name, cfg, cb, 12.0, threads)
            out["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": threads, "kind": "port",
                                   "sample": f"{n} train steps of batch {cb} in {dt:.1f} s (oracle/maf_torch.py)"}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
