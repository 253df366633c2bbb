"""TrainerModule with jaxili's surface (jaxili/train.py), driving the fused B200 kernels.

Kept from the reference: constructor signature and defaults (train.py:48-62), the config /
``hparams.json`` / ``metrics/`` / ``<base>/<model_class>/version_N`` layout (123-162, 542-556),
``init_optimizer`` (246-300: clip_by_global_norm -> warmup-cosine schedule -> adam | sgd | adamw),
``create_functions`` returning ``train_step(state, batch) -> (state, {"loss"})`` and
``eval_step(state, batch) -> {"loss"}`` (317-342), the epoch loops with the reference's averaging
rules (427-487), strict-improvement model selection (489-519), flax ``EarlyStopping`` semantics
(382, 401-411), ``bind_model`` (625-636) and ``load_from_checkpoints`` (638-684).

Replaced: XLA autodiff + optax tree-maps -> ``mafb200_loss_grad`` + ``mafb200_clip_adam``; orbax ->
``.npz`` keyed by the flax parameter path; TensorBoard/W&B -> JSON-lines logger (neither orbax nor
pytorch-lightning is installable here).  ``state`` is updated in place and returned, which keeps
``state, m = train_step(state, batch)`` working.

Data parallelism: if ``torch.distributed`` is initialised, every rank runs the kernels on its shard
with 1/B_global scaling and the flat gradient (+ loss) is all-reduced once per step over NCCL before
the (replicated, bit-identical) clip + Adam update.
"""

from __future__ import annotations

import json
import os
import time
import warnings
from collections import defaultdict
from copy import copy, deepcopy
from typing import Any, Callable, Dict, Iterator, Optional

import numpy as np
import torch

from . import layout, parallel
from .loss import jaxili_loss_dict, loss_nll_nle, loss_nll_npe
from .model import NDE_w_Standardization, ConditionalMAF, key_to_seed
from .nn import jax_nn_dict
from .utils import handle_non_serializable

try:  # progress bar is optional plumbing
    from tqdm import tqdm
except Exception:  # pragma: no cover
    tqdm = None


class EarlyStopping:
    """flax.training.early_stopping.EarlyStopping semantics: an epoch improves iff
    ``best - metric > min_delta`` (or best is inf); otherwise ``should_stop`` is evaluated on the
    patience counter *before* it is incremented, so training stops patience+1 epochs after the
    last improvement."""

    def __init__(self, min_delta=0.0, patience=0, best_metric=float("inf"), patience_count=0,
                 should_stop=False, has_improved=False):
        self.min_delta, self.patience = min_delta, patience
        self.best_metric, self.patience_count = best_metric, patience_count
        self.should_stop, self.has_improved = should_stop, has_improved

    def update(self, metric):
        if np.isinf(self.best_metric) or self.best_metric - metric > self.min_delta:
            return EarlyStopping(self.min_delta, self.patience, metric, 0, False, True)
        should_stop = self.patience_count >= self.patience or self.should_stop
        return EarlyStopping(self.min_delta, self.patience, self.best_metric, self.patience_count + 1,
                             should_stop, False)


class JsonLogger:
    """Stand-in for the pytorch-lightning loggers: ``<save_dir>/version_N`` + metrics.jsonl."""

    def __init__(self, save_dir, version=None, name=""):
        root = os.path.join(save_dir, name) if name else save_dir
        if version is None:
            os.makedirs(root, exist_ok=True)
            existing = [int(d.split("_")[1]) for d in os.listdir(root)
                        if d.startswith("version_") and d.split("_")[1].isdigit()]
            version = f"version_{max(existing) + 1 if existing else 0}"
        self.log_dir = os.path.join(root, version) if version != "" else root
        os.makedirs(self.log_dir, exist_ok=True)

    def log_metrics(self, metrics, step=None):
        with open(os.path.join(self.log_dir, "metrics.jsonl"), "a") as f:
            f.write(json.dumps({"step": step, **{k: float(v) for k, v in metrics.items()}}) + "\n")

    def finalize(self, status):
        pass


class TrainState:
    """Counterpart of jaxili's TrainState (train.py:29-38): step, params, opt_state, tx, batch_stats, rng.
    ``step`` is the optax count and lives on the device (int32) so a training step never syncs."""

    def __init__(self, step, apply_fn, params, tx, opt_state, batch_stats=None, rng=None):
        self.step, self.apply_fn, self.params = step, apply_fn, params
        self.tx, self.opt_state, self.batch_stats, self.rng = tx, opt_state, batch_stats, rng


class ResidentLoader:
    """GPU-resident dataset with on-device epoch shuffling: batches are views of
    device tensors, so ``train_step`` does no host->device copy."""

    def __init__(self, theta, x, batch_size, shuffle=True, drop_last=True, seed=42, device=None):
        device = device or torch.device("cuda", torch.cuda.current_device())
        self.theta = torch.as_tensor(theta, dtype=torch.float32).to(device).contiguous()
        self.x = torch.as_tensor(x, dtype=torch.float32).to(device).contiguous()
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), shuffle, drop_last
        self.gen = torch.Generator(device=device).manual_seed(seed)

    def __len__(self):
        n = self.theta.shape[0]
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def __iter__(self):
        n = self.theta.shape[0]
        if self.shuffle:
            perm = torch.randperm(n, generator=self.gen, device=self.theta.device)
            theta, x = self.theta[perm], self.x[perm]
        else:
            theta, x = self.theta, self.x
        for i in range(len(self)):
            s = slice(i * self.batch_size, (i + 1) * self.batch_size)
            yield [theta[s], x[s]]


class ResidentDataset:
    """A split of a GPU-resident simulation set: ``theta`` / ``x`` are device tensors."""

    def __init__(self, theta, x):
        self.theta, self.x = theta, x

    def __len__(self):
        return self.theta.shape[0]

    def __getitem__(self, idx):
        return self.theta[idx], self.x[idx]


def split_resident(theta, x, train_test_split, key, device=None):
    """On-device random split with the reference's size arithmetic (jaxili/inference/npe.py:255-286: ``int()`` of
    the cumulative fractions): one upload of the whole simulation set, a device permutation seeded from ``key``,
    three gathers.  Returns (train, val, test-or-None) ``ResidentDataset``s."""
    from .model import key_to_seed
    device = device or torch.device("cuda", torch.cuda.current_device())
    th = torch.as_tensor(theta, dtype=torch.float32).to(device)
    xx = torch.as_tensor(x, dtype=torch.float32).to(device)
    n = th.shape[0]
    fr = list(train_test_split)
    if len(fr) not in (2, 3):
        raise ValueError("train_test_split should have 2 or 3 elements.")
    assert abs(sum(fr) - 1.0) < 1e-6, "The sum of the split fractions should be 1."
    if key is None:
        key = np.random.randint(0, 1000)
    gen = torch.Generator(device=device).manual_seed(key_to_seed(key) % (2**63))
    perm = torch.randperm(n, generator=gen, device=device)
    n_tr = int(fr[0] * n)
    n_tv = int((fr[0] + fr[1]) * n)
    parts = [perm[:n_tr], perm[n_tr:n_tv], perm[n_tv:] if len(fr) == 3 else None]
    return tuple(None if p is None else ResidentDataset(th[p].contiguous(), xx[p].contiguous()) for p in parts)


def _capture(body, stream):
    """Capture ``body()`` (library launches only) into a CUDA graph.  Relaxed capture mode and a collected,
    paused garbage collector: an unrelated object dying mid-capture (a pinned buffer, another graph) would
    otherwise issue a CUDA call that invalidates the capture."""
    import gc
    gc.collect()
    was = gc.isenabled()
    gc.disable()
    try:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream, capture_error_mode="relaxed"):
            body()
    finally:
        if was:
            gc.enable()
    return graph


class HostBatchPipeline:
    """Streams host batches through the fused step without ever idling the GPU on the host: a thin wrapper
    over the native executor ``mafb200_pipeline_*`` (csrc/pipeline.inc).

    The reference's epoch loop hands numpy batches to a jitted step and relies on JAX's asynchronous
    dispatch, reading the loss once per epoch (jaxili/train.py:427-451).  The same contract here:
      * two device batch buffers; the H2D copy of batch i+1 runs on a copy stream while the compute stream
        executes step i (events order buffer reuse in both directions);
      * the step (flow fwd+bwd and the cooperative tail, or flow + reduce + peer all-reduce + clip/Adam under
        torch.distributed) is one CUDA graph per buffer;
      * every step's loss is copied D2H into a pinned ring on a third stream; the host reads the ring after
        ``drain()`` (one synchronisation per epoch, like ``.item()`` at train.py:449).
    Batches that are pinned torch tensors are copied from where they lie; anything else goes through the
    executor's own pinned staging slots, whose reuse is guarded by the copy events."""

    def __init__(self, trainer, rows):
        import ctypes as C
        from . import _lib
        self.t, self.rows = trainer, rows
        eng = trainer.model.engine()
        self.eng = eng
        state = trainer.state
        _, world = parallel.world_info()
        self.peer = world > 1 and parallel.setup_peer_allreduce(eng, state.step)
        if world > 1 and not self.peer:
            raise NotImplementedError("the streamed step under torch.distributed needs the peer-memory all-reduce")
        flat = state.params.flat
        P = flat.numel()
        self.flat, self.grad = flat, trainer._gbuf[:P]
        self._keep = (flat, state.opt_state["m"], state.opt_state["v"], state.step, trainer._gbuf, state.tx)
        self.lib, self._C, self._check = eng.lib, C, _lib.check
        h = C.c_void_p()
        cur = torch.cuda.current_stream(eng.device).cuda_stream
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        self._check(self.lib.mafb200_pipeline_create(
            eng._h, rows, ptr(flat), ptr(state.opt_state["m"]), ptr(state.opt_state["v"]), ptr(state.step),
            ptr(self.grad), C.byref(state.tx), 1.0 / (rows * world), _lib.DP_SLOT if self.peer else 0,
            C.c_void_p(cur), C.byref(h)))
        self._p = h
        if self.peer:
            parallel.reset_peer_epochs(eng)          # create's throw-away step advanced, then rewound, the count
        self._check(self.lib.mafb200_pipeline_capture(self._p))
        self.i = self.base = 0
        self._out = np.empty(8192, np.float32)
        # pinned batches are copied asynchronously from where they lie, on the executor's own copy stream, which
        # torch's caching host allocator knows nothing about: keep them alive until drain() has seen their copies
        self._inflight = []
        trainer._packed = True

    def __del__(self):
        p, self._p = getattr(self, "_p", None), None
        if p:
            try:
                self.lib.mafb200_pipeline_destroy(p)
            except Exception:
                pass

    def fence(self):
        """Order the pipeline behind everything enqueued on torch's current stream."""
        cur = torch.cuda.current_stream(self.eng.device).cuda_stream
        self._check(self.lib.mafb200_pipeline_fence(self._p, self._C.c_void_p(cur)))

    def _host(self, a, cols):
        """(pointer, pinned?, keep-alive) of a host batch [rows, cols]."""
        if isinstance(a, torch.Tensor):
            if a.dtype != torch.float32 or not a.is_contiguous():
                a = a.to(torch.float32).contiguous()
            if a.numel() != self.rows * cols:
                raise ValueError(f"batch has {a.numel()} values, the pipeline expects {self.rows} x {cols}")
            return a.data_ptr(), a.is_pinned(), a
        a = np.ascontiguousarray(a, dtype=np.float32)
        if a.size != self.rows * cols:
            raise ValueError(f"batch has {a.size} values, the pipeline expects {self.rows} x {cols}")
        return a.ctypes.data, False, a

    def submit(self, batch):
        xb, yb = self.t._batch_xy(batch)
        px, pin_x, kx = self._host(xb, self.eng.n_in)
        if self.eng.n_cond:
            py, pin_y, ky = self._host(yb, self.eng.n_cond)
        else:
            py, pin_y, ky = None, True, None
        if self.i - self.base >= 8192:
            raise RuntimeError("loss ring full: call drain() at least every 8192 steps")
        pinned = pin_x and pin_y
        self._check(self.lib.mafb200_pipeline_submit(self._p, self._C.c_void_p(px),
                                                     self._C.c_void_p(py) if py else None, 1 if pinned else 0))
        if pinned:
            self._inflight.append((kx, ky))     # released in drain(), after the copies have completed
        self.i += 1

    def drain(self):
        """Wait for everything submitted; returns the per-step losses since the previous drain (numpy)."""
        n = self._C.c_int64()
        self._check(self.lib.mafb200_pipeline_drain(
            self._p, self._C.c_void_p(self._out.ctypes.data), self._out.size, self._C.byref(n),
            self._C.c_void_p(self.t._gbuf[-4:-3].data_ptr())))     # where the plain step leaves its loss
        self.base = self.i
        self._inflight.clear()                  # drain() returns after every submitted copy and step has finished
        return self._out[: n.value].copy()


class TrainerModule:
    """Training loop, evaluation, logging and checkpointing around the fused kernels."""

    def __init__(self, model_class, model_hparams: Dict[str, Any], optimizer_hparams: Dict[str, Any],
                 loss_fn: Callable, exmp_input: Any, seed: int = 42,
                 logger_params: Dict[str, Any] = None, enable_progress_bar: bool = True,
                 debug: bool = False, check_val_every_epoch: int = 1, nde_class: str = "NPE", **kwargs):
        self.model_class = model_class
        self.model_hparams = model_hparams
        self.loss_fn = loss_fn
        self.optimizer_hparams = optimizer_hparams
        self.enable_progress_bar = enable_progress_bar
        self.debug = debug
        self.seed = seed
        self.check_val_every_epoch = check_val_every_epoch
        self.nde_class = nde_class
        assert nde_class == "NPE" or nde_class == "NLE", \
            "Choose a valid class of Neural Density Estimator. (NPE or NLE)"
        self.exmp_input = exmp_input
        if self.nde_class == "NLE":
            self.exmp_input = (self.exmp_input[1], self.exmp_input[0])
        self.generate_config(logger_params)
        self.config.update(kwargs)
        self.model = self.model_class(**self.model_hparams)
        if not isinstance(self.model, (ConditionalMAF, NDE_w_Standardization)):
            raise NotImplementedError("the B200 trainer drives ConditionalMAF / NDE_w_Standardization only")
        self.init_apply_fn()
        self.init_logger(logger_params)
        self.create_jitted_functions()
        self.init_model(self.exmp_input)
        self._bufs = {}
        self._buf_events = {}
        self._pipeline = None
        self._pipelines = {}

    # ------------------------------------------------------------------ setup
    def init_logger(self, logger_params: Optional[Dict] = None):
        """jaxili/train.py:123-162 (JSON logger instead of TensorBoard / W&B)."""
        if logger_params is None:
            logger_params = dict()
        log_dir = logger_params.get("log_dir", None)
        if log_dir is None:
            base_log_dir = logger_params.get("base_log_dir", "checkpoints/")
            log_dir = os.path.join(base_log_dir, self.config["model_class"])
            if "logger_name" in logger_params:
                log_dir = os.path.join(log_dir, logger_params["logger_name"])
            version = None
        else:
            version = ""
        logger_type = logger_params.get("logger_type", "TensorBoard").lower()
        assert logger_type in ("tensorboard", "wandb", "json"), f'Unknown logger type "{logger_type}"'
        self.logger = JsonLogger(save_dir=log_dir, version=version, name="")
        log_dir = self.logger.log_dir
        if not os.path.isfile(os.path.join(log_dir, "hparams.json")):
            os.makedirs(os.path.join(log_dir, "metrics/"), exist_ok=True)
            try:
                with open(os.path.join(log_dir, "hparams.json"), "w") as f:
                    json.dump(self.config, f, indent=4, default=handle_non_serializable)
            except Exception:
                warnings.warn("Could not save hyperparameters.", Warning)
        self.log_dir = log_dir

    def init_model(self, exmp_input: Any):
        """jaxili/train.py:164-190: fresh parameters; the optimizer is created later."""
        variables = self.run_model_init(exmp_input, self.seed)
        self.state = TrainState(step=None, apply_fn=self.apply_fn, params=variables["params"],
                                batch_stats=None, rng=self.seed, tx=None, opt_state=None)

    def init_apply_fn(self):
        self.apply_fn = self.model.log_prob

    def generate_config(self, logger_params):
        """jaxili/train.py:196-211."""
        self.config = {
            "model_class": self.model_class.__name__,
            "model_hparams": dict(self.model_hparams),
            "loss_fn": self.loss_fn.__name__,
            "optimizer_hparams": self.optimizer_hparams,
            "logger_params": logger_params,
            "enable_progress_bar": self.enable_progress_bar,
            "debug": self.debug,
            "check_val_every_epoch": self.check_val_every_epoch,
            "seed": self.seed,
        }
        if "activation" in self.model_hparams.keys():
            act = self.model_hparams["activation"]
            self.config["model_hparams"]["activation"] = act if isinstance(act, str) else act.__name__

    def run_model_init(self, exmp_input: Any, init_rng: Any) -> Dict:
        return self.model.init(init_rng, *exmp_input, method="log_prob")

    def print_tabulate(self, exmp_input: Any):
        eng = self.model.engine()
        print(f"{type(self.model).__name__}: {eng.n_params} parameters, dims {eng.dims} x {eng.n_layers} layers")

    def init_optimizer(self, num_epochs: int, num_steps_per_epoch: int):
        """jaxili/train.py:246-300: optax chain restated as one fused kernel's descriptor."""
        hparams = copy(self.optimizer_hparams)
        optimizer_name = hparams.pop("optimizer_name", "adam").lower()
        assert optimizer_name in ("adam", "sgd", "adamw"), f'Unknown optimizer "{optimizer_name}"'
        lr = hparams.pop("lr", 1e-3)
        warmup = hparams.pop("warmup", num_steps_per_epoch)
        decay_steps = hparams.pop("decay_steps", int(num_epochs // 2 * num_steps_per_epoch))
        clip = hparams.pop("gradient_clip", 5.0)
        wd = hparams.pop("weight_decay", None)
        if optimizer_name == "sgd":
            weight_decay = 0.0 if wd is None else wd          # add_decayed_weights only for sgd (287-288)
        elif optimizer_name == "adamw":
            weight_decay = hparams.pop("adamw_weight_decay", 1e-4)  # popped in the reference -> optax default
        else:
            weight_decay = 0.0
        eng = self.model.engine()
        tx = eng.make_opt_desc(lr=float(lr), warmup=float(warmup), decay_steps=float(decay_steps),
                               clip=float(clip), b1=float(hparams.pop("b1", 0.9)),
                               b2=float(hparams.pop("b2", 0.999)), eps=float(hparams.pop("eps", 1e-8)),
                               weight_decay=float(weight_decay), kind=optimizer_name)
        flat = self.model.flat_params(self.state.params)
        params = layout.tree_from_flat(flat, self.model.dims, self.model.n_layers,
                                       wrap_nde=isinstance(self.model, NDE_w_Standardization))
        dev = flat.device
        P = flat.numel()
        self._gbuf = torch.zeros(((P + 3) // 4) * 4 + 4, dtype=torch.float32, device=dev)  # grad + loss
        self.state = TrainState(
            step=torch.zeros(1, dtype=torch.int32, device=dev), apply_fn=self.state.apply_fn, params=params,
            tx=tx, opt_state={"m": torch.zeros_like(flat), "v": torch.zeros_like(flat)},
            batch_stats=self.state.batch_stats, rng=self.state.rng)
        self._packed = False

    def create_jitted_functions(self):
        """jaxili/train.py:302-315 (nothing to jit: the step is already two fused launches)."""
        self.train_step, self.eval_step = self.create_functions()

    # ------------------------------------------------------------------ the step
    def _to_device(self, a, cols, slot):
        """host batch -> device through a pinned staging buffer (async H2D on the current stream)."""
        eng = self.model.engine()
        if isinstance(a, torch.Tensor) and a.is_cuda:
            return a.to(torch.float32).reshape(-1, cols).contiguous()
        a = np.asarray(a, dtype=np.float32).reshape(-1, cols)
        key = (slot, a.shape)
        if key not in self._bufs:
            self._bufs[key] = (torch.empty(a.shape, dtype=torch.float32).pin_memory(),
                               torch.empty(a.shape, dtype=torch.float32, device=eng.device))
        pin, dev = self._bufs[key]
        ev = self._buf_events.get(key)
        if ev is not None:
            ev.synchronize()                   # the previous H2D out of this staging buffer has finished
        else:
            ev = self._buf_events[key] = torch.cuda.Event()
        pin.numpy()[...] = a
        dev.copy_(pin, non_blocking=True)
        ev.record()
        return dev

    def _batch_xy(self, batch):
        """(density variable, conditioning variable) for this trainer's loss (jaxili/loss.py:55-57, 81-83)."""
        thetas, xs = batch
        if self.loss_fn is loss_nll_nle or getattr(self.loss_fn, "__name__", "") == "loss_nll_nle":
            return xs, thetas
        if self.loss_fn is loss_nll_npe or getattr(self.loss_fn, "__name__", "") == "loss_nll_npe":
            return thetas, xs
        raise NotImplementedError(
            "the fused backward implements loss_nll_npe / loss_nll_nle; other losses have no gradient here")

    def _nll_loss(self):
        try:
            self._batch_xy([None, None])
            return True
        except NotImplementedError:
            return False

    def create_functions(self):
        """jaxili/train.py:317-342."""
        def train_step(state: TrainState, batch: Any):
            if state.tx is None:
                raise ValueError("optimizer not initialised: call init_optimizer / train_model first")
            eng = self.model.engine()
            xd, yd = self._batch_xy(batch)
            x = self._to_device(xd, eng.n_in, "x")
            y = self._to_device(yd, eng.n_cond, "y") if eng.n_cond else None
            flat = state.params.flat
            P = flat.numel()
            grad, loss = self._gbuf[:P], self._gbuf[-4:-3]
            _, world = parallel.world_info()
            if world > 1 and parallel.setup_peer_allreduce(eng, state.step):
                # latency-bound gradient (<= 1 MB): one-shot all-reduce over NVLink peer memory, fused
                # with the norm partials; replicas stay bit-identical (fixed rank order)
                eng.train_step(flat, x, y, loss, grad, state.opt_state["m"], state.opt_state["v"], state.step,
                               state.tx, inv_global_b=parallel.inv_global_batch(x.shape[0]),
                               packed_current=self._packed, dp_slot=True)
                self._packed = True
                return state, {"loss": loss[0].clone()}
            if world == 1:
                eng.train_step(flat, x, y, loss, grad, state.opt_state["m"], state.opt_state["v"], state.step,
                               state.tx, packed_current=self._packed)
                self._packed = True
                return state, {"loss": loss[0].clone()}
            parallel.dp_step(
                lambda g: eng.loss_grad(flat, x, y, loss, grad, inv_global_b=parallel.inv_global_batch(x.shape[0]),
                                        packed_current=self._packed),
                lambda g: eng.clip_adam(flat, grad, state.opt_state["m"], state.opt_state["v"], state.step,
                                        state.tx, norm_current=world == 1),
                self._gbuf)                   # one NCCL all-reduce: flat gradient + loss
            self._packed = True
            return state, {"loss": loss[0].clone()}

        def eval_step(state: TrainState, batch: Any):
            eng = self.model.engine()
            xd, yd = self._batch_xy(batch)
            x = self._to_device(xd, eng.n_in, "ex")
            y = self._to_device(yd, eng.n_cond, "ey") if eng.n_cond else None
            flat = self.model.flat_params(state.params)
            lp = eng.log_prob(flat, x, y)
            return {"loss": -lp.mean()}

        return train_step, eval_step

    def make_graph_step(self, theta_dev, x_dev, batch_size, steps_per_launch=1):
        """CUDA-graph replay of one whole step over a GPU-resident dataset: the batch is selected on
        the device from the optax step counter, so one graph launch = forward + backward + reduction
        + clip + Adam with no host work.  Returns ``step() -> None`` (loss in ``self.last_loss``).
        ``steps_per_launch`` > 1 captures that many consecutive steps in the one graph (``step.steps``)."""
        state = self.state
        eng = self.model.engine()
        _, world = parallel.world_info()
        peer = world > 1 and parallel.setup_peer_allreduce(eng, state.step)
        if world > 1 and not peer:
            # capturing the NCCL all-reduce together with the kernels hung on 2xB200 (torch 2.11 /
            # NCCL 2.28.9); without the peer-memory path multi-rank steps are launched eagerly
            raise NotImplementedError("make_graph_step under torch.distributed needs the peer-memory all-reduce")
        xd, yd = self._batch_xy([theta_dev, x_dev])
        n_batches = xd.shape[0] // batch_size
        flat = state.params.flat
        P = flat.numel()
        grad, loss = self._gbuf[:P], self._gbuf[-4:-3]
        self.last_loss = loss
        eng.pack(flat)

        inv = 1.0 / (batch_size * world)

        def one():
            # flow kernel + one cooperative tail kernel; with peers the one-shot exchange over NVLink peer
            # memory happens inside the tail (own kernel, graph-capturable -- unlike the NCCL call)
            eng.train_step(flat, xd, yd, loss, grad, state.opt_state["m"], state.opt_state["v"], state.step,
                           state.tx, batch_rows=batch_size, inv_global_b=inv, batch_index=state.step,
                           n_batches=n_batches, packed_current=True, dp_slot=peer)

        def body():
            for _ in range(max(1, int(steps_per_launch))):
                one()

        s = torch.cuda.Stream(device=eng.device)
        s.wait_stream(torch.cuda.current_stream(eng.device))
        with torch.cuda.stream(s):
            body()                       # warm-up outside capture
            s.synchronize()
            graph = _capture(body, s)
        torch.cuda.current_stream(eng.device).wait_stream(s)
        self._packed = True
        self._graph = graph

        def step():
            graph.replay()
        step.steps = max(1, int(steps_per_launch))
        step.check = self._check_dp      # call at a host sync point: raises if a peer rank went missing
        return step

    # ------------------------------------------------------------------ loops
    def train_model(self, train_loader: Iterator, val_loader: Iterator, test_loader: Optional[Iterator] = None,
                    num_epochs: int = 500, min_delta: float = 1e-3, patience: int = 20) -> Dict[str, Any]:
        """jaxili/train.py:344-425."""
        self.init_optimizer(num_epochs, len(train_loader))
        self.on_training_start()
        best_eval_metrics = None
        best_epoch = None
        early_stop = EarlyStopping(min_delta, patience)
        pbar = self.tracker(range(1, num_epochs + 1), desc="Epochs")
        epoch_idx = 0
        for epoch_idx in pbar:
            train_metrics = self.train_epoch(train_loader)
            self.logger.log_metrics(train_metrics, step=epoch_idx)
            self.on_training_epoch_end(epoch_idx)
            if epoch_idx % self.check_val_every_epoch == 0:
                eval_metrics = self.eval_model(val_loader, log_prefix="val/")
                self.on_validation_epoch_end(epoch_idx, eval_metrics, val_loader)
                self.logger.log_metrics(eval_metrics, step=epoch_idx)
                self.save_metrics(f"eval_epoch_{str(epoch_idx).zfill(3)}", eval_metrics)
                if self.is_new_model_better(eval_metrics, best_eval_metrics):
                    best_eval_metrics = eval_metrics
                    best_eval_metrics.update(train_metrics)
                    best_epoch = epoch_idx
                    self.save_model(step=epoch_idx)
                    self.save_metrics("best_eval", best_eval_metrics)
                early_stop = early_stop.update(eval_metrics["val/loss"])
                if early_stop.should_stop:
                    print(f"Neural network training stopped after {epoch_idx} epochs.")
                    print(f"Early stopping with best validation metric: {early_stop.best_metric}")
                    print(f"Best model saved at epoch {best_epoch}")
                    print(f"Early stopping parameters: min_delta={min_delta}, patience={patience}")
                    break
                if self.enable_progress_bar and hasattr(pbar, "set_description"):
                    pbar.set_description(
                        f"Epochs: Val loss {eval_metrics['val/loss']:.3f}/ Best val loss {early_stop.best_metric:.3f}")
        if test_loader is not None:
            test_metrics = self.eval_model(test_loader, log_prefix="test/")
            self.logger.log_metrics(test_metrics, step=epoch_idx)
            self.save_metrics("test", test_metrics)
            best_eval_metrics.update(test_metrics)
        self.logger.finalize("success")
        return best_eval_metrics

    def train_epoch(self, train_loader: Iterator) -> Dict[str, Any]:
        """jaxili/train.py:427-451: plain mean of the step losses, one host sync per epoch."""
        metrics = defaultdict(float)
        num_train_steps = len(train_loader)
        start_time = time.time()
        if self._streamable():
            return self._train_epoch_streamed(train_loader, num_train_steps, start_time)
        for batch in train_loader:
            self.state, step_metrics = self.train_step(self.state, batch)
            for key in step_metrics:
                metrics["train/" + key] = metrics["train/" + key] + step_metrics[key] / num_train_steps
        metrics = {key: float(metrics[key].item()) if hasattr(metrics[key], "item") else float(metrics[key])
                   for key in metrics}
        self._check_dp()
        metrics["epoch_time"] = time.time() - start_time
        return metrics

    def _streamable(self):
        """The pipelined epoch serves the stock step only: a subclass that overrides ``create_functions``
        (jaxili/train.py:317-328 invites that) or a non-NLL loss keeps the plain loop."""
        if getattr(self, "stream_host_batches", True) is False or self.state is None or self.state.tx is None:
            return False
        if type(self).create_functions is not TrainerModule.create_functions:
            return False
        try:
            self._batch_xy([None, None])
        except NotImplementedError:
            return False
        _, world = parallel.world_info()
        eng = self.model.engine()
        return world == 1 or (not eng.wide and parallel.setup_peer_allreduce(eng, self.state.step))

    def host_pipeline(self, rows):
        """The (cached) ``HostBatchPipeline`` for batches of ``rows`` rows and the current optimiser state."""
        cache = self._pipelines
        p = cache.get(rows)
        if p is None or p.t.state is not self.state or p.flat is not self.state.params.flat:
            if p is None and len(cache) >= 4:        # a loader has at most a full and a ragged batch size
                cache.clear()
            p = cache[rows] = HostBatchPipeline(self, rows)
        self._pipeline = p
        return p

    def _train_epoch_streamed(self, train_loader, num_train_steps, start_time):
        pipe, total, n = None, 0.0, 0
        dev_total = None                             # losses of device-resident steps: summed on the device
        for batch in train_loader:
            first = batch[0]
            rows = first.shape[0]
            if isinstance(first, torch.Tensor) and first.is_cuda:
                # device-resident batches need no staging: plain step; its loss word is added to a device scalar
                # (stream-ordered, before the next step overwrites it) and read once at the end of the epoch,
                # like the reference's single .item() (jaxili/train.py:446-449)
                if pipe is not None:
                    total += float(pipe.drain().sum())
                    pipe = None                      # re-fenced behind the plain step when used again
                self.state, m = self.train_step(self.state, batch)
                loss_t = m["loss"]
                if isinstance(loss_t, torch.Tensor):
                    loss_t = loss_t.reshape(-1)[:1].to(torch.float32)
                    dev_total = loss_t.clone() if dev_total is None else dev_total.add_(loss_t)
                else:
                    total += float(loss_t)
                n += 1
                continue
            if pipe is None or pipe.rows != rows:
                if pipe is not None:
                    total += float(pipe.drain().sum())
                pipe = self.host_pipeline(rows)
                eng = self.model.engine()
                if getattr(eng, "_packed_from_flow", None) != pipe.flat.data_ptr():
                    eng.pack(pipe.flat)          # another parameter set was packed into the handle in between
                pipe.fence()
            if pipe.i - pipe.base >= 8192:
                total += float(pipe.drain().sum())
            pipe.submit(batch)
            n += 1
        if pipe is not None:
            total += float(pipe.drain().sum())
        if dev_total is not None:
            total += float(dev_total.item())
        self._check_dp()
        return {"train/loss": total / max(num_train_steps, 1), "epoch_time": time.time() - start_time}

    def _check_dp(self):
        """Raise when the peer-memory gradient exchange lost a rank (the kernels then leave params / m / v / step
        untouched and report NaN losses; the error word is read at the loops' existing host syncs)."""
        _, world = parallel.world_info()
        if world > 1 and self.model.engine().dp_error():
            raise RuntimeError("data-parallel exchange timed out: a peer rank did not reach the step "
                               "(its process died or stalled for longer than the exchange time-out)")

    def eval_model(self, data_loader: Iterator, log_prefix: Optional[str] = "") -> Dict[str, Any]:
        """jaxili/train.py:453-487: size-weighted mean over the batches.  With the stock eval_step and an NLL loss
        the whole loop stays on the device: every batch adds -sum(log_prob) and its row count to two device
        scalars (mafb200_eval_accumulate), read once at the end -- the reference's single ``.item()``."""
        if type(self).create_functions is TrainerModule.create_functions and self._nll_loss():
            eng = self.model.engine()
            acc = torch.zeros(2, dtype=torch.float32, device=eng.device)
            flat = self.model.flat_params(self.state.params)
            first = True
            for batch in data_loader:
                xd, yd = self._batch_xy(batch)
                x = self._to_device(xd, eng.n_in, "ex")
                y = self._to_device(yd, eng.n_cond, "ey") if eng.n_cond else None
                eng.eval_accumulate(flat, x, y, acc, packed_current=not first)
                first = False
            a = acc.cpu().numpy()
            return {(log_prefix + "loss"): float(a[0] / max(a[1], 1.0))}
        metrics = defaultdict(float)
        num_elements = 0
        for batch in data_loader:
            step_metrics = self.eval_step(self.state, batch)
            batch_size = batch[0].shape[0] if isinstance(batch, (list, tuple)) else batch.shape[0]
            for key in step_metrics:
                metrics[key] = metrics[key] + step_metrics[key] * batch_size
            num_elements += batch_size
        return {(log_prefix + key): float((metrics[key] / num_elements).item()) for key in metrics}

    def is_new_model_better(self, new_metrics: Dict[str, Any], old_metrics: Dict[str, Any]) -> bool:
        """jaxili/train.py:489-519."""
        if old_metrics is None:
            return True
        for key, is_larger in [("val/val_metric", False), ("val/acc", True), ("val/loss", False)]:
            if key in new_metrics:
                if is_larger:
                    return new_metrics[key] > old_metrics[key]
                else:
                    return new_metrics[key] < old_metrics[key]
        assert False, f"No known metrics to log on: {new_metrics}"

    def tracker(self, iterator: Iterator, **kwargs) -> Iterator:
        if self.enable_progress_bar and tqdm is not None:
            return tqdm(iterator, **kwargs)
        return iterator

    def save_metrics(self, filename: str, metrics: Dict[str, Any]):
        with open(os.path.join(self.log_dir, f"metrics/{filename}.json"), "w") as f:
            json.dump(metrics, f, indent=4)

    def on_training_start(self):
        pass

    def on_training_epoch_end(self, epoch_idx: int):
        pass

    def on_validation_epoch_end(self, epoch_idx: int, eval_metrics: Dict[str, Any], val_loader: Iterator):
        pass

    # ------------------------------------------------------------------ checkpoints
    def save_model(self, step: int = 0):
        """Params only, as in jaxili/train.py:598-611; ``.npz`` keyed by the flax parameter path
        (max_to_keep=1: the file is overwritten)."""
        tree = layout.tree_to_numpy(self.state.params)
        flatmap = {"params/" + k: v for k, v in layout.flatten_paths(tree).items()}
        np.savez(os.path.join(self.log_dir, "checkpoint.npz"), __step__=np.int64(step), **flatmap)

    def load_model(self):
        """jaxili/train.py:613-623."""
        z = np.load(os.path.join(self.log_dir, "checkpoint.npz"))
        tree = layout.unflatten_paths({k[len("params/"):]: z[k] for k in z.files if k.startswith("params/")})
        eng = self.model.engine()
        flat = layout.flat_from_tree(tree, self.model.dims, self.model.n_layers, device=eng.device)
        params = layout.tree_from_flat(flat, self.model.dims, self.model.n_layers,
                                       wrap_nde=isinstance(self.model, NDE_w_Standardization))
        self.state = TrainState(step=self.state.step, apply_fn=self.apply_fn, params=params,
                                tx=self.state.tx, opt_state=self.state.opt_state,
                                batch_stats=self.state.batch_stats, rng=self.state.rng)
        self._packed = False

    def bind_model(self):
        """jaxili/train.py:625-636."""
        return self.model.bind({"params": self.state.params})

    @classmethod
    def load_from_checkpoints(cls, model_class, checkpoint: str, exmp_input: Any) -> Any:
        """jaxili/train.py:638-684."""
        hparams_file = os.path.join(checkpoint, "hparams.json")
        assert os.path.isfile(hparams_file), "Could not find hparams file."
        with open(hparams_file, "r") as f:
            hparams = json.load(f)
        assert hparams["model_class"] == model_class.__name__, \
            "Model class does not match. Check that you are using the correct architecture."
        hparams.pop("model_class")
        assert hparams["loss_fn"] in jaxili_loss_dict, "Unknown loss function."
        hparams["loss_fn"] = jaxili_loss_dict[hparams["loss_fn"]]
        if "activation" in hparams["model_hparams"].keys():
            hparams["model_hparams"]["activation"] = jax_nn_dict[hparams["model_hparams"]["activation"]]
        if not hparams["logger_params"]:
            hparams["logger_params"] = dict()
        hparams["logger_params"]["log_dir"] = checkpoint
        trainer = cls(model_class=model_class, exmp_input=exmp_input, **hparams)
        trainer.load_model()
        return trainer
