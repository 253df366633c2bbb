"""MafEngine: thin host wrapper around one native ``mafb200_handle``.

Device memory, streams and (elsewhere) ``torch.distributed`` are torch's; every FLOP
of the hot path runs in the hand-written sm_100a kernels behind the C ABI.  All
tensors are FP32, contiguous, on the engine's CUDA device; calls enqueue on torch's
current stream and do not synchronise.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .masks import maf_degrees


def _f32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _np_ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class MafEngine:
    """One ConditionalMAF (+ standardisation constants) bound to one GPU."""

    def __init__(self, n_in, n_cond, n_layers, layers, activation="relu", use_reverse=True,
                 seed=42, shift=None, scale=None, mean=None, std=None, device=None):
        if activation not in _lib.ACTIVATIONS:
            raise ValueError(
                f"activation {activation!r} is not supported by the fused kernels; "
                f"choose one of {sorted(_lib.ACTIVATIONS)}")
        if not (1 <= len(layers) <= _lib.MAX_HIDDEN):
            raise ValueError("layers must have between 1 and 5 entries")
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.MafB200Error("no CUDA device: jaxili_b200 has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device) \
            if not isinstance(device, torch.device) else device
        self.n_in, self.n_cond, self.n_layers = int(n_in), int(n_cond), int(n_layers)
        self.layers = [int(h) for h in layers]
        self.activation, self.use_reverse, self.seed = activation, bool(use_reverse), seed
        self.degrees = maf_degrees(self.n_in, self.layers, self.n_layers, seed)
        desc = _lib.MafDesc()
        desc.n_in, desc.n_cond, desc.n_layers = self.n_in, self.n_cond, self.n_layers
        desc.n_hidden = len(self.layers)
        for i, h in enumerate(self.layers):
            desc.hidden[i] = h
        desc.activation = _lib.ACTIVATIONS[activation]
        desc.use_reverse = int(self.use_reverse)
        self._std = [_f32(shift), _f32(scale), _f32(mean), _f32(std)]
        handle = C.c_void_p()
        _lib.check(self.lib.mafb200_create(
            C.byref(desc), self.degrees.ctypes.data_as(C.c_void_p),
            *[_np_ptr(a) for a in self._std], self.device.index, C.byref(handle)))
        self._h = handle
        self.n_params = int(self.lib.mafb200_param_count(self._h))
        self.dims = [self.n_in + self.n_cond, *self.layers, 2 * self.n_in]

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self.lib.mafb200_destroy(h)
            except Exception:
                pass
            self._h = None

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _pc(self, params, packed_current, image="flow"):
        """PACKED_CURRENT flag for a call on ``params``: honoured only when the handle's packed image (``flow``: the
        weight images / masked copy, ``sample``: the sorted sampler image) was last written from this very buffer.
        The image lives in the shared handle: a call with other parameters in between (another checkpoint, a
        posterior built on an older state) re-packs it, and the next ``packed_current=True`` call on the trainer's
        buffer must pack again instead of running on foreign weights."""
        key = "_packed_from_" + image
        ok = bool(packed_current) and getattr(self, key, None) == params.data_ptr()
        setattr(self, key, params.data_ptr())
        if self.wide and image == "sample":
            setattr(self, "_packed_from_flow", params.data_ptr())     # the wide sampler runs on the masked copy
        return _lib.PACKED_CURRENT if ok else 0

    def _chk(self, t, name, cols=None, align=16):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise TypeError(f"{name} must be a contiguous float32 CUDA tensor")
        if t.device != self.device:
            raise ValueError(f"{name} is on {t.device}, engine is on {self.device}")
        if t.data_ptr() % align:
            raise ValueError(f"{name} must be {align}-byte aligned")
        if cols is not None and (t.dim() != 2 or t.shape[1] != cols):
            raise ValueError(f"{name} must have shape (rows, {cols}), got {tuple(t.shape)}")
        return C.c_void_p(t.data_ptr())

    def set_standardization(self, shift=None, scale=None, mean=None, std=None):
        self._std = [_f32(shift), _f32(scale), _f32(mean), _f32(std)]
        _lib.check(self.lib.mafb200_set_standardization(self._h, *[_np_ptr(a) for a in self._std]))

    def mask(self):
        """0/1 uint8 vector in flat parameter order (biases 1)."""
        m = np.empty(self.n_params, dtype=np.uint8)
        _lib.check(self.lib.mafb200_get_mask(self._h, m.ctypes.data_as(C.c_void_p)))
        return m

    def query(self):
        v = [C.c_int32() for _ in range(4)]
        _lib.check(self.lib.mafb200_query(self._h, *[C.byref(a) for a in v]))
        return dict(smem_bytes=v[0].value, tile_rows=v[1].value, max_ctas=v[2].value, stash_all=v[3].value)

    @property
    def wide(self):
        """True when the MADE does not fit shared memory and runs layer by layer (wide path)."""
        if not hasattr(self, "_wide"):
            self._wide = self.query()["smem_bytes"] == 0
        return self._wide

    def set_wide_core(self, core):
        """Wide path GEMM core: "tf32" (tcgen05, default) or "ffma" (FP32, exact)."""
        _lib.check(self.lib.mafb200_set_wide_core(self._h, {"ffma": 0, "tf32": 1}[core]))

    def eval_accumulate(self, params, x, y, acc, packed_current=False):
        """acc[0] += -sum(log_prob(x | y)), acc[1] += rows, on the device (eval_model's size-weighted mean)."""
        px = self._chk(x, "x", self.n_in, align=4)
        py = self._chk(y, "y", self.n_cond, align=4) if self.n_cond else None
        B = x.shape[0]
        if self.n_cond and y.shape[0] != B:
            raise ValueError("x and y must have the same number of rows")
        if B:
            _lib.check(self.lib.mafb200_eval_accumulate(
                self._h, self._chk(params, "params"), px, py, B, self._chk(acc, "acc", align=4),
                self._pc(params, packed_current), self._stream()))

    def set_flow_kernel(self, mode):
        """Fused-kernel selection: "auto" (chain kernel for training, barrier-phased otherwise), "phased", "chain"."""
        _lib.check(self.lib.mafb200_set_flow_kernel(self._h, {"auto": 0, "phased": 1, "chain": 2}[mode]))

    def flow_timing(self, enable=True):
        _lib.check(self.lib.mafb200_flow_timing(self._h, 1 if enable else 0))

    def flow_elapsed(self):
        """(total device ms, launches) of the bracketed flow-kernel launches."""
        ms, n = C.c_double(), C.c_int64()
        _lib.check(self.lib.mafb200_flow_elapsed(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # ------------------------------------------------------------------ hot path
    def pack(self, params):
        _lib.check(self.lib.mafb200_pack(self._h, self._chk(params, "params"), self._stream()))
        self._packed_from_flow = params.data_ptr()

    def invalidate_packed(self):
        """Forget what the packed images were written from (call after changing parameters outside the library)."""
        self._packed_from_flow = self._packed_from_sample = None

    def log_prob(self, params, x, y, out=None, packed_current=False):
        px = self._chk(x, "x", self.n_in, align=4)     # misaligned row tiles take the kernels' element-wise staging path
        py = self._chk(y, "y", self.n_cond, align=4) if self.n_cond else None
        B = x.shape[0]
        if self.n_cond and y.shape[0] != B:
            raise ValueError("x and y must have the same number of rows")
        if out is None:
            out = torch.empty(B, dtype=torch.float32, device=self.device)
        if B:
            _lib.check(self.lib.mafb200_log_prob(
                self._h, self._chk(params, "params"), px, py, B, self._chk(out, "out"),
                self._pc(params, packed_current), self._stream()))
        return out

    def forward(self, params, x, y, packed_current=False):
        """ConditionalMAF.__call__: returns (u, log_det) for standardised-or-not inputs."""
        px = self._chk(x, "x", self.n_in, align=4)     # misaligned row tiles take the kernels' element-wise staging path
        py = self._chk(y, "y", self.n_cond, align=4) if self.n_cond else None
        B = x.shape[0]
        u = torch.empty((B, self.n_in), dtype=torch.float32, device=self.device)
        ld = torch.empty(B, dtype=torch.float32, device=self.device)
        if B:
            _lib.check(self.lib.mafb200_forward(
                self._h, self._chk(params, "params"), px, py, B, None, self._chk(u, "u"),
                self._chk(ld, "ld"), self._pc(params, packed_current), self._stream()))
        return u, ld

    def inverse(self, params, u, y, packed_current=False):
        """ConditionalMAF.backward with per-row conditioning: returns (x, log_det)."""
        pu = self._chk(u, "u", self.n_in)
        py = self._chk(y, "y", self.n_cond, align=4) if self.n_cond else None
        N = u.shape[0]
        if self.n_cond and y.shape[0] != N:
            raise ValueError("u and y must have the same number of rows")
        x = torch.empty_like(u)
        ld = torch.empty(N, dtype=torch.float32, device=self.device)
        if N:
            _lib.check(self.lib.mafb200_inverse(
                self._h, self._chk(params, "params"), pu, py, N, self._chk(x, "x"),
                self._chk(ld, "ld"), self._pc(params, packed_current, "sample"), self._stream()))
        return x, ld

    def log_prob_grad_y(self, params, x, y, packed_current=False):
        """(log_prob [B], d log_prob / d y [B, n_cond]): gradient w.r.t. the conditioning input."""
        px = self._chk(x, "x", self.n_in, align=4)     # misaligned row tiles take the kernels' element-wise staging path
        py = self._chk(y, "y", self.n_cond, align=4)
        B = x.shape[0]
        if y.shape[0] != B:
            raise ValueError("x and y must have the same number of rows")
        lp = torch.empty(B, dtype=torch.float32, device=self.device)
        dy = torch.empty((B, self.n_cond), dtype=torch.float32, device=self.device)
        if B:
            _lib.check(self.lib.mafb200_log_prob_grad_y(
                self._h, self._chk(params, "params"), px, py, B, self._chk(lp, "lp"), self._chk(dy, "dy"),
                self._pc(params, packed_current), self._stream()))
        return lp, dy

    def hmc_trajectory(self, params, x_rep, state, n_steps, packed_current=False):
        """``n_steps`` leapfrog steps of all chains (``n_steps=0``: evaluate gradient and log density only).
        ``state``: dict of CUDA tensors eps[1], inv_mass[C], prior_kind[C] (int32), lo[C], hi[C], z[n,C],
        p[n,C], g[n,C], logdens[n], theta[n,C], scratch[n*(C+1)] -- updated in place."""
        n = state["z"].shape[0]
        px = self._chk(x_rep, "x", self.n_in, align=4)
        if x_rep.shape[0] != n:
            raise ValueError("x must hold one observation row per chain")
        st = _lib.MafHmcState()
        for name, shape, dt in (("eps", (1,), torch.float32), ("inv_mass", (self.n_cond,), torch.float32),
                                ("prior_kind", (self.n_cond,), torch.int32), ("lo", (self.n_cond,), torch.float32),
                                ("hi", (self.n_cond,), torch.float32), ("z", (n, self.n_cond), torch.float32),
                                ("p", (n, self.n_cond), torch.float32), ("g", (n, self.n_cond), torch.float32),
                                ("logdens", (n,), torch.float32), ("theta", (n, self.n_cond), torch.float32),
                                ("scratch", (n * (self.n_cond + 1),), torch.float32)):
            t = state[name]
            if tuple(t.shape) != shape or t.dtype != dt or not t.is_contiguous() or t.device != self.device:
                raise ValueError(f"hmc state '{name}' must be a contiguous {dt} tensor of shape {shape} on {self.device}")
            setattr(st, name, t.data_ptr())
        _lib.check(self.lib.mafb200_hmc_trajectory(
            self._h, self._chk(params, "params"), px, n, C.byref(st), int(n_steps),
            self._pc(params, packed_current), self._stream()))

    # ---- data parallel over NVLink peer memory
    def dp_init(self, rank, world):
        """Allocate this rank's symmetric gradient buffer; returns its 64-byte CUDA IPC handle."""
        buf = C.create_string_buffer(64)
        _lib.check(self.lib.mafb200_dp_init(self._h, int(rank), int(world), buf))
        return buf.raw

    def dp_connect(self, handles):
        """``handles``: the ranks' IPC handles, rank-major (list of 64-byte strings)."""
        blob = b"".join(handles)
        _lib.check(self.lib.mafb200_dp_connect(self._h, C.c_char_p(blob)))

    def dp_bind_step(self, step):
        """``step``: the int32 CUDA tensor that clip_adam increments (kept alive by the caller)."""
        _lib.check(self.lib.mafb200_dp_bind_step(self._h, C.c_void_p(step.data_ptr())))
        self._dp_step = step

    def dp_reset(self):
        _lib.check(self.lib.mafb200_dp_reset(self._h))

    def dp_allreduce(self, out_grad, out_loss):
        """One-shot peer all-reduce of the slots written by ``loss_grad(dp_slot=True)``; leaves the
        norm partials current (``clip_adam(norm_current=True)``)."""
        _lib.check(self.lib.mafb200_dp_allreduce(self._h, self._chk(out_grad, "out_grad"),
                                                 self._chk(out_loss, "out_loss"), self._stream()))

    def dp_error(self):
        e = C.c_int32()
        _lib.check(self.lib.mafb200_dp_error(self._h, C.byref(e)))
        return e.value

    def loss_grad(self, params, x, y, out_loss, out_grad, batch_rows=None, inv_global_b=None,
                  batch_index=None, n_batches=1, packed_current=False, dp_slot=False, flow_only=False):
        """out_loss[0], out_grad <- NLL and its gradient over rows [0, batch_rows) of x/y (or the
        device-selected batch ``batch_index[0] % n_batches``).  ``flow_only`` (measurement hook): only the fused
        flow kernel is launched, out_loss / out_grad stay untouched."""
        px = self._chk(x, "x", self.n_in, align=4)     # misaligned row tiles take the kernels' element-wise staging path
        py = self._chk(y, "y", self.n_cond, align=4) if self.n_cond else None
        B = x.shape[0] if batch_rows is None else int(batch_rows)
        if batch_index is not None:
            if x.shape[0] < B * n_batches:
                raise ValueError("x holds fewer than n_batches * batch_rows rows")
            bi = C.c_void_p(batch_index.data_ptr())
        else:
            bi = None
        inv = 1.0 / B if inv_global_b is None else float(inv_global_b)
        _lib.check(self.lib.mafb200_loss_grad(
            self._h, self._chk(params, "params"), px, py, B, inv, bi, int(n_batches),
            None if dp_slot else self._chk(out_loss, "out_loss"), None if dp_slot else self._chk(out_grad, "out_grad"),
            (self._pc(params, packed_current)) | (_lib.DP_SLOT if dp_slot else 0) |
            (_lib.FLOW_ONLY if flow_only else 0), self._stream()))

    @staticmethod
    def make_opt_desc(lr=1e-3, warmup=0.0, decay_steps=1e9, clip=5.0, b1=0.9, b2=0.999, eps=1e-8,
                      weight_decay=0.0, kind="adam", init_factor=0.01, end_factor=0.01):
        o = _lib.MafOptDesc()
        o.lr, o.warmup, o.decay_steps = lr, warmup, decay_steps
        o.init_factor, o.end_factor = init_factor, end_factor
        o.b1, o.b2, o.eps, o.clip, o.weight_decay = b1, b2, eps, clip, weight_decay
        o.kind = {"adam": 0, "sgd": 1, "adamw": 2}[kind]
        return o

    def train_step(self, params, x, y, out_loss, out_grad, m, v, step, opt, batch_rows=None, inv_global_b=None,
                   batch_index=None, n_batches=1, packed_current=False, dp_slot=False):
        """loss_grad [+ peer all-reduce] + clip_adam as one call: flow kernel + one cooperative tail kernel.
        ``dp_slot=True``: data parallel over NVLink peer memory (``dp_init`` / ``dp_connect`` / ``dp_bind_step``
        done, ``inv_global_b`` = 1 / global batch); the exchange happens inside the tail kernel."""
        if step.dtype != torch.int32 or not step.is_cuda:
            raise TypeError("step must be an int32 CUDA tensor")
        self._packed_from_sample = None        # the update refreshes the flow images only
        px = self._chk(x, "x", self.n_in, align=4)     # misaligned row tiles take the kernels' element-wise staging path
        py = self._chk(y, "y", self.n_cond, align=4) if self.n_cond else None
        B = x.shape[0] if batch_rows is None else int(batch_rows)
        if batch_index is not None:
            if x.shape[0] < B * n_batches:
                raise ValueError("x holds fewer than n_batches * batch_rows rows")
            bi = C.c_void_p(batch_index.data_ptr())
        else:
            bi = None
        inv = 1.0 / B if inv_global_b is None else float(inv_global_b)
        _lib.check(self.lib.mafb200_train_step(
            self._h, self._chk(params, "params"), px, py, B, inv, bi, int(n_batches),
            self._chk(out_loss, "out_loss"), self._chk(out_grad, "out_grad"),
            self._chk(m, "m") if m is not None else None, self._chk(v, "v") if v is not None else None,
            C.c_void_p(step.data_ptr()), C.byref(opt),
            (self._pc(params, packed_current)) | (_lib.DP_SLOT if dp_slot else 0), self._stream()))

    def clip_adam(self, params, grad, m, v, step, opt, norm_current=False):
        """In-place optax-chain update; ``step`` is an int32 CUDA tensor (the optax count)."""
        self._packed_from_sample = None        # the update refreshes the flow images only
        if step.dtype != torch.int32 or not step.is_cuda:
            raise TypeError("step must be an int32 CUDA tensor")
        _lib.check(self.lib.mafb200_clip_adam(
            self._h, self._chk(params, "params"), self._chk(grad, "grad"),
            self._chk(m, "m") if m is not None else None, self._chk(v, "v") if v is not None else None,
            C.c_void_p(step.data_ptr()), C.byref(opt), _lib.NORM_CURRENT if norm_current else 0,
            self._stream()))

    def sample_from_noise(self, params, u, y_obs, rows_per_obs=None, out=None, packed_current=False):
        pu = self._chk(u, "u", self.n_in)
        N = u.shape[0]
        y_obs = y_obs.reshape(-1, self.n_cond) if self.n_cond else y_obs
        n_obs = y_obs.shape[0] if self.n_cond else 1
        if rows_per_obs is None:
            rows_per_obs = max(1, -(-N // n_obs))
        if out is None:
            out = torch.empty_like(u)
        if N:
            _lib.check(self.lib.mafb200_sample_from_noise(
                self._h, self._chk(params, "params"), pu,
                self._chk(y_obs, "y_obs", align=4) if self.n_cond else None, n_obs, int(rows_per_obs), N,
                self._chk(out, "out"), self._pc(params, packed_current, "sample"), self._stream()))
        return out

    def sample_philox(self, params, N, y_obs, seed=0, row_offset=0, rows_per_obs=None, out=None,
                      packed_current=False):
        y_obs = y_obs.reshape(-1, self.n_cond) if self.n_cond else y_obs
        n_obs = y_obs.shape[0] if self.n_cond else 1
        if rows_per_obs is None:
            rows_per_obs = max(1, -(-N // n_obs))
        if out is None:
            out = torch.empty((N, self.n_in), dtype=torch.float32, device=self.device)
        if N:
            _lib.check(self.lib.mafb200_sample_philox(
                self._h, self._chk(params, "params"), int(seed) & (2**64 - 1), int(row_offset),
                self._chk(y_obs, "y_obs", align=4) if self.n_cond else None, n_obs, int(rows_per_obs), int(N),
                self._chk(out, "out"), self._pc(params, packed_current, "sample"), self._stream()))
        return out


def ffma_peak_tflops(device=0, iters=4096, packed=False):
    """Measured FP32 FMA-pipe throughput of the GPU (register-tiled GEMM operand pattern): plain FFMA or, with
    packed=True, the packed FFMA2 the fused kernels issue."""
    lib = _lib.load()
    out = C.c_float()
    _lib.check(lib.mafb200_fp32_peak(int(device), int(iters), 1 if packed else 0, C.byref(out)))
    return float(out.value)


def fp32_peak_tflops(device=0):
    """Roofline denominator of the fused kernels: the better of the FFMA and FFMA2 probes, and both figures."""
    a, b = ffma_peak_tflops(device), ffma_peak_tflops(device, packed=True)
    return max(a, b), {"ffma": a, "ffma2": b}


def tf32_peak_tflops(device=0, iters=20000):
    """Measured TF32 tcgen05 tensor-pipe throughput (resident operands, back-to-back MMAs)."""
    lib = _lib.load()
    out = C.c_float()
    _lib.check(lib.mafb200_tf32_peak(int(device), int(iters), C.byref(out)))
    return float(out.value)
