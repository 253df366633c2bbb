"""CPU tests of the host-side logic and of the C-ABI boundary (no compute calls without a GPU)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from jaxili_b200 import _lib, layout, masks, parallel
from oracle import maf_numpy as onp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    from jaxili_b200 import build
    path = build.build()
    assert os.path.isfile(path)
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "mafb200.h")).read()
    declared = set(re.findall(r"\b(mafb200_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    raw = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/mafb200.h but not exported"
    assert declared == set(_lib.SYMBOLS), "ctypes table and header disagree"
    assert lib.mafb200_version() >= 100
    assert lib.mafb200_last_error() is not None


def test_product_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from jaxili_b200.engine import MafEngine
    with pytest.raises(_lib.MafB200Error):
        MafEngine(2, 2, 2, [8])
    from jaxili_b200.model import ConditionalMAF
    with pytest.raises(_lib.MafB200Error):
        ConditionalMAF(2, 2, 2, [8], activation="relu", seed=1).init(0)
    # the C entry point reports the missing device itself (no CPU fallback behind the ABI either)
    lib = _lib.load()
    desc = _lib.MafDesc()
    desc.n_in, desc.n_cond, desc.n_layers, desc.n_hidden = 2, 2, 1, 1
    desc.hidden[0] = 4
    deg = np.zeros(4, np.int32)
    h = ctypes.c_void_p()
    rc = lib.mafb200_create(ctypes.byref(desc), deg.ctypes.data_as(ctypes.c_void_p), None, None, None, None, 0,
                            ctypes.byref(h))
    assert rc != 0 and b"no CPU fallback" in lib.mafb200_last_error()


def test_product_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under jaxili_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "jaxili_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f"{f} imports the oracle"


def test_degrees_and_masks_match_oracle():
    for n_in, n_cond, L, hid in [(10, 10, 5, [50, 50]), (2, 2, 5, [50, 50]), (8, 5, 5, [50, 50]), (3, 5, 4, [128, 128]),
                                 (6, 0, 2, [16, 8, 12])]:
        np.random.seed(999)                   # the builder must not disturb the caller's global RNG state
        before = np.random.get_state()[1].copy()
        deg = masks.maf_degrees(n_in, hid, L, 42)
        assert (np.random.get_state()[1] == before).all()
        spec = onp.MafSpec(n_in, n_cond, L, hid, seed=42)
        assert masks.maf_layer_seeds(42, L) == spec.seeds
        for k in range(L):
            for a, b in zip(masks.masks_from_degrees(n_in, n_cond, hid, deg[k]), spec.masks[k]):
                assert (a == b).all()
    with pytest.raises(ValueError):
        masks.made_hidden_degrees(1, [4], 0)


def test_layout_round_trip_and_flax_paths():
    dims, L = [20, 50, 50, 20], 5
    P = layout.n_params(dims, L)
    assert P == 23100
    flat = torch.arange(P, dtype=torch.float32)
    tree = layout.tree_from_flat(flat, dims, L, wrap_nde=True)
    paths = layout.flatten_paths(tree)
    assert "nde/layer_list_0/ConditionalMADE_0/layers_0/Dense_0/kernel" in paths
    assert "nde/layer_list_4/ConditionalMADE_0/layers_4/Dense_0/bias" in paths
    assert len(paths) == 5 * 3 * 2
    assert paths["nde/layer_list_0/ConditionalMADE_0/layers_0/Dense_0/kernel"].shape == (20, 50)
    assert paths["nde/layer_list_0/ConditionalMADE_0/layers_2/Dense_0/kernel"][0, 1] == 20 * 50 + 50 + 1
    # views: writing a leaf writes the flat buffer
    paths["nde/layer_list_1/ConditionalMADE_0/layers_0/Dense_0/bias"][3] = -7
    assert flat[4620 + 1000 + 3] == -7
    assert layout.flat_from_tree(tree, dims, L) is flat
    rebuilt = layout.flat_from_tree(layout.unflatten_paths(layout.flatten_paths(layout.tree_to_numpy(tree))), dims, L)
    assert torch.equal(rebuilt, flat)
    # the oracle uses the same flat order
    spec = onp.MafSpec(10, 10, 5, [50, 50])
    lay = spec.unflatten(flat.numpy())
    assert (lay[1][0][1] == tree["nde"]["layer_list_1"]["ConditionalMADE_0"]["layers_0"]["Dense_0"]["bias"].numpy()).all()
    with pytest.raises(ValueError):
        bad = layout.tree_to_numpy(tree)
        bad["nde"]["layer_list_0"]["ConditionalMADE_0"]["layers_0"]["Dense_0"]["kernel"] = np.zeros((3, 3))
        layout.flat_from_tree(bad, dims, L)


def test_early_stopping_known_answers():
    """Notebook known answers: best epoch 45 -> stop at 66 with patience 20."""
    from jaxili_b200.train import EarlyStopping
    for best in (45, 75, 249, 189):
        es = EarlyStopping(min_delta=1e-3, patience=20)
        stopped = None
        for epoch in range(1, 1000):
            metric = 1.0 - 0.01 * epoch if epoch <= best else 1.0 - 0.01 * best + 1e-4
            es = es.update(metric)
            if es.should_stop:
                stopped = epoch
                break
        assert stopped == best + 21
    es = EarlyStopping(min_delta=1e-3, patience=2).update(1.0)
    assert es.has_improved and not es.update(0.9995).has_improved   # improvement below min_delta does not count


def test_split_sizes_and_loaders():
    """Split sizes follow the reference's int() arithmetic: 35000/9999/5001 for N = 50000
    (notebooks/npe_validation.ipynb:1875-1879); loaders: shuffle+drop_last for train only."""
    from jaxili_b200.inference.npe import NDEDataset, split_indices
    from jaxili_b200.utils import create_data_loader, numpy_collate, validate_theta_x
    tr, va, te = split_indices(50_000, [0.7, 0.2, 0.1], key=0)
    assert (len(tr), len(va), len(te)) == (35_000, 9_999, 5_001)
    assert len(set(tr) | set(va) | set(te)) == 50_000
    tr2, va2, te2 = split_indices(1000, [0.7, 0.3], key=1)
    assert te2 is None and len(tr2) + len(va2) == 1000
    with pytest.raises(ValueError):
        split_indices(10, [1.0], key=0)
    with pytest.raises(AssertionError):
        split_indices(10, [0.5, 0.2, 0.1], key=0)
    theta = np.arange(300, dtype=np.float32).reshape(100, 3)
    x = -theta[:, :2].copy()
    ds = NDEDataset(theta, x)
    train, val = create_data_loader(ds, ds, train=[True, False], batch_size=32)
    assert len(train) == 3 and len(val) == 4
    batches = list(train)
    assert all(b[0].shape == (32, 3) and b[1].shape == (32, 2) for b in batches)
    assert all((b[1] == -b[0][:, :2]).all() for b in batches)            # rows stay paired
    first_epoch = np.concatenate([b[0][:, 0] for b in batches])
    second_epoch = np.concatenate([b[0][:, 0] for b in train])
    assert not np.array_equal(first_epoch, second_epoch)                 # reshuffled every epoch
    vb = list(val)
    assert vb[-1][0].shape[0] == 4 and (np.concatenate([b[0] for b in vb]) == theta).all()
    assert numpy_collate([(np.ones(2), np.zeros(1)), (np.ones(2), np.zeros(1))])[0].shape == (2, 2)
    with pytest.raises(AssertionError):
        validate_theta_x(theta.astype(np.float64), x)
    with pytest.raises(AssertionError):
        validate_theta_x(theta, x[:5])
    with pytest.raises(AssertionError):
        validate_theta_x(list(theta), x)


def test_keys_activations_and_sharding():
    from jaxili_b200 import nn
    from jaxili_b200.model import key_to_seed
    assert key_to_seed(7) == 7 and key_to_seed(None) == 0
    assert key_to_seed(np.array([0, 42], dtype=np.uint32)) == 42
    assert key_to_seed(np.array([1, 2], dtype=np.uint32)) == (1 << 32) ^ 2
    assert nn.activation_name(nn.relu) == "relu" and nn.activation_name("swish") == "silu"

    def custom(x):
        return x
    with pytest.raises(ValueError):
        nn.activation_name(custom)
    assert set(nn.jax_nn_dict) >= {"relu", "silu", "tanh", "gelu", "elu", "sigmoid", "softplus", "leaky_relu"}
    # row sharding covers every row exactly once
    for n, w in [(10, 3), (16384, 8), (5, 8)]:
        blocks = [parallel.shard_rows(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        assert max(b[1] - b[0] for b in blocks) - min(b[1] - b[0] for b in blocks) <= 1


def test_bench_helpers_flops():
    """SURVEY 8(d) algorithmic FLOP table."""
    sys.path.insert(0, ROOT)
    import bench
    assert bench.train_flops_per_sample(bench.WORKLOADS["c1"][0]) == 87_000
    assert bench.train_flops_per_sample(bench.WORKLOADS["c2"][0]) == 135_000
    assert bench.train_flops_per_sample(bench.WORKLOADS["c3"][0]) == 118_500
    assert bench.sample_flops_per_sample(bench.WORKLOADS["c4"][0]) == 450_000
    th, x = bench.synth_data("c3", bench.WORKLOADS["c3"][0], 100, 0)
    assert th.shape == (100, 5) and x.shape == (100, 8) and x.dtype == np.float32


_DP_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from jaxili_b200 import parallel
from oracle import maf_numpy as onp
from tests.util import CONFIGS, make_case
rank, world = int(sys.argv[2]), int(sys.argv[3])
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[4]}", rank=rank, world_size=world)
spec, flat, std, x, y = make_case(CONFIGS["c3"], 48, seed=9)
P = spec.n_params
params = flat.astype(np.float64)
st = onp.AdamState(P, np.float64)
lo, hi = parallel.shard_rows(48, rank, world)
assert hi - lo == 24
gbuf = torch.zeros(P + 1, dtype=torch.float64)
for step in range(3):
    def compute_local(g):
        l, gr = onp.nll_loss_and_grad(spec, params, std, x[lo:hi].astype(np.float64), y[lo:hi].astype(np.float64),
                                      inv_b=parallel.inv_global_batch(hi - lo))
        g[:P] = torch.from_numpy(gr); g[P] = l
    def apply_update(g):
        global params
        params, _, _ = onp.clip_adam_step(params, g[:P].numpy(), st, 1e-2, 0.1, 100, clip=0.5)
    parallel.dp_step(compute_local, apply_update, gbuf)
np.save(sys.argv[5] + f"/rank{rank}.npy", np.concatenate([params, [gbuf[P].item()]]))
dist.destroy_process_group()
'''


def test_data_parallel_world_size_2_gloo(tmp_path):
    """N>1 host path on CPU: two gloo ranks with the oracle standing in for the kernels; replicas end
    bit-identical and equal to the single-process run on the global batch."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_DP_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), "2", str(port), str(tmp_path)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out.decode()[-2000:]
    a, b = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert np.array_equal(a, b)
    from tests.util import CONFIGS, make_case
    spec, flat, std, x, y = make_case(CONFIGS["c3"], 48, seed=9)
    params = flat.astype(np.float64)
    st = onp.AdamState(spec.n_params, np.float64)
    for _ in range(3):
        loss, g = onp.nll_loss_and_grad(spec, params, std, x.astype(np.float64), y.astype(np.float64))
        params, _, _ = onp.clip_adam_step(params, g, st, 1e-2, 0.1, 100, clip=0.5)
    assert np.max(np.abs(a[:-1] - params)) < 1e-12
    assert abs(a[-1] - loss) < 1e-12 * abs(loss)


def test_hmc_restatement_is_symplectic_and_reversible():
    """oracle/hmc_numpy.py on an analytic Gaussian likelihood: energy error O(eps^2), exact reversibility."""
    from oracle import hmc_numpy
    rng = np.random.default_rng(0)
    n, C = 16, 4
    prec = np.array([1.0, 4.0, 0.25, 2.0])

    def ll_grad(theta):
        return -0.5 * (theta ** 2 * prec).sum(1), -theta * prec

    kind = np.array([0, 1, 0, 1]); lo = np.array([0.0, -3.0, 0.5, -1.0]); hi = np.array([2.0, 3.0, 1.0, 4.0])
    z = rng.normal(size=(n, C)); p = rng.normal(size=(n, C)); im = np.ones(C)
    ld, g, _ = hmc_numpy.log_density_and_grad(z, kind, lo, hi, ll_grad)
    # analytic gradient against central differences
    h = 1e-6
    for c in range(C):
        dz = np.zeros(C); dz[c] = h
        num = (hmc_numpy.log_density_and_grad(z + dz, kind, lo, hi, ll_grad)[0]
               - hmc_numpy.log_density_and_grad(z - dz, kind, lo, hi, ll_grad)[0]) / (2 * h)
        assert np.allclose(num, g[:, c], rtol=1e-5, atol=1e-6)
    errs = []
    for eps in (0.02, 0.01):
        z1, p1, g1, ld1, _ = hmc_numpy.leapfrog(z, p, g, eps, im, int(0.4 / eps), kind, lo, hi, ll_grad)
        errs.append(np.abs((-ld1 + 0.5 * (p1 ** 2).sum(1)) - (-ld + 0.5 * (p ** 2).sum(1))).max())
    assert errs[1] < errs[0] / 3.0                       # second-order integrator
    zb, pb, _, _, _ = hmc_numpy.leapfrog(z1, -p1, g1, 0.01, im, 40, kind, lo, hi, ll_grad)
    assert np.allclose(zb, z, atol=1e-9) and np.allclose(-pb, p, atol=1e-9)


def test_mcmc_host_pieces():
    from jaxili_b200.posterior.mcmc_posterior import _DualAveraging, _warmup_windows, MCMCPosterior
    from jaxili_b200.priors import Uniform, Normal
    assert _warmup_windows(10) == (0, [])
    init, ends = _warmup_windows(500)
    assert init == 75 and ends == [100, 150, 250, 450]
    init, ends = _warmup_windows(100)
    assert init == 15 and ends == [90]
    # dual averaging drives a monotone acceptance curve to the target
    da, eps = _DualAveraging(1.0, 0.8), 1.0
    for _ in range(400):
        eps = da.update(float(np.exp(-eps)))             # accept(eps) = exp(-eps): target at eps = -ln 0.8
    assert abs(da.final() + np.log(0.8)) < 0.02
    u = Uniform(-np.ones(3, np.float32), 2 * np.ones(3, np.float32))
    t = torch.tensor([[0.0, 0.5, 1.9], [0.0, 2.5, 0.0]])
    lp = u.log_prob(t)
    assert lp.shape == (2, 3) and torch.allclose(lp[0], torch.full((3,), -np.log(3.0))) and lp[1, 1] == -np.inf
    assert u.sample(key=1, sample_shape=(7,), device="cpu").shape == (7, 3)
    nrm = Normal(np.zeros(2, np.float32), 2.0)
    assert torch.allclose(nrm.log_prob(torch.zeros(1, 2)), torch.full((1, 2), -np.log(2.0) - 0.5 * np.log(2 * np.pi)))
    k, lo, hi = nrm.bounds()
    assert k.tolist() == [0, 0] and hi.tolist() == [2.0, 2.0]
    with pytest.raises(NotImplementedError):
        MCMCPosterior(None, None, u, mcmc_method="wrong_method")
    post = MCMCPosterior(None, None, u)
    assert post.mcmc_method == "nuts_numpyro" and post.mcmc_kwargs == {}
    assert post.log_prior(t).shape == (2,)
    with pytest.raises(ValueError):
        post.sample(10)


def test_utils_like_reference():
    """jaxili/tests/test_utils.py:6-49 against the shims (the realnvp / mdn hyper-parameter checks belong to
    density estimators outside this path)."""
    from jaxili_b200 import nn
    from jaxili_b200.utils import check_density_estimator, check_hparams_maf, validate_theta_x
    check_density_estimator("maf")
    check_density_estimator("mdn")
    check_density_estimator("realnvp")
    with pytest.raises(ValueError):
        check_density_estimator("blablabla")
    rng = np.random.default_rng(0)
    theta = rng.standard_normal((100, 2)).astype(np.float32)
    x = rng.standard_normal((100, 2)).astype(np.float32)
    validate_theta_x(torch.from_numpy(theta), torch.from_numpy(x))
    _, _, batch_size = validate_theta_x(theta, x)
    assert batch_size == 100
    with pytest.raises(AssertionError):
        validate_theta_x(theta, x[1:])
    check_hparams_maf({"n_layers": 5, "layers": [50, 50, 50], "activation": nn.relu, "use_reverse": True})
    with pytest.raises(AssertionError):
        check_hparams_maf({})
    with pytest.raises(AssertionError):
        check_hparams_maf({"n_layers": 5.0, "layers": [50], "activation": nn.relu, "use_reverse": True})


_PEER_FALLBACK_WORKER = r"""
import sys, os
root, rank, world, port = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
sys.path.insert(0, root)
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port, RANK=str(rank), WORLD_SIZE=str(world))
import torch, torch.distributed as dist
dist.init_process_group("gloo", rank=rank, world_size=world)
from jaxili_b200 import parallel

class StubEngine:
    # stands in for MafEngine: rank 1 cannot map its peers (no NVLink / IPC)
    n_params = 1000
    device = torch.device("cpu")
    def __init__(self): self.calls = []
    def dp_init(self, r, w): self.calls.append("init"); return bytes(64)
    def dp_connect(self, handles):
        self.calls.append("connect")
        if rank == 1: raise RuntimeError("cudaIpcOpenMemHandle: peer access unsupported")
    def dp_reset(self): self.calls.append("reset")
    def dp_bind_step(self, s): self.calls.append("bind")

eng = StubEngine()
step = torch.zeros(1, dtype=torch.int32)
parallel.torch.cuda.synchronize = lambda *a, **k: None      # CPU stand-in
first = parallel.setup_peer_allreduce(eng, step)
second = parallel.setup_peer_allreduce(eng, step)          # remembered: no second attempt, no collective
assert first is False and second is False, (first, second)
assert eng.calls == ["init", "connect"], eng.calls
assert "reset" not in eng.calls and "bind" not in eng.calls
dist.barrier()
dist.destroy_process_group()
"""


def test_peer_mapping_failure_falls_back_on_every_rank(tmp_path):
    """If ANY rank cannot map the peers' buffers, ALL ranks must take the NCCL path together (a split decision
    would deadlock the first collective)."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "peer_worker.py"
    script.write_text(_PEER_FALLBACK_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), "2", str(port)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out.decode()[-2000:]


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The drop-in boundary is a C ABI: the header must compile as C99 (no C++ types in the signatures) and a C
    program must link against the library and call a non-compute entry."""
    hdr = os.path.join(ROOT, "include", "mafb200.h")
    r = subprocess.run(["gcc", "-x", "c", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", hdr],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include "mafb200.h"\n'
                   'int main(void) { MafDesc d; d.n_in = 0; printf("%d %d\\n", mafb200_version(), (int)sizeof(d));\n'
                   '  return mafb200_last_error() == 0; }\n')
    exe = tmp_path / "abi"
    libdir = os.path.join(ROOT, "jaxili_b200", "csrc")
    r = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", libdir, "-lmafb200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.split()[0].isdigit(), (r.returncode, r.stdout, r.stderr)


def test_kernel_resource_budgets():
    """Regression guard on the built cubin (cuobjdump, no GPU): the fused flow kernels keep their accumulators in
    registers (one 512-thread CTA per SM: <= 128 registers, no local-memory stack) and the cooperative tail kernel
    (256-thread blocks) stays within 128 registers, which is what lets all ceil(P/128) blocks be co-resident two per
    SM; the chain kernel's training pass fits 480 threads (<= 128 registers) without spilling."""
    import shutil
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(tool):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([tool, "-res-usage", _lib.LIB_PATH], capture_output=True, text=True).stdout
    usage = {}
    name = None
    for line in out.splitlines():
        line = line.strip()
        if line.startswith("Function "):
            name = line[len("Function "):].rstrip(":")
        elif line.startswith("REG:") and name:
            usage[name] = {k: int(v) for k, v in (f.split(":") for f in line.split() if f.split(":")[1].isdigit())}
    flow512 = {n: u for n, u in usage.items() if "maf_flow_kernel" in n and "ELi512E" in n}
    assert len(flow512) >= 8
    for n, u in flow512.items():
        assert u["REG"] <= 128 and u["STACK"] == 0, (n, u)
    tails = {n: u for n, u in usage.items() if "step_tail_kernel" in n}
    assert len(tails) == 2 and all(u["REG"] <= 128 for u in tails.values()), tails
    chains = {n: u for n, u in usage.items() if "maf_chain_kernelILb1E" in n}
    assert len(chains) >= 4 and all(u["REG"] <= 128 for u in chains.values()), chains
    samplers = {n: u for n, u in usage.items() if "maf_sample_kernel" in n}
    assert samplers and all(u["STACK"] == 0 for u in samplers.values()), samplers


def test_sass_shows_blackwell_native_paths():
    """SASS of the built library (cuobjdump, no GPU): the wide GEMM core is tcgen05 (UTCHMMA tiles fed by UTMALDG,
    accumulators read back with LDTM), the fused flow kernel stages its operands with bulk copies (UBLKCP) behind
    mbarriers (SYNCS) and runs packed FFMA2 in its micro-GEMMs; no legacy tensor-core path (HMMA) anywhere."""
    import shutil
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(tool):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([tool, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    per_fn, name = {}, None
    for line in sass.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
            per_fn[name] = set()
        elif name:
            for m in ("UTCHMMA", "UTMALDG", "LDTM", "UBLKCP", "SYNCS", "FFMA2", "HMMA"):
                if m in line and not (m == "HMMA" and "UTCHMMA" in line):
                    per_fn[name].add(m)
    tc = [s for n, s in per_fn.items() if "wide_tc_gemm_kernel" in n]
    assert tc and all({"UTCHMMA", "UTMALDG", "LDTM"} <= s for s in tc)
    flow = [s for n, s in per_fn.items() if "maf_flow_kernel" in n]
    assert flow and all({"UBLKCP", "SYNCS", "FFMA2"} <= s for s in flow)
    assert not any("HMMA" in s for s in per_fn.values())


def test_hmc_restatement_samples_the_right_distribution():
    """The restated transition (leapfrog in the unconstrained space + Metropolis) targets prior x likelihood: a
    Gaussian likelihood under a Uniform(-1, 2) prior and a Normal prior, moments against quadrature."""
    from oracle import hmc_numpy
    rng = np.random.default_rng(3)
    n, C = 512, 2
    mu, sig = np.array([0.6, -0.3]), np.array([0.8, 0.5])

    def ll_grad(theta):
        d = (theta - mu) / sig
        return -0.5 * (d * d).sum(1), -d / sig

    kind = np.array([1, 0]); lo = np.array([-1.0, 0.0]); hi = np.array([2.0, 1.0])   # U(-1, 2) x N(0, 1)
    z = rng.normal(size=(n, C)) * 0.1
    ld, g, th = hmc_numpy.log_density_and_grad(z, kind, lo, hi, ll_grad)
    eps, im, keep = 0.25, np.ones(C), []
    for it in range(160):
        p = rng.normal(size=(n, C))
        z1, p1, g1, ld1, th1 = hmc_numpy.leapfrog(z, p, g, eps, im, int(rng.integers(3, 9)), kind, lo, hi, ll_grad)
        dh = (ld1 - 0.5 * (p1 ** 2).sum(1)) - (ld - 0.5 * (p ** 2).sum(1))
        acc = np.log(rng.uniform(size=n)) < dh
        z, g, ld, th = (np.where(acc[:, None], z1, z), np.where(acc[:, None], g1, g), np.where(acc, ld1, ld),
                        np.where(acc[:, None], th1, th))
        if it >= 60:
            keep.append(th.copy())
    s = np.concatenate(keep)
    # dimension 0: N(0.6, 0.8^2) truncated to (-1, 2); dimension 1: N(-0.3, 0.5^2) x N(0, 1) = N(-0.24, 0.2)
    t = np.linspace(-1, 2, 20001)
    w = np.exp(-0.5 * ((t - mu[0]) / sig[0]) ** 2)
    m0 = (t * w).sum() / w.sum()
    v0 = ((t - m0) ** 2 * w).sum() / w.sum()
    v1 = 1.0 / (1.0 / sig[1] ** 2 + 1.0)
    m1 = v1 * mu[1] / sig[1] ** 2
    assert s[:, 0].min() > -1 and s[:, 0].max() < 2
    assert abs(s[:, 0].mean() - m0) < 0.03 and abs(s[:, 0].var() - v0) < 0.03
    assert abs(s[:, 1].mean() - m1) < 0.02 and abs(s[:, 1].var() - v1) < 0.02


def test_bench_graph_steps_divide_the_timed_steps(monkeypatch):
    """bench.py times EXACTLY --steps steps: the number of steps captured per graph launch must divide it."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    monkeypatch.delenv("MAFB200_BENCH_GRAPH_STEPS", raising=False)
    for K in (1, 7, 20, 50, 64, 300, 5000, 4999):
        s = bench.graph_steps(K)
        assert 1 <= s <= 25 and K % s == 0, (K, s)
    assert bench.graph_steps(20) == 20 and bench.graph_steps(5000) == 25 and bench.graph_steps(7) == 1
    monkeypatch.setenv("MAFB200_BENCH_GRAPH_STEPS", "4")
    assert bench.graph_steps(20) == 4
