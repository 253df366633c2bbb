#!/usr/bin/env python
"""Golden fixtures from the REAL jaxili (run this on a machine that has jax, flax, optax, distrax and jaxili).

    pip install jaxili            # or: pip install -e /path/to/sachaguer/jaxili
    python tools/jaxili_golden.py [--out tests/golden]

This image cannot run it (no JAX wheels, no network: SURVEY F2), which is why parity with jaxili itself is
"unpinned" in DESIGN.md.  The script turns that into "pinnable by one command elsewhere": it drives jaxili's own
classes the way its tests and trainer do and writes flax-layout ``.npz`` files that
``tests/test_jaxili_golden.py`` picks up when they are present under tests/golden/jaxili_*.npz and compares against
BOTH the CPU oracle (oracle/maf_numpy.py) and the CUDA path (through jaxili_b200.layout).

What is recorded per case (all float32 unless noted):
  hparams                       n_in, n_cond, n_layers, layers, activation name, use_reverse, seed
  params/<flax path>            every leaf of ``variables["params"]`` (kernel (in,out), bias), flax path as the key
  x, y                          inputs (density variable, conditioning variable)
  shift, scale                  the distrax.ScalarAffine of NDE_w_Standardization (cases with the wrapper)
  log_prob                      model.apply(variables, x, y, method="log_prob")      jaxili/model.py:903-921, 1073-1098
  u, log_det                    model.apply(variables, x, y)  (ConditionalMAF.__call__)                     848-874
  x_rec, log_det_rec            model.apply(variables, u, y, method="backward")                       876-901, 745-773
  loss, grads/<flax path>       jax.value_and_grad of loss_nll_npe                       jaxili/loss.py:35-58
  after3/<flax path>, losses3   parameters and losses after three train steps with the optimiser chain of
                                TrainerModule.init_optimizer (jaxili/train.py:246-300) and train_step (330-335)
  opt                           the optimiser hyper-parameters used (lr, warmup, decay_steps, gradient_clip)

Nothing here is imported by the product; it is test tooling like oracle/.
"""
import argparse
import os

import numpy as np


def flat_paths(tree, prefix):
    from flax.traverse_util import flatten_dict
    return {prefix + "/".join(k): np.asarray(v) for k, v in flatten_dict(tree).items()}


def optimizer_like_trainer(optax, lr, warmup, decay_steps, gradient_clip):
    """The chain TrainerModule.init_optimizer builds for optimizer_name="adam" (jaxili/train.py:257-300)."""
    sched = optax.warmup_cosine_decay_schedule(init_value=0.01 * lr, peak_value=lr, warmup_steps=warmup,
                                               decay_steps=decay_steps, end_value=0.01 * lr)
    return optax.chain(optax.clip_by_global_norm(gradient_clip), optax.adam(sched))


def run_case(tag, out_dir, n_in, n_cond, n_layers, layers, act_name, use_reverse, B, wrapper, seed_data):
    import distrax
    import jax
    import jax.numpy as jnp
    import optax
    from jaxili.loss import loss_nll_npe
    from jaxili.model import ConditionalMAF, Identity, NDE_w_Standardization

    act = getattr(jax.nn, act_name)
    rng = np.random.default_rng(seed_data)
    x = rng.standard_normal((B, n_in)).astype(np.float32)
    y = rng.standard_normal((B, n_cond)).astype(np.float32)
    maf = ConditionalMAF(n_in, n_cond, n_layers, list(layers), use_reverse=use_reverse, seed=42, activation=act)
    rec = dict(n_in=n_in, n_cond=n_cond, n_layers=n_layers, layers=np.array(layers), activation=act_name,
               use_reverse=use_reverse, seed=42, x=x, y=y, wrapper=wrapper)
    if wrapper:
        shift = (rng.standard_normal(n_in) * 2).astype(np.float32)
        scale = rng.uniform(0.5, 2.0, n_in).astype(np.float32)
        model = NDE_w_Standardization(nde=maf, embedding_net=Identity(),
                                      transformation=distrax.ScalarAffine(scale=scale, shift=shift))
        rec.update(shift=shift, scale=scale)
        x = (x * scale + shift).astype(np.float32)          # data-space inputs
        rec["x"] = x
    else:
        model = maf
    variables = model.init(jax.random.PRNGKey(0), jnp.asarray(x), jnp.asarray(y))
    # trained-scale weights exercise the exp / clamp paths: scale the flax init (0.01-ish) up, deterministically
    variables = jax.tree_util.tree_map(lambda a: a * 8.0, variables)
    params = variables["params"]
    rec.update(flat_paths(params, "params/"))
    rec["log_prob"] = np.asarray(model.apply(variables, jnp.asarray(x), jnp.asarray(y), method="log_prob"))
    if not wrapper:
        u, ld = model.apply(variables, jnp.asarray(x), jnp.asarray(y))
        xr, ldr = model.apply(variables, u, jnp.asarray(y), method="backward")
        rec.update(u=np.asarray(u), log_det=np.asarray(ld), x_rec=np.asarray(xr), log_det_rec=np.asarray(ldr))
    batch = [jnp.asarray(x), jnp.asarray(y)]                 # [theta, x] as numpy_collate yields it
    loss, grads = jax.value_and_grad(lambda p: loss_nll_npe(model, p, batch))(params)
    rec["loss"] = np.float32(loss)
    rec.update(flat_paths(grads, "grads/"))
    opt = dict(lr=5e-4, warmup=0.1, decay_steps=1_000_000, gradient_clip=5.0)
    tx = optimizer_like_trainer(optax, **opt)
    state = tx.init(params)
    p = params
    losses = []
    for _ in range(3):                                       # train_step: value_and_grad + apply_gradients
        l, g = jax.value_and_grad(lambda q: loss_nll_npe(model, q, batch))(p)
        upd, state = tx.update(g, state, p)
        p = optax.apply_updates(p, upd)
        losses.append(float(l))
    rec.update(flat_paths(p, "after3/"))
    rec["losses3"] = np.array(losses, np.float32)
    rec["opt"] = np.array([opt["lr"], opt["warmup"], opt["decay_steps"], opt["gradient_clip"]], np.float64)
    path = os.path.join(out_dir, f"jaxili_{tag}.npz")
    np.savez_compressed(path, **rec)
    print("wrote", path, "loss", float(loss))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                  "tests", "golden"))
    args = ap.parse_args()
    import jax
    jax.config.update("jax_platform_name", "cpu")
    os.makedirs(args.out, exist_ok=True)
    # jaxili/tests/test_nn.py:10-61 (ConditionalMAF, silu, 3 and 4 layers, with / without reverse)
    run_case("maf_silu_rev", args.out, 3, 5, 3, (128, 128), "silu", True, 10, False, 1)
    run_case("maf_silu_norev", args.out, 3, 5, 4, (128, 128), "silu", False, 10, False, 2)
    # jaxili/tests/test_nn.py:182-237 (NDE_w_Standardization around a [50,50] relu MAF)
    run_case("nde_relu", args.out, 2, 5, 3, (50, 50), "relu", True, 10, True, 3)
    # the default NPE network of jaxili/inference/npe.py:39-45 at the c2 shape
    run_case("c2_relu", args.out, 10, 10, 5, (50, 50), "relu", True, 96, True, 4)


if __name__ == "__main__":
    main()
