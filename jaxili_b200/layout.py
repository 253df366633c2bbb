"""Flat parameter buffer <-> flax-style parameter tree.

jaxili's parameters live in a flax pytree:

    params / [nde /] layer_list_{k} / ConditionalMADE_0 / layers_{0,2,4,..} / Dense_0 / {kernel, bias}

(``layers_{2i}``: the activation callables occupy the odd slots of the list in
jaxili/model.py:592-596).  The kernels work on ONE flat FP32 buffer ordered
layer-major, linear-major, kernel (in, out) row-major then bias.  ``ParamTree`` is a nested dict
whose leaves are *views* into that buffer, so converting between the two is free.
"""

from __future__ import annotations

import numpy as np
import torch


class ParamTree(dict):
    """Nested dict of tensor views with the owning flat buffer attached as ``.flat``."""

    flat: torch.Tensor = None


def layer_shapes(dims):
    return [((dims[i], dims[i + 1]), (dims[i + 1],)) for i in range(len(dims) - 1)]


def n_params(dims, n_layers):
    return n_layers * sum(a * b + b for a, b in zip(dims[:-1], dims[1:]))


def tree_from_flat(flat, dims, n_layers, wrap_nde=False):
    """Build the flax-layout tree of views over ``flat`` (torch tensor or numpy array)."""
    root = ParamTree()
    off = 0
    for k in range(n_layers):
        made = {}
        for i, (ws, bs) in enumerate(layer_shapes(dims)):
            nw = ws[0] * ws[1]
            kernel = flat[off:off + nw].reshape(ws)
            off += nw
            bias = flat[off:off + bs[0]]
            off += bs[0]
            made[f"layers_{2 * i}"] = {"Dense_0": {"kernel": kernel, "bias": bias}}
        root[f"layer_list_{k}"] = {"ConditionalMADE_0": made}
    assert off == flat.shape[0]
    if wrap_nde:
        outer = ParamTree()
        outer["nde"] = dict(root)
        outer.flat = flat
        return outer
    root.flat = flat
    return root


def flat_from_tree(tree, dims, n_layers, device=None):
    """Flatten any tree with the layout above (leaves: numpy arrays or tensors) to a new FP32
    flat tensor; returns ``tree.flat`` itself when the tree already owns one."""
    flat = getattr(tree, "flat", None)
    if flat is not None:
        return flat
    if "nde" in tree and "layer_list_0" not in tree:
        tree = tree["nde"]
    parts = []
    for k in range(n_layers):
        made = tree[f"layer_list_{k}"]["ConditionalMADE_0"]
        for i, (ws, bs) in enumerate(layer_shapes(dims)):
            leaf = made[f"layers_{2 * i}"]["Dense_0"]
            for name, shape in (("kernel", ws), ("bias", bs)):
                a = leaf[name]
                a = a.detach().to("cpu").numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
                if tuple(a.shape) != tuple(shape):
                    raise ValueError(f"layer_list_{k}/layers_{2 * i}/{name}: expected {shape}, got {a.shape}")
                parts.append(a.astype(np.float32).reshape(-1))
    out = torch.from_numpy(np.concatenate(parts))
    return out.to(device) if device is not None else out


def tree_to_numpy(tree):
    """Deep copy of a parameter tree with numpy leaves (checkpoint format)."""
    if isinstance(tree, dict):
        return {k: tree_to_numpy(v) for k, v in tree.items()}
    return tree.detach().cpu().numpy() if isinstance(tree, torch.Tensor) else np.asarray(tree)


def flatten_paths(tree, prefix=""):
    """{'a/b/c': leaf} view of a nested dict (npz checkpoint keys follow the flax path)."""
    out = {}
    for k, v in tree.items():
        p = f"{prefix}/{k}" if prefix else k
        if isinstance(v, dict):
            out.update(flatten_paths(v, p))
        else:
            out[p] = v
    return out


def unflatten_paths(flatmap):
    root = {}
    for path, v in flatmap.items():
        node = root
        parts = path.split("/")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = v
    return root
