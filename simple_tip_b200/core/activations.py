"""Activation-trace producer for PyTorch models and the reference's on-disk badge layout
(SURVEY.md §8 f4).

The reference obtains activation traces from a multi-output Keras model
(`handler_model.BaseModel.get_activations / walk_activations`, handler_model.py:175-206) as host
NumPy arrays and can persist them in badges of 100 inputs
(`activation_persistor.py:13-35`: `<root>/activations/<case_study>/model_<id>/<dataset>/layer_<i>/badge_<b>.npy`,
labels in `.../labels/badge_<b>.npy`).  Here the same two things for a `torch.nn.Module`:

* `TransparentModel` returns the outputs of selected layers for a batch through forward hooks.  The
  tensors stay on the model's device: handed to `DSA` / `LSA` / `KMNC.buckets` / `DeepGini.calculate`
  they are scored without a host round trip (those entry points accept CUDA tensors).
* `persist_badges` / `load_badges` write and read the reference's `.npy` layout, so traces produced
  here can be consumed by the reference's scripts and vice versa.
"""
from __future__ import annotations

import os
from typing import Dict, Generator, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

BADGE_SIZE = 100      # activation_persistor.py:11

LayerId = Union[int, str]


class TransparentModel:
    """Selected layer outputs of a torch module, like the reference's transparent Keras model
    (handler_model.py:193-206: the outputs of `model.layers[i]` for i in `activation_layers`,
    plus the model output when `include_last_layer`).

    activation_layers: indices into `list(model.children())` (what Keras' `model.layers` is for a
    Sequential) or dotted submodule names from `model.named_modules()`."""

    def __init__(self, model, activation_layers: Optional[Sequence[LayerId]], include_last_layer: bool = True):
        if activation_layers is None:
            raise ValueError("No activation layers specified")          # handler_model.py:196-197
        self.model = model
        self.include_last_layer = include_last_layer
        children = list(model.children())
        named = dict(model.named_modules())
        self._modules = []
        for lid in activation_layers:
            if isinstance(lid, (int, np.integer)):
                if not 0 <= int(lid) < len(children):
                    raise IndexError(f"layer index {lid} out of range (model has {len(children)} layers)")
                self._modules.append(children[int(lid)])
            else:
                if lid not in named:
                    raise KeyError(f"no submodule named {lid!r}")
                self._modules.append(named[lid])

    def get_activations(self, x) -> List["torch.Tensor"]:     # noqa: F821
        """One deterministic forward pass (eval mode, no grad); returns one tensor per selected layer
        (+ the model output), each [batch, ...], on the model's device."""
        import torch

        captured: Dict[int, torch.Tensor] = {}
        hooks = [m.register_forward_hook(lambda _m, _i, out, k=k: captured.__setitem__(k, out.detach()))
                 for k, m in enumerate(self._modules)]
        was_training = self.model.training
        self.model.eval()
        try:
            with torch.no_grad():
                out = self.model(x)
        finally:
            for h in hooks:
                h.remove()
            self.model.train(was_training)
        missing = [k for k in range(len(self._modules)) if k not in captured]
        if missing:
            raise RuntimeError(f"selected layers {missing} did not run in the forward pass")
        acts = [captured[k] for k in range(len(self._modules))]
        if self.include_last_layer:
            acts.append(out.detach())
        return acts

    def walk_activations(self, batches: Iterable) -> Generator[List["torch.Tensor"], None, None]:   # noqa: F821
        """Activations badge by badge for a (potentially large) dataset (handler_model.py:175-180)."""
        for badge in batches:
            yield self.get_activations(badge)

    def collect(self, x, batch_size: int = BADGE_SIZE) -> List["torch.Tensor"]:   # noqa: F821
        """All of x in badges of `batch_size`; per layer one [N, ...] tensor on the model's device."""
        import torch

        per_layer: Optional[List[List[torch.Tensor]]] = None
        for start in range(0, x.shape[0], batch_size):
            acts = self.get_activations(x[start:start + batch_size])
            if per_layer is None:
                per_layer = [[] for _ in acts]
            for lst, a in zip(per_layer, acts):
                lst.append(a)
        if per_layer is None:
            return []
        return [torch.cat(lst, dim=0) for lst in per_layer]


def _badge_dir(root: str, case_study: str, model_id: int, dataset: str) -> str:
    return os.path.join(root, "activations", case_study, f"model_{model_id}", dataset)


def persist_badge(root: str, case_study: str, model_id: int, dataset: str, badge_id: int, activations: Sequence,
                  labels) -> None:
    """One badge in the reference's layout (activation_persistor.py:13-35; `root` is "/assets" there)."""
    path = _badge_dir(root, case_study, model_id, dataset)
    for layer_i, layer_at in enumerate(activations):
        folder = os.path.join(path, f"layer_{layer_i}")
        os.makedirs(folder, exist_ok=True)
        np.save(os.path.join(folder, f"badge_{badge_id}.npy"), _to_numpy(layer_at))
    folder = os.path.join(path, "labels")
    os.makedirs(folder, exist_ok=True)
    np.save(os.path.join(folder, f"badge_{badge_id}.npy"), _to_numpy(labels))


def persist_badges(root: str, case_study: str, model_id: int, dataset: str, transparent_model: TransparentModel, x, y,
                   badge_size: int = BADGE_SIZE) -> int:
    """Walks (x, y) in badges and persists every layer's activations; returns the number of badges
    (activation_persistor.py:52-72)."""
    n_badges = 0
    for badge_id, start in enumerate(range(0, x.shape[0], badge_size)):
        acts = transparent_model.get_activations(x[start:start + badge_size])
        persist_badge(root, case_study, model_id, dataset, badge_id, acts, y[start:start + badge_size])
        n_badges += 1
    return n_badges


def load_badges(root: str, case_study: str, model_id: int, dataset: str, layers: Optional[Sequence[int]] = None
                ) -> Tuple[List[np.ndarray], np.ndarray]:
    """Reads a persisted dataset back: (one [N, ...] array per layer, labels [N]), badges in order."""
    path = _badge_dir(root, case_study, model_id, dataset)
    if layers is None:
        layers = sorted(int(f.split("_")[1]) for f in os.listdir(path) if f.startswith("layer_"))
    label_dir = os.path.join(path, "labels")
    badge_ids = sorted(int(f[len("badge_"):-len(".npy")]) for f in os.listdir(label_dir) if f.endswith(".npy"))
    labels = np.concatenate([np.load(os.path.join(label_dir, f"badge_{b}.npy")) for b in badge_ids])
    acts = [np.concatenate([np.load(os.path.join(path, f"layer_{layer}", f"badge_{b}.npy")) for b in badge_ids])
            for layer in layers]
    return acts, labels


def _to_numpy(a) -> np.ndarray:
    if isinstance(a, np.ndarray):
        return a
    try:
        import torch

        if isinstance(a, torch.Tensor):
            return a.detach().cpu().numpy()
    except ImportError:      # pragma: no cover
        pass
    return np.asarray(a)
