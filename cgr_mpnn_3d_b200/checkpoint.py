"""Checkpoint compatibility with the reference.

The reference saves the WHOLE module (``torch.save(self.model, path)``, training/trainer.py:208) and reloads it with
``torch.load(path, map_location=...)`` (test.py:93, cli_tool/activation_energy_predictor.py:62).  Such a pickle names
the classes ``cgr_mpnn_3D.models.GNN.GNN`` / ``DMPNNConv`` -- which resolve to this repository's drop-in classes -- but
a model trained with the real reference also carries torch_geometric objects inside every ``DMPNNConv``
(``MessagePassing`` internals) and ``pooling_fn = torch_geometric.nn.global_add_pool``.  torch_geometric is not needed
to run this implementation, so :func:`load_reference_checkpoint` unpickles with a ``find_class`` that substitutes inert
stand-ins for anything under ``torch_geometric`` (or any other module that cannot be imported), then rebuilds a clean
B200 ``GNN`` from the recovered hyper-parameters and ``state_dict``.

The other direction needs nothing: ``torch.save(model, path)`` of the drop-in module pickles as
``cgr_mpnn_3D.models.GNN.GNN`` and loads at the reference's call sites (device caches are dropped in ``__getstate__``).
"""
from __future__ import annotations

import importlib
import pickle
from typing import Any, Dict

import torch


class _Inert:
    """Stand-in for an object of a class that cannot be imported: accepts any construction and any state."""

    def __init__(self, *args, **kwargs):
        pass

    def __new__(cls, *args, **kwargs):
        return object.__new__(cls)

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        elif isinstance(state, tuple) and len(state) == 2:          # (dict, slots) protocol
            for part in state:
                if isinstance(part, dict):
                    self.__dict__.update(part)

    def __call__(self, *args, **kwargs):
        raise RuntimeError("this object is a stand-in for a class that is not installed "
                           f"({type(self).__module__}.{type(self).__qualname__})")


_STUBS: Dict[str, type] = {}


def _stub_class(module: str, name: str) -> type:
    key = f"{module}.{name}"
    if key not in _STUBS:
        _STUBS[key] = type(name.split(".")[-1], (_Inert,), {"__module__": module, "__qualname__": name})
    return _STUBS[key]


class _ReferenceUnpickler(pickle.Unpickler):
    def find_class(self, module: str, name: str) -> Any:
        if module.split(".")[0] == "torch_geometric":
            if name == "global_add_pool":                       # GNN.pooling_fn (GNN.py:23,49)
                from .model import global_add_pool
                return global_add_pool
            return _stub_class(module, name)
        try:
            return super().find_class(module, name)
        except (ImportError, AttributeError):
            try:
                importlib.import_module(module.split(".")[0])
            except ImportError:
                return _stub_class(module, name)                 # whole package missing (e.g. torch_scatter)
            raise


class _PickleModule:
    """The ``pickle_module`` object torch.load expects (it subclasses ``Unpickler``)."""
    __name__ = "pickle"
    Unpickler = _ReferenceUnpickler
    load = staticmethod(lambda f, **kw: _ReferenceUnpickler(f, **kw).load())
    loads = staticmethod(pickle.loads)
    dump = staticmethod(pickle.dump)
    dumps = staticmethod(pickle.dumps)
    HIGHEST_PROTOCOL = pickle.HIGHEST_PROTOCOL
    PickleError = pickle.PickleError
    UnpicklingError = pickle.UnpicklingError


def _arch_from_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, Any]:
    """Hyper-parameters implied by the parameter shapes (state_dict layout of GNN.py:53-74)."""
    depth = 0
    while f"convs.{depth}.lin.weight" in sd:
        depth += 1
    if depth == 0 or "edge_init.weight" not in sd or "edge_to_node.weight" not in sd:
        raise ValueError("not a CGR-MPNN-3D GNN state_dict (edge_init / convs / edge_to_node missing)")
    hidden = int(sd["edge_init.weight"].shape[0])
    fa = int(sd["edge_to_node.weight"].shape[1]) - hidden
    fb = int(sd["edge_init.weight"].shape[1]) - fa
    return dict(num_node_features=fa, num_edge_features=fb, depth=depth, hidden_sizes=[hidden] * depth,
                use_learnable_skip="skip_weights.0" in sd)


def load_reference_checkpoint(path, map_location="cpu", **gnn_kwargs):
    """Load a checkpoint written by the reference -- a whole-module pickle (trainer.py:208) or a plain state_dict --
    into a fresh B200 ``GNN``.  ``gnn_kwargs`` override what cannot be read from a bare state_dict
    (``dropout_ps``, ``activation_fn``)."""
    from .model import GNN
    obj = torch.load(path, map_location=map_location, pickle_module=_PickleModule, weights_only=False)
    if isinstance(obj, dict):
        sd = obj.get("state_dict", obj) if not all(isinstance(v, torch.Tensor) for v in obj.values()) else obj
        kw = _arch_from_state_dict(sd)
    else:
        sd = {k: v.detach() for k, v in obj.state_dict().items()}
        kw = _arch_from_state_dict(sd)
        for name in ("hidden_sizes", "dropout_ps", "activation_fn", "use_learnable_skip", "depth"):
            if name in obj.__dict__:
                kw[name] = obj.__dict__[name]
    kw.update(gnn_kwargs)
    model = GNN(**kw)
    model.load_state_dict(sd, strict=True)
    dev = torch.device(map_location) if isinstance(map_location, (str, torch.device)) else None
    return model.to(dev) if dev is not None else model
