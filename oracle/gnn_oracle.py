"""ORACLE — test infrastructure only, never imported by the product path.

CPU restatement (pure torch, no torch_geometric) of the reference hot path
``cgr_mpnn_3D/models/GNN.py``.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.

Pinning status: the reference's own tests hold NO numeric vector for this path
(``tests/test_trainer.py:37-38`` only constructs ``GNN``), so the
oracle is pinned instead against outputs of the *unmodified* reference
``GNN.py`` executed in the build container under a test-only torch_geometric
stand-in (``tests/_pyg_shim``); the generating script is
``tests/golden/make_golden.py`` and the committed vectors live in
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks this file
against them bit-for-bit.

Each function cites the reference lines it follows.  PyG semantics restated
(recalled from PyG 2.6):

* ``MessagePassing.propagate(edge_index, x=None, edge_attr=h)`` with
  ``aggr="add"``: ``zeros(max(edge_index[1])+1, H).scatter_add_(0, dst, h)``.
* ``global_add_pool(h, batch)``: ``batch is None`` -> ``h.sum(0, keepdim=True)``,
  else scatter-sum with ``dim_size = batch.max()+1``.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


def propagate_add(edge_index: torch.Tensor, h: torch.Tensor) -> torch.Tensor:
    """reference GNN.py:134 + 143-145 (message returns edge_attr, sum-aggregated at dst)."""
    dst = edge_index[1]
    n_out = int(dst.max()) + 1          # PyG: size=None, x=None -> dim_size = max+1
    out = h.new_zeros((n_out, h.shape[1]))
    # sequential fp32 accumulation in ascending bond id (verified == scatter_add_)
    out.index_add_(0, dst, h)
    return out


def global_add_pool(h: torch.Tensor, batch: Optional[torch.Tensor]) -> torch.Tensor:
    """reference GNN.py:110 (pooling_fn)."""
    if batch is None:
        return h.sum(dim=-2, keepdim=True)
    nb = int(batch.max()) + 1
    out = h.new_zeros((nb, h.shape[1]))
    out.index_add_(0, batch, h)
    return out


class OracleDMPNNConv(nn.Module):
    """reference GNN.py:113-145."""

    def __init__(self, hidden_size: int):
        super().__init__()
        self.lin = nn.Linear(hidden_size, hidden_size)      # GNN.py:129

    def forward(self, edge_index, edge_attr):
        row = edge_index[0]                                  # GNN.py:132
        a_message = propagate_add(edge_index, edge_attr)     # GNN.py:134
        e = edge_attr.size(0)
        rev_message = torch.flip(edge_attr.view(e // 2, 2, -1), dims=[1]).view(e, -1)  # GNN.py:136-138
        return a_message, self.lin(a_message[row] - rev_message)                        # GNN.py:141


class OracleGNN(nn.Module):
    """reference GNN.py:8-110; same parameter tree and registration order."""

    def __init__(self, num_node_features: int, num_edge_features: int, depth: int = 3,
                 hidden_sizes: Optional[List[int]] = None, dropout_ps: Optional[List[float]] = None,
                 activation_fn=F.relu, use_learnable_skip: bool = False):
        super().__init__()
        self.depth = depth                                               # GNN.py:45
        self.hidden_sizes = hidden_sizes or [300] * depth                # GNN.py:46
        self.dropout_ps = dropout_ps or [0.02] * depth                   # GNN.py:47
        self.activation_fn = activation_fn
        self.use_learnable_skip = use_learnable_skip
        self.edge_init = nn.Linear(num_node_features + num_edge_features, self.hidden_sizes[0])  # GNN.py:53-55
        self.convs = nn.ModuleList()
        for i in range(self.depth):                                      # GNN.py:58-60
            self.convs.append(OracleDMPNNConv(self.hidden_sizes[i]))
        self.edge_to_node = nn.Linear(num_node_features + self.hidden_sizes[-1], self.hidden_sizes[-1])  # GNN.py:63-65
        self.ffn = nn.Linear(self.hidden_sizes[-1], 1)                   # GNN.py:68
        if self.use_learnable_skip:                                      # GNN.py:71-74
            self.skip_weights = nn.ParameterList(
                [nn.Parameter(torch.tensor(1.0)) for _ in range(self.depth)])

    def forward(self, data, dropout_masks: Optional[List[torch.Tensor]] = None):
        """``dropout_masks[l]``: optional externally supplied keep-masks (bool [E,H]) so a
        device dropout stream can be replayed exactly; default is torch's own ``F.dropout``."""
        x, edge_index, edge_attr, batch = data.x, data.edge_index, data.edge_attr, data.batch
        row = edge_index[0]                                                               # GNN.py:85
        h_0 = self.activation_fn(self.edge_init(torch.cat([x[row], edge_attr], dim=1)))  # GNN.py:86
        h = h_0
        for l in range(self.depth):                                                       # GNN.py:90
            _, h = self.convs[l](edge_index, h)                                           # GNN.py:91
            if self.use_learnable_skip:                                                   # GNN.py:94-97
                h = h + self.skip_weights[l] * h_0
            else:
                h = h + h_0
            h = self.activation_fn(h)
            if dropout_masks is not None and self.training and self.dropout_ps[l] > 0:
                h = h * dropout_masks[l].to(h.dtype) / (1.0 - self.dropout_ps[l])
            else:
                h = F.dropout(h, self.dropout_ps[l], training=self.training)              # GNN.py:100-102
        s = propagate_add(edge_index, h)                                                  # GNN.py:105 (conv's lin output discarded)
        q = torch.cat([x, s], dim=1)                                                      # GNN.py:106
        h = self.activation_fn(self.edge_to_node(q))                                      # GNN.py:107
        return self.ffn(global_add_pool(h, batch)).squeeze(-1)                            # GNN.py:110


def mse_sum_loss(pred: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """reference train.py:120 (``MSELoss(reduction="sum")``), trainer.py:142."""
    return ((pred - y) ** 2).sum()


def scale_normalised_error(out: torch.Tensor, ref: torch.Tensor) -> float:
    """SURVEY.md §8c parity metric: ``max_b |out-ref| / max(|ref_b|, mean|ref|)``."""
    out = out.detach().double().flatten().cpu()
    ref = ref.detach().double().flatten().cpu()
    if ref.numel() == 0:
        return 0.0
    denom = torch.maximum(ref.abs(), ref.abs().mean().clamp_min(1e-30))
    return float(((out - ref).abs() / denom).max())


def tensor_error(out: torch.Tensor, ref: torch.Tensor) -> float:
    """Gradient metric: max-abs difference normalised by the reference tensor's max-abs."""
    out = out.detach().double().cpu()
    ref = ref.detach().double().cpu()
    scale = float(ref.abs().max().clamp_min(1e-30))
    return float((out - ref).abs().max()) / scale
