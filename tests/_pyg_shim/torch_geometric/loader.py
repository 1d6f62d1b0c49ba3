"""TEST-ONLY: minimal ``torch_geometric.loader.DataLoader`` (see package docstring).

What the reference's trainer constructs at ``training/trainer.py:105-118``: a ``torch.utils.data.DataLoader`` whose
collate is PyG's ``Batch.from_data_list`` -- restated by ``cgr_mpnn_3d_b200.data.collate_host`` (x / edge_attr / y
concatenated, ``edge_index`` offset by the cumulative node counts, ``batch``, ``ptr``).  Items are per-reaction graphs
with numpy or tensor fields ``x, edge_index, edge_attr, y``.  Worker processes are not used (``num_workers`` is
accepted and ignored) so that the shuffled order only depends on the torch seed.
"""
import numpy as np
import torch
import torch.utils.data


def _np(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


def _collate(items):
    from cgr_mpnn_3d_b200.data import Graph, collate_host
    graphs = [Graph(x=_np(g.x).astype(np.float32), edge_index=_np(g.edge_index).astype(np.int64),
                    edge_attr=_np(g.edge_attr).astype(np.float32), y=_np(g.y).astype(np.float32).reshape(-1)[:1])
              for g in items]
    return collate_host(graphs)


class DataLoader(torch.utils.data.DataLoader):
    def __init__(self, dataset, batch_size=1, shuffle=False, num_workers=0, pin_memory=False, **kwargs):
        kwargs.pop("collate_fn", None)
        super().__init__(dataset, batch_size=batch_size, shuffle=shuffle, num_workers=0, pin_memory=False,
                         collate_fn=_collate, **kwargs)
