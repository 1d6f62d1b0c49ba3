"""TEST-ONLY: minimal ``torch_geometric.data`` surface (see package docstring)."""


class Data:
    """Attribute container like ``tg.data.Data`` (reference ``data/ChemDataset.py:81-94`` builds one per reaction)."""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        return int(self.x.shape[0])

    @property
    def num_node_features(self):
        return int(self.x.shape[1])

    @property
    def num_edge_features(self):
        return int(self.edge_attr.shape[1])
