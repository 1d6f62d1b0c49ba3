"""TEST-ONLY: see package docstring."""
import torch


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", **kwargs):
        super().__init__()
        if aggr != "add":
            raise NotImplementedError("shim only implements aggr='add'")
        self.aggr = aggr

    def propagate(self, edge_index, size=None, **kwargs):
        # flow="source_to_target": aggregate message(...) at edge_index[1]; x=None and
        # size=None => dim_size = int(index.max()) + 1 (PyG SumAggregation -> scatter).
        msg_kwargs = {k: v for k, v in kwargs.items() if k != "x"}
        msg = self.message(**msg_kwargs)
        index = edge_index[1]
        dim_size = int(index.max()) + 1
        out = msg.new_zeros((dim_size,) + tuple(msg.shape[1:]))
        return out.scatter_add_(0, index.view(-1, 1).expand_as(msg), msg)


def global_add_pool(x, batch, size=None):
    if batch is None:
        return x.sum(dim=-2, keepdim=True)
    dim_size = int(batch.max()) + 1 if size is None else size
    out = x.new_zeros((dim_size,) + tuple(x.shape[1:]))
    return out.scatter_add_(0, batch.view(-1, 1).expand_as(x), x)
