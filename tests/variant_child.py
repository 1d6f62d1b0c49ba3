"""Child process of tests/test_gpu_variants.py: one inference forward and one training forward + backward of a fixed
model / batch, written to an .npz.  The parent sets the environment switch under test (the library reads its
experiment switches once per process) and compares against its own default-configuration results."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tests.util import build_model
from cgr_mpnn_3d_b200.data import make_batch


def run(path=None):
    meta = dict(fa=846, fb=14, depth=3, hidden=400, skip=True, wseed=5, act="relu")
    data = make_batch(256, seed=31, kind="t1x", fa=846).to("cuda")
    model = build_model(meta, engine="auto").eval()
    model.tile_policy = "throughput"          # wide-slice regime: the persistent atom projection and the 2-tile clusters
    with torch.no_grad():
        out = model(data)
    model.train()
    model.zero_grad(set_to_none=True)
    out_t = model(data)
    torch.nn.functional.mse_loss(out_t, data.y, reduction="sum").backward()
    grads = torch.cat([q.grad.reshape(-1) for q in model.parameters()])
    res = {"out": out.cpu().numpy(), "out_t": out_t.detach().cpu().numpy(), "grads": grads.cpu().numpy()}
    if path:
        np.savez(path, **res)
    return res


if __name__ == "__main__":
    run(sys.argv[1])
