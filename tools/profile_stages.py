"""Per-stage CUDA-event breakdown (cgr_profile_*) of a forward or a training step."""
import argparse, collections, ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from cgr_mpnn_3D.models.GNN import GNN
from cgr_mpnn_3d_b200 import _lib
from cgr_mpnn_3d_b200.data import make_batch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64); ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--train", action="store_true"); ap.add_argument("--engine", default="auto")
a = ap.parse_args()
torch.manual_seed(0)
m = GNN(846, 14, depth=4, hidden_sizes=[400] * 4, dropout_ps=[0.0] * 4, activation_fn=F.relu, use_learnable_skip=True).cuda()
m.engine = a.engine
d = make_batch(a.batch, seed=0, fa=846).to("cuda")
lib = _lib.load()

def step():
    if a.train:
        m.zero_grad(set_to_none=True)
        ((m(d) - d.y) ** 2).sum().backward()
    else:
        with torch.no_grad():
            m(d)
m.train(a.train)
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(a.iters): step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / a.iters
lib.cgr_profile_enable(1)
for _ in range(a.iters): step()
torch.cuda.synchronize()
tot, cnt = collections.OrderedDict(), collections.OrderedDict()
name = ctypes.create_string_buffer(64); ms = ctypes.c_float()
for i in range(lib.cgr_profile_count()):
    if lib.cgr_profile_get(i, name, 64, ctypes.byref(ms)) == 0:
        k = name.value.decode(); tot[k] = tot.get(k, 0) + ms.value; cnt[k] = cnt.get(k, 0) + 1
lib.cgr_profile_enable(0)
print(f"wall per step {wall*1e6:.1f} us; sum of ranges per step {sum(tot.values())/a.iters*1e3:.1f} us")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"  {k:22s} {cnt[k]/a.iters:5.1f} launches/step  {tot[k]/cnt[k]*1e3:8.1f} us avg  {tot[k]/a.iters*1e3:9.1f} us/step")
