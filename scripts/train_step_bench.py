"""Training-path timing: forward + backward of one HGTConv layer (autograd path, pyhgt_b200/autograd.py)."""
import argparse, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyhgt_b200 import HGTConv, synth

ap = argparse.ArgumentParser(); ap.add_argument("--scale", type=float, default=0.25)
ap.add_argument("--iters", type=int, default=4); ap.add_argument("--profile", type=int, default=1); a = ap.parse_args()
dev = torch.device("cuda:0")
g = synth.make_mag_shaped(a.scale)
torch.manual_seed(0)
m = HGTConv(256, 256, 4, 4, 8, 0.0, True, False).to(dev).train()
HGTConv.keep_att = False
x = torch.randn(g.num_nodes, 256, device=dev, requires_grad=True)
nt, ei, et = g.node_type.to(dev), g.edge_index.to(dev), g.edge_type.to(dev)
w = torch.randn(g.num_nodes, 256, device=dev)
for i in range(a.iters):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = m(x, nt, ei, et)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    (out * w).sum().backward()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    m.zero_grad(); x.grad = None
    print("iter %d: N=%d E=%d fwd %.1f ms  bwd %.1f ms  -> %.1f M edges/s fwd+bwd, peak mem %.1f GB"
          % (i, g.num_nodes, g.num_edges, (t1 - t0) * 1e3, (t2 - t1) * 1e3, g.num_edges / (t2 - t0) / 1e6,
             torch.cuda.max_memory_allocated() / 1e9), flush=True)
if not a.profile:
    sys.exit(0)
if not a.profile:
    sys.exit(0)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    out = m(x, nt, ei, et); (out * w).sum().backward(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=10, max_name_column_width=60))
