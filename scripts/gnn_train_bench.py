"""BASELINE config 4 on one GPU (scaled): 3-layer HGT stack, forward + backward, dropout 0."""
import argparse, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyhgt_b200 import HGTConv, synth
from pyhgt_b200.model import GNN

ap = argparse.ArgumentParser(); ap.add_argument("--scale", type=float, default=0.25); a = ap.parse_args()
dev = torch.device("cuda:0")
g = synth.make_mag_shaped(a.scale)
torch.manual_seed(0)
HGTConv.keep_att = False
m = GNN(256, 256, 4, 4, 8, 3, 0.0, "hgt", True, True, False).to(dev).train()
x = torch.randn(g.num_nodes, 256, device=dev)
nt, ei, et, tm = (t.to(dev) for t in (g.node_type, g.edge_index, g.edge_type, g.edge_time))
w = torch.randn(g.num_nodes, 256, device=dev)
for i in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = m(x, nt, tm, ei, et)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    (out * w).sum().backward()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    m.zero_grad()
    print("iter %d: N=%d E=%d 3 layers: fwd %.1f ms  bwd %.1f ms  -> %.1f M edge-layers/s fwd+bwd, peak mem %.1f GB"
          % (i, g.num_nodes, g.num_edges, (t1 - t0) * 1e3, (t2 - t1) * 1e3, 3 * g.num_edges / (t2 - t0) / 1e6,
             torch.cuda.max_memory_allocated() / 1e9), flush=True)
