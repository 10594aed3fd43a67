"""Time ShardedGraph.build (device-side partition passes) for one rank of a world of W on the C2 graph."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyhgt_b200 import synth, sharded

W = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
g = synth.make_mag_shaped(1.0)
torch.zeros(1, device=dev)
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sh = sharded.ShardedGraph.build(g.node_type, g.edge_index, g.edge_type, None, 4, 4, 0, W, dev)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("world %d rank 0: build %.2f s  (owned %d, halo %d, local edges %d)" % (W, t1 - t0, sh.n_owned, sh.n_halo, sh.n_local_edges), flush=True)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    sharded.ShardedGraph.build(g.node_type, g.edge_index, g.edge_type, None, 4, 4, 0, W, dev)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cpu_time_total", row_limit=14, max_name_column_width=50))
