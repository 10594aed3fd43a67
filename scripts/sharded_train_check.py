"""2-GPU check of the sharded forward+backward (BASELINE config 4): gradients of the dst-sharded layer (reverse halo
all-to-all + all-reduced parameter gradients) against the single-GPU full-graph autograd result, plus timing.
    python -m torch.distributed.run --nproc-per-node 2 scripts/sharded_train_check.py"""
import os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyhgt_b200 import HGTConv, synth, sharded

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.02
g = synth.make_mag_shaped(scale)
HGTConv.keep_att = False
torch.manual_seed(0)
ref = HGTConv(256, 256, 4, 4, 8, 0.0, True, True).to(dev).train()
m = HGTConv(256, 256, 4, 4, 8, 0.0, True, True).to(dev).train()
m.load_state_dict(ref.state_dict())
gen = torch.Generator().manual_seed(1)
x = torch.randn(g.num_nodes, 256, generator=gen)
w = torch.randn(g.num_nodes, 256, generator=gen)
xr = x.to(dev).requires_grad_(True)
out_ref = ref(xr, g.node_type.to(dev), g.edge_index.to(dev), g.edge_type.to(dev), g.edge_time.to(dev))
(out_ref * w.to(dev)).sum().backward()
sh = sharded.ShardedGraph.build(g.node_type, g.edge_index, g.edge_type, g.edge_time, 4, 4, rank, world, dev,
                                halo_mode="nccl")
x_own = x[sh.owned_global].to(dev).requires_grad_(True)
w_own = w[sh.owned_global].to(dev)
for it in range(3):
    m.zero_grad(); x_own.grad = None
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    out = sh.forward_train(m, x_own)
    (out * w_own).sum().backward()
    sh.allreduce_grads(m)
    torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
own = sh.owned_global.to(dev)
e_out = (out - out_ref[own]).abs().max().item()
e_x = (x_own.grad - xr.grad[own]).abs().max().item() / max(xr.grad.abs().max().item(), 1e-12)
e_p = max(((a.grad - b.grad).abs().max() / b.grad.abs().max().clamp_min(1e-12)).item()
          for a, b in zip(m.parameters(), ref.parameters()) if b.grad is not None)
print("rank %d/%d: N=%d E=%d  fwd+bwd %.1f ms  max|out-ref| %.2e  rel err d x_own %.2e  rel err d params %.2e"
      % (rank, world, g.num_nodes, g.num_edges, dt * 1e3, e_out, e_x, e_p), flush=True)
assert e_out < 1e-3 and e_x < 2e-3 and e_p < 2e-3
dist.destroy_process_group()
