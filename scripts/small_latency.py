"""Latency of one HGTConv forward on a sampled-subgraph-sized input (BASELINE config 1: 1k nodes / 5k edges),
with and without the per-graph plan build — the regime of pyHGT's own training loop (new graph every batch)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyhgt_b200 import HGTConv, synth, clear_plan_cache

dev = torch.device("cuda:0")
for name, g, d, H in (("c1", synth.make_c1(), 64, 4), ("c1-sorted", synth.make_random(1000, 5000, 2, 1, seed=1, sorted_types=True), 64, 4),
                      ("oag-batch", synth.make_oag_shaped(0.05), 400, 8)):
    torch.manual_seed(0)
    m = HGTConv(d, d, g.num_types, g.num_relations, H, 0.2, True, True).to(dev).eval()
    x = torch.randn(g.num_nodes, d, device=dev)
    nt, ei, et, tm = (t.to(dev) for t in (g.node_type, g.edge_index, g.edge_type, g.edge_time))
    with torch.no_grad():
        for _ in range(5):
            m(x, nt, ei, et, tm)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            m(x, nt, ei, et, tm)
        torch.cuda.synchronize()
        warm = (time.perf_counter() - t0) / 50 * 1e3
        t0 = time.perf_counter()
        for _ in range(20):
            clear_plan_cache()
            m(x, nt, ei, et, tm)
        torch.cuda.synchronize()
        cold = (time.perf_counter() - t0) / 20 * 1e3
        # sync-free plan: the host knows the per-type counts and the <type, relation> pairs (what data.to_torch passes)
        from pyhgt_b200 import plan as P
        sorted_ok = bool((g.node_type[1:] >= g.node_type[:-1]).all())
        meta = {"type_count": torch.bincount(g.node_type, minlength=g.num_types).tolist() + [0], "sorted": sorted_ok,
                "pairs": sorted({(int(a), int(b)) for a, b in zip(g.node_type[g.edge_index[0]].tolist(), g.edge_type.tolist())})}
        for _ in range(3):
            clear_plan_cache(); P.get_plan(nt, ei, et, tm, g.num_types, g.num_relations, host_meta=meta); m(x, nt, ei, et, tm)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            clear_plan_cache()
            P.get_plan(nt, ei, et, tm, g.num_types, g.num_relations, host_meta=meta)
            m(x, nt, ei, et, tm)
        torch.cuda.synchronize()
        free = (time.perf_counter() - t0) / 50 * 1e3
    print("%s: sync-free plan rebuilt every call: %.3f ms" % (name, free))
    if sorted_ok:
        # CUDA-graph replay of plan build + layer for a padded signature, new host batch every call
        from pyhgt_b200 import graphed
        sig = graphed.GraphSignature(meta["type_count"][:-1], g.num_edges, meta["pairs"], g.num_relations, d)
        gf = graphed.GraphedForward(lambda x_, nt_, tm_, ei_, et_: m(x_, nt_, ei_, et_, tm_), sig, dev)
        xh = x.cpu()
        for _ in range(3):
            gf(xh, g.node_type, g.edge_time, g.edge_index, g.edge_type)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            gf(xh, g.node_type, g.edge_time, g.edge_index, g.edge_type)
        torch.cuda.synchronize()
        per_call = (time.perf_counter() - t0) / 50 * 1e3
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(gf.stream):
            ev0.record()
            for _ in range(20):
                gf.graph.replay()
            ev1.record()
        torch.cuda.synchronize()
        print("%s: CUDA-graph replay incl. host padding + H2D of the batch: %.3f ms per call; device time of one replay "
              "(plan build + layer): %.3f ms" % (name, per_call, ev0.elapsed_time(ev1) / 20))
    print("%s: N=%d E=%d d=%d  forward %.3f ms (plan cached)  %.3f ms (plan rebuilt)  -> %.1f / %.1f M edges/s"
          % (name, g.num_nodes, g.num_edges, d, warm, cold, g.num_edges / warm / 1e3, g.num_edges / cold / 1e3))
