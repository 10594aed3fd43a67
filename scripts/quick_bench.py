"""Development helper: time one HGTConv forward on a synthetic config and print the per-kernel breakdown
(CUPTI via torch.profiler).  Not the contract bench — see bench.py."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyhgt_b200 import HGTConv, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--linear", type=int, default=0)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--att", type=int, default=0)
    ap.add_argument("--profile", type=int, default=1)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    t0 = time.time()
    if a.config == "c2":
        g, d, H, rte = synth.make_mag_shaped(a.scale), 256, 8, False
    elif a.config == "c3":
        g, d, H, rte = synth.make_oag_shaped(a.scale), 400, 8, True
    elif a.config == "c5":
        g, d, H, rte = synth.make_powerlaw(int(16_000_000 * a.scale)), 128, 8, False
    else:
        raise SystemExit("unknown config")
    print("graph %s: N=%d E=%d built in %.1fs" % (g.name, g.num_nodes, g.num_edges, time.time() - t0), flush=True)
    torch.manual_seed(0)
    m = HGTConv(d, d, g.num_types, g.num_relations, H, 0.2, True, rte).to(dev).eval()
    m.edge_variant, m.linear_impl = a.variant, a.linear
    HGTConv.keep_att = bool(a.att)
    x = torch.randn(g.num_nodes, d, device=dev)
    nt, ei, et, tm = (t.to(dev) for t in (g.node_type, g.edge_index, g.edge_type, g.edge_time))
    with torch.no_grad():
        torch.cuda.synchronize(); t0 = time.time()
        out = m(x, nt, ei, et, tm if rte else None)
        torch.cuda.synchronize()
        print("first forward (incl. plan build) %.1f ms" % ((time.time() - t0) * 1e3), flush=True)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
        for i in range(a.iters):
            ev[i].record()
            out = m(x, nt, ei, et, tm if rte else None)
        ev[a.iters].record()
        torch.cuda.synchronize()
        ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.iters)]
        best = min(ts)
        print("forward ms: %s  -> best %.2f ms = %.1f M edges/s" % (["%.2f" % t for t in ts], best,
                                                                    g.num_edges / best / 1e3), flush=True)
        print("out finite:", bool(torch.isfinite(out).all()), "mean|out|=%.4f" % out.abs().mean().item())
        if a.profile:
            from torch.profiler import profile, ProfilerActivity
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                for _ in range(2):
                    m(x, nt, ei, et, tm if rte else None)
                torch.cuda.synchronize()
            print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))


if __name__ == "__main__":
    main()
