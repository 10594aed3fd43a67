"""Host-side logic of the multi-GPU path (pyhgt_b200/sharded.py) on CPU with the gloo backend, world_size 2:
partition + halo all-to-all must hand every rank exactly the rows it needs, so that the ORACLE run on the local
shard reproduces the full-graph oracle on the owned rows.  (The CUDA kernels are exercised by the -m gpu tests.)"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import hgt_oracle
from pyhgt_b200 import sharded, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        g = synth.make_random(600, 6000, 3, 4, seed=5, isolated_frac=0.2, self_loops=50, duplicate_edges=80)
        d, H = 32, 4
        params = hgt_oracle.init_params(d, d, 3, 4, H, use_norm=True, use_RTE=True, seed=1)
        x = torch.randn(g.num_nodes, d, generator=torch.Generator().manual_seed(2))
        kw = dict(num_types=3, num_relations=4, n_heads=H, use_norm=True, use_RTE=True)
        full, _ = hgt_oracle.hgt_forward_ref_port(params, x, g.node_type, g.edge_index, g.edge_type, g.edge_time, **kw)
        sh = sharded.ShardedGraph.build(g.node_type, g.edge_index, g.edge_type, g.edge_time, 3, 4, rank, world,
                                        torch.device("cpu"))
        x_local = sh.exchange(x[sh.owned_global])
        # the exchange delivered exactly the halo rows
        assert torch.equal(x_local, x[sh.local_global])
        loc, _ = hgt_oracle.hgt_forward_ref_port(params, x_local, sh.node_type, sh.edge_index, sh.edge_type,
                                                 sh.edge_time, **kw)
        err = (loc[sh.own_rows] - full[sh.owned_global]).abs().max().item()
        owned = torch.zeros(g.num_nodes, dtype=torch.int64)
        owned[sh.owned_global] = 1
        dist.all_reduce(owned)
        edges = torch.tensor([sh.n_local_edges])
        dist.all_reduce(edges)
        ret[rank] = (err, bool((owned == 1).all()), int(edges.item()) == g.num_edges, sh.n_owned, sh.n_local_edges,
                     sh.active_per_type)
    finally:
        dist.destroy_process_group()


def test_two_rank_shards_reproduce_full_graph():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        err, each_node_once, all_edges, n_owned, n_edges, active = ret[r]
        assert err < 1e-5, "rank %d: sharded oracle differs from full oracle by %g" % (r, err)
        assert each_node_once and all_edges
        assert sum(active) == n_owned
    # cost-balanced: neither rank holds more than 60 % of the edges
    e = [ret[r][4] for r in range(world)]
    assert max(e) <= 0.6 * sum(e)


def test_partition_is_contiguous_inside_each_type():
    g = synth.make_mag_shaped(scale=0.002, seed=3)
    owner = sharded.partition_owner(g.node_type, g.edge_index, 4, 4)
    for t in range(4):
        o = owner[g.node_type == t]
        assert torch.all(o[1:] >= o[:-1])          # non-decreasing = contiguous blocks
    deg = torch.bincount(g.edge_index[1], minlength=g.num_nodes)
    per_rank = torch.zeros(4).index_add_(0, owner, (2 * deg + 1).float())
    assert per_rank.max() <= 1.25 * per_rank.mean()
