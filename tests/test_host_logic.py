"""Host-side logic that needs no GPU: batch padding for CUDA-graph replay, the pull plan of the sharded exchange, the
per-pair compaction runs, the disjoint backward sub-tables."""
import numpy as np
import pytest
import torch

from pyhgt_b200 import graphed, sharded, synth


def test_pad_batch_preserves_the_graph_and_pads_harmlessly():
    b = synth.make_random(400, 3000, 3, 4, seed=1, sorted_types=True)
    pairs = {(int(b.node_type[s]), int(r)) for s, r in zip(b.edge_index[0].tolist(), b.edge_type.tolist())}
    counts = torch.bincount(b.node_type, minlength=3).tolist()
    sig = graphed.GraphSignature([c + 7 for c in counts], 3600, pairs, 4, 16)
    x = torch.randn(400, 16)
    px, pnt, ptm, pei, pet, new_id = graphed.pad_batch(sig, x, b.node_type, b.edge_time, b.edge_index, b.edge_type)
    assert px.shape == (sig.n_nodes, 16) and pei.shape == (2, 3600)
    assert np.array_equal(px[new_id], x.numpy())                              # features moved with their nodes
    assert np.array_equal(pnt[new_id], b.node_type.numpy())                   # types preserved, layout type-contiguous
    assert np.all(np.diff(pnt) >= 0) and pnt[-1] == 3                         # trailing node of out-of-range type
    E = b.edge_type.numel()
    assert np.array_equal(pei[0, :E], new_id[b.edge_index[0].numpy()]) and np.array_equal(pei[1, :E], new_id[b.edge_index[1].numpy()])
    assert np.all(pei[:, E:] == sig.n_nodes - 1)                              # padding edges: self loops on that node
    assert np.array_equal(pet[:E], b.edge_type.numpy()) and np.array_equal(ptm[:E], b.edge_time.numpy())
    # a batch with more nodes of a type than the signature allows, or a new <type, relation> pair, is refused
    with pytest.raises(ValueError):
        graphed.pad_batch(graphed.GraphSignature(counts, 100, pairs, 4, 16), x, b.node_type, b.edge_time, b.edge_index,
                          b.edge_type)
    with pytest.raises(ValueError):
        graphed.pad_batch(graphed.GraphSignature([c + 7 for c in counts], 3600, list(pairs)[:3], 4, 16), x, b.node_type,
                          b.edge_time, b.edge_index, b.edge_type)


@pytest.mark.parametrize("world", [2, 4])
def test_pull_plan_is_a_staggered_permutation_and_compaction_runs_cover_every_edge(world):
    g = synth.make_random(900, 9000, 3, 2, seed=7, isolated_frac=0.1, self_loops=40)
    for rank in range(world):
        sh = sharded.ShardedGraph.build(g.node_type, g.edge_index, g.edge_type, None, 3, 2, rank, world, torch.device("cpu"))
        n_local = sh.n_owned + sh.n_halo
        order = sh.pull_order.long()
        assert sorted(order.tolist()) == list(range(n_local))                 # every local row pulled exactly once
        owners = sh.pull_rank.long()[order]
        first = owners[:world].tolist()
        present = sorted(set(owners.tolist()))
        # consecutive work items cycle through the owners, starting right after this rank
        assert first[0] == next(p for p in [(rank + 1 + i) % world for i in range(world)] if p in present)
        assert len(set(first[:len(present)])) == len(present)
        # the owner's row index is what the owner itself calls that node
        assert torch.equal(sh.pull_row.long()[sh.own_rows], torch.arange(sh.n_owned))
        # every local edge's source row lies inside a K'/V' run of its <source type, relation> pair
        runs = dict(sh.kv_runs)
        src = sh.edge_index[0]
        t_src = sh.node_type[src]
        type_row0 = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(torch.bincount(sh.node_type, minlength=3), 0)])
        rel = src - type_row0[t_src]
        for (t, r), rr in runs.items():
            sel = (t_src == t) & (sh.edge_type == r)
            rows = rel[sel]
            inside = torch.zeros_like(rows, dtype=torch.bool)
            for (a, b) in rr:
                inside |= (rows >= a) & (rows < b)
            assert bool(inside.all()), (rank, t, r)
        for t, r in set(zip(t_src.tolist(), sh.edge_type.tolist())):
            assert (t, r) in runs
