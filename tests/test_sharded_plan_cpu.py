"""Structural invariants of the per-rank shard plan (pyhgt_b200/sharded.py: ShardedGraph.build) — pure host logic, no
process group: the build is a deterministic function of (graph, rank, world), so all ranks' plans can be built in one
process and checked against each other.  Covers what the round-2 multi-GPU kernels rely on: the peer-memory pull tables,
the rank-staggered processing order, the per-<type, relation> compaction runs and the all_to_all split symmetry."""
import torch

from pyhgt_b200 import sharded, synth

WORLD = 4
T, R = 3, 5


def _shards():
    g = synth.make_random(900, 9000, T, R, seed=11, isolated_frac=0.15, self_loops=40, duplicate_edges=60)
    shs = [sharded.ShardedGraph.build(g.node_type, g.edge_index, g.edge_type, g.edge_time, T, R, r, WORLD,
                                      torch.device("cpu"), halo_mode="nccl") for r in range(WORLD)]
    return g, shs


def test_pull_tables_address_the_owner_rows():
    g, shs = _shards()
    for sh in shs:
        n_local = sh.local_global.numel()
        assert n_local == sh.n_owned + sh.n_halo
        pr, prow = sh.pull_rank.long(), sh.pull_row.long()
        for i in range(n_local):
            assert int(shs[int(pr[i])].owned_global[int(prow[i])]) == int(sh.local_global[i])
        # local node order is type-sorted, owned rows before halo rows inside a type
        nt = sh.node_type
        assert torch.all(nt[1:] >= nt[:-1])
        own = torch.zeros(n_local, dtype=torch.bool)
        own[sh.own_rows] = True
        for t in range(T):
            o = own[nt == t].to(torch.int8)
            assert torch.all(o[1:] <= o[:-1])


def test_pull_order_is_a_rank_staggered_permutation():
    _, shs = _shards()
    for sh in shs:
        po = sh.pull_order.long()
        assert torch.equal(torch.sort(po)[0], torch.arange(po.numel()))
        owners = sh.pull_rank.long()[po]
        present = torch.unique(sh.pull_rank.long()).tolist()
        # the first items cycle through every source rank once, starting behind this rank
        head = owners[:len(present)].tolist()
        assert sorted(head) == sorted(present)
        ring = sorted(present, key=lambda p: (p - sh.rank - 1) % WORLD)
        assert head == ring


def test_kv_runs_cover_exactly_the_rows_a_pair_needs():
    _, shs = _shards()
    for sh in shs:
        assert sh.kv_runs is not None
        nt = sh.node_type
        type_row0 = [int((nt < t).sum()) for t in range(T + 1)]
        runs = dict(sh.kv_runs)
        src, rel = sh.edge_index[0], sh.edge_type
        for t in range(T):
            for r in range(R):
                need = torch.unique(src[(rel == r) & (nt[src] == t)]) - type_row0[t]
                got = [i for a, b in runs.get((t, r), ()) for i in range(a, b)]
                assert sorted(need.tolist()) == got, (sh.rank, t, r)
        # compaction is worth something on this graph: fewer projected <pair, row> items than rows x relations
        total = sum(b - a for _, rr in sh.kv_runs for a, b in rr)
        assert total < sum(int((nt == t).sum()) for t in range(T)) * R


def test_all_to_all_splits_are_symmetric_and_stats_match():
    _, shs = _shards()
    for a in shs:
        assert a.recv_splits[a.rank] == 0 and a.send_splits[a.rank] == 0
        assert sum(a.recv_splits) == a.n_halo
        for b in shs:
            assert a.recv_splits[b.rank] == b.send_splits[a.rank]
        st = a.halo_stats(64)
        assert st["halo_rows_rank"] == a.n_halo and st["halo_bytes_rank"] == a.n_halo * 64 * 4
        # the rows a peer sends are the rows this rank lists as halo, in arrival order (by owner, then id)
        off = 0
        for b in shs:
            k = a.recv_splits[b.rank]
            sent = b.owned_global[b.send_idx.long()[sum(b.send_splits[:a.rank]):sum(b.send_splits[:a.rank + 1])]]
            assert torch.equal(sent, a.halo_global[off:off + k])
            off += k
