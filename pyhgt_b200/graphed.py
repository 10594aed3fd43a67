"""CUDA-graph replay for the sampled-subgraph regime (pyHGT's own training / inference loop feeds a NEW small graph every
batch: OAG/train_paper_field.py:241, ~10^3-10^5 nodes).  At that size a layer is launch-bound, so the whole sequence

    per-graph plan build (CSR by destination, gather rows, tiles)  ->  HGT layers

is captured ONCE into a CUDA graph for a static *signature* — padded node count per type, padded edge count, the set of
<source type, relation> pairs — and replayed per batch; the host then issues a few copies and one graph launch.  This works
because the plan build is sync-free when the host knows the signature (plan.get_plan(host_meta=...)): no read-back, tile
counts stay on the device, grids are sized by upper bounds.

Batches are padded to the signature on the host (vectorised numpy):
  * nodes stay type-contiguous (the layout `to_torch` produces, data.py:232-235); type t gets `type_counts[t]` slots, real
    nodes first, the rest isolated zero-feature nodes whose output rows are dropped;
  * one extra node of out-of-range type closes the array; padding edges are self loops on it (they match no
    <source type, target type, relation> triple and their destination row is discarded), so they cannot touch a real row.
"""
import numpy as np
import torch

from . import _lib
from . import plan as _plan


class GraphSignature:
    """Static shape of a family of batches."""

    def __init__(self, type_counts, n_edges, pairs, num_relations, feat_dim, use_time=True):
        self.type_counts = [int(c) for c in type_counts]
        self.n_edges = int(n_edges)
        self.pairs = sorted({(int(s), int(r)) for s, r in pairs})
        self.num_types = len(self.type_counts)
        self.num_relations = int(num_relations)
        self.feat_dim = int(feat_dim)
        self.use_time = bool(use_time)
        self.n_nodes = sum(self.type_counts) + 1            # + the trailing out-of-range-type node
        self.row0 = np.concatenate([[0], np.cumsum(self.type_counts)]).astype(np.int64)
        self.pair_mask = np.zeros(self.num_types * self.num_relations, dtype=bool)
        for s_, r_ in self.pairs:
            if 0 <= s_ < self.num_types and 0 <= r_ < self.num_relations:
                self.pair_mask[s_ * self.num_relations + r_] = True
        self.node_type = np.repeat(np.arange(self.num_types + 1, dtype=np.int64), self.type_counts + [1])   # static

    def fits(self, counts, n_edges, pairs):
        return (len(counts) == self.num_types and all(c <= C for c, C in zip(counts, self.type_counts))
                and n_edges <= self.n_edges and set(pairs) <= set(self.pairs))

    def host_meta(self):
        return {"type_count": self.type_counts + [1], "sorted": True, "pairs": self.pairs}


def pad_batch(sig, node_feature, node_type, edge_time, edge_index, edge_type, out=None):
    """Host tensors of one batch (type-contiguous node order) -> padded numpy arrays of the signature's shape and the
    new index of every real node.  `out` = (x, edge_time, edge_index, edge_type) numpy views to fill in place (pinned
    staging).  Raises if the batch does not fit."""
    nt = node_type.numpy()
    if nt.size and np.any(nt[1:] < nt[:-1]):
        raise ValueError("pad_batch needs type-contiguous nodes (what to_torch emits)")
    T = sig.num_types
    if nt.size and (nt[0] < 0 or nt[-1] >= T):
        raise ValueError("pad_batch: node types must lie in [0, %d)" % T)
    counts = np.bincount(nt, minlength=T)[:T] if nt.size else np.zeros(T, dtype=np.int64)
    src, dst = edge_index[0].numpy(), edge_index[1].numpy()
    et = edge_type.numpy()
    E = et.size
    ok = len(counts) == T and bool(np.all(counts <= np.asarray(sig.type_counts))) and E <= sig.n_edges
    if ok and E:
        if et.min() < 0 or et.max() >= sig.num_relations:
            ok = False
        else:
            present = np.bincount(nt[src] * sig.num_relations + et, minlength=T * sig.num_relations) > 0
            ok = not bool(np.any(present & ~sig.pair_mask))
    if not ok:
        raise ValueError("batch (type counts %s, %d edges) does not fit the signature (%s, %d edges) or has new "
                         "<type, relation> pairs" % (counts.tolist(), E, sig.type_counts, sig.n_edges))
    old0 = np.concatenate([[0], np.cumsum(counts)])
    shift = sig.row0[:T] - old0[:T]
    new_id = np.arange(nt.size, dtype=np.int64) + shift[nt] if nt.size else np.zeros(0, dtype=np.int64)
    if out is None:
        out = (np.empty((sig.n_nodes, sig.feat_dim), dtype=np.float32), np.empty(sig.n_edges, dtype=np.int64),
               np.empty((2, sig.n_edges), dtype=np.int64), np.empty(sig.n_edges, dtype=np.int64))
    x, etm, ei, ety = out
    x.fill(0.0)
    x[new_id] = node_feature.numpy()
    pad_node = sig.n_nodes - 1
    ei[0, :E] = new_id[src]
    ei[1, :E] = new_id[dst]
    ei[:, E:] = pad_node
    ety[:E] = et
    ety[E:] = 0
    etm[E:] = 120
    if edge_time is not None:
        etm[:E] = edge_time.numpy()
    else:
        etm[:E] = 120
    return x, sig.node_type, etm, ei, ety, new_id


class GraphedForward:
    """Capture `fn(node_feature, node_type, edge_time, edge_index, edge_type) -> [N, d]` (an HGTConv / GNN forward under
    no_grad; note GNN's argument order, model.py:69) for one signature and replay it per batch.

        sig = GraphSignature(type_counts=[...], n_edges=..., pairs=[...], num_relations=R, feat_dim=F)
        g = GraphedForward(lambda x, nt, tm, ei, et: gnn(x, nt, tm, ei, et), sig, device)
        out = g(node_feature, node_type, edge_time, edge_index, edge_type)     # host tensors of one batch
    """

    def __init__(self, fn, sig, device):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.HgtError("GraphedForward needs a CUDA device")
        self.fn, self.sig, self.dev = fn, sig, dev
        i64 = dict(dtype=torch.int64, device=dev)
        self.x = torch.zeros((sig.n_nodes, sig.feat_dim), dtype=torch.float32, device=dev)
        self.nt = torch.zeros(sig.n_nodes, **i64)
        self.ei = torch.zeros((2, sig.n_edges), **i64)
        self.et = torch.zeros(sig.n_edges, **i64)
        self.tm = torch.zeros(sig.n_edges, **i64)
        # pinned staging for the per-batch copies
        self.h = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in (self.x, self.nt, self.tm, self.ei, self.et)]
        self.graph = None
        self.out = None
        self.plan = None
        self._staged = None
        self._nt_done = False
        self._pins = []
        self.stream = torch.cuda.Stream(device=dev)

    def _run(self):
        s = self.sig
        self.plan = _plan.rebuild_plan(self.nt, self.ei, self.et, self.tm if s.use_time else None, s.num_types,
                                       s.num_relations, s.host_meta())
        with torch.no_grad():
            return self.fn(self.x, self.nt, self.tm, self.ei, self.et)

    def _stage(self, batch):
        """Pad the batch straight into the pinned staging buffers and enqueue the copies (node_type is static)."""
        if self._staged is not None:
            self._staged.synchronize()                       # the previous batch's copies have left the staging buffers
        hx, hnt, htm, hei, het = self.h
        out = (hx.numpy(), htm.numpy(), hei.numpy(), het.numpy())
        new_id = pad_batch(self.sig, *batch, out=out)[5]
        if not self._nt_done:
            hnt.copy_(torch.from_numpy(self.sig.node_type))
            self.nt.copy_(hnt, non_blocking=True)
            self._nt_done = True
        for h, dst in ((hx, self.x), (htm, self.tm), (hei, self.ei), (het, self.et)):
            dst.copy_(h, non_blocking=True)
        self._staged = torch.cuda.Event()
        self._staged.record(self.stream)
        return new_id

    def __call__(self, node_feature, node_type, edge_time, edge_index, edge_type):
        cur = torch.cuda.current_stream(self.dev)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            new_id = self._stage((node_feature, node_type, edge_time, edge_index, edge_type))
            if self.graph is None:
                for _ in range(2):                          # eager warm-up: pointer tables, pinned-block cache
                    self._run()
                self.stream.synchronize()
                # the table uploads captured below re-read their pinned sources at every replay: keep them (self._pins)
                _plan._PIN_KEEP = self._pins
                try:
                    self.graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self.graph, stream=self.stream):
                        self.out = self._run()
                finally:
                    _plan._PIN_KEEP = None
            self.graph.replay()
            idx = torch.from_numpy(new_id).pin_memory().to(self.dev, non_blocking=True)
            res = self.out.index_select(0, idx)
        cur.wait_stream(self.stream)
        res.record_stream(cur)
        return res
