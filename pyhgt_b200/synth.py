"""Seeded synthetic heterographs in the tensor layout pyHGT's ``to_torch`` produces
(reference: pyHGT/data.py:226-256): nodes concatenated type by type (``node_type`` piecewise
constant), ``edge_index`` int64 [2,E] with row 0 = source and row 1 = target, edges appended in
<target_type, source_type, relation> blocks (NOT sorted by destination), ``edge_type`` int64 [E],
``edge_time`` int64 [E] in [0,240).

Shapes follow BASELINE.json ``configs`` / SURVEY.md §8(d):
  c1  2 types / 1 relation, 1k nodes, 5k edges (d=64, H=4)
  c2  ogbn-mag-shaped: 4 types / 4 relations, 1,939,743 nodes, 21,111,007 edges (d=256, H=8)
  c3  OAG-CS-shaped sampled subgraph: 6 types / 10 relations, 200k nodes, 5M edges (d=400, H=8)
  c5  power-law heterograph, 4 types / 8 relations, N = E/10 (d=128, H=8)
"""
from dataclasses import dataclass

import torch


@dataclass
class HeteroGraph:
    node_type: torch.Tensor     # [N] int64
    edge_index: torch.Tensor    # [2,E] int64
    edge_type: torch.Tensor     # [E] int64
    edge_time: torch.Tensor     # [E] int64 in [0,240)
    num_types: int
    num_relations: int
    name: str = ""

    @property
    def num_nodes(self):
        return self.node_type.numel()

    @property
    def num_edges(self):
        return self.edge_type.numel()


# (name, src_type, dst_type, count) — ogbn-mag's raw relations (ogbn-mag/preprocess_ogbn_mag.py:33-42
# adds rev_* and self on top; BASELINE.json fixes R=4).
MAG_NODE_COUNTS = (736389, 1134649, 8740, 59965)          # paper, author, institution, field
MAG_RELATIONS = (
    ("writes", 1, 0, 7145660),
    ("cites", 0, 0, 5416271),
    ("has_topic", 0, 3, 7505078),
    ("affiliated_with", 1, 2, 1043998),
)

OAG_TYPE_FRACTIONS = (0.45, 0.35, 0.08, 0.02, 0.05, 0.05)  # paper, author, field, venue, affiliation, extra
# (src_type, dst_type, weight): 9 typed relations + relation 9 = 'self' over every node (data.py:183-186)
OAG_RELATIONS = (
    (1, 0, 0.22), (0, 1, 0.22), (0, 0, 0.16), (0, 2, 0.10), (2, 0, 0.10),
    (0, 3, 0.04), (3, 0, 0.04), (1, 4, 0.04), (4, 1, 0.04),
)


def _type_layout(counts):
    node_type = torch.cat([torch.full((c,), t, dtype=torch.int64) for t, c in enumerate(counts)])
    starts = [0]
    for c in counts:
        starts.append(starts[-1] + c)
    return node_type, starts


def _zipf_ids(gen, n_ids, count, alpha):
    """Power-law ids in [0, n_ids): inverse-CDF of p(k) ~ (k+1)^-alpha, then a fixed shuffle so
    hubs are spread over the id range."""
    u = torch.rand(count, generator=gen, dtype=torch.float64)
    if abs(alpha - 1.0) < 1e-9:
        k = torch.exp(u * torch.log(torch.tensor(float(n_ids) + 1.0))) - 1.0
    else:
        a = 1.0 - alpha
        hi = (float(n_ids) + 1.0) ** a
        k = (u * (hi - 1.0) + 1.0) ** (1.0 / a) - 1.0
    k = k.clamp_(0, n_ids - 1).to(torch.int64)
    mix = (k * 2654435761) % n_ids
    return mix


def _assemble(blocks, counts, T, R, gen, name):
    node_type, _ = _type_layout(counts)
    src = torch.cat([b[0] for b in blocks]) if blocks else torch.zeros(0, dtype=torch.int64)
    dst = torch.cat([b[1] for b in blocks]) if blocks else torch.zeros(0, dtype=torch.int64)
    rel = torch.cat([torch.full((b[0].numel(),), b[2], dtype=torch.int64) for b in blocks]) \
        if blocks else torch.zeros(0, dtype=torch.int64)
    e = src.numel()
    etime = torch.randint(0, 240, (e,), generator=gen, dtype=torch.int64)
    return HeteroGraph(node_type, torch.stack([src, dst]), rel, etime, T, R, name)


def make_c1(seed=1, n_nodes=1000, n_edges=5000):
    """BASELINE config 1: random 2-type / 1-relation graph; node types NOT sorted."""
    g = torch.Generator().manual_seed(seed)
    node_type = torch.randint(0, 2, (n_nodes,), generator=g, dtype=torch.int64)
    src = torch.randint(0, n_nodes, (n_edges,), generator=g, dtype=torch.int64)
    dst = torch.randint(0, n_nodes, (n_edges,), generator=g, dtype=torch.int64)
    etime = torch.randint(0, 240, (n_edges,), generator=g, dtype=torch.int64)
    return HeteroGraph(node_type, torch.stack([src, dst]), torch.zeros(n_edges, dtype=torch.int64),
                       etime, 2, 1, "c1")


def make_mag_shaped(scale=1.0, seed=2, dst_zipf=None):
    """BASELINE config 2 (scale=1.0).  Endpoints uniform inside each type's id range;
    ``dst_zipf=alpha`` gives the skewed-destination variant."""
    g = torch.Generator().manual_seed(seed)
    counts = [max(2, int(round(c * scale))) for c in MAG_NODE_COUNTS]
    starts = [0]
    for c in counts:
        starts.append(starts[-1] + c)
    blocks = []
    for r, (_, s, t, cnt) in enumerate(MAG_RELATIONS):
        m = max(1, int(round(cnt * scale)))
        src = torch.randint(0, counts[s], (m,), generator=g, dtype=torch.int64) + starts[s]
        if dst_zipf is None:
            dst = torch.randint(0, counts[t], (m,), generator=g, dtype=torch.int64)
        else:
            dst = _zipf_ids(g, counts[t], m, dst_zipf)
        blocks.append((src, dst + starts[t], r))
    name = "c2-ogbn-mag-shaped" + ("" if scale == 1.0 else "-x%g" % scale)
    return _assemble(blocks, counts, 4, 4, g, name)


def make_oag_shaped(scale=1.0, seed=3, n_nodes=200_000, n_edges=5_000_000):
    """BASELINE config 3: 6 types / 10 relations incl. a 'self' relation on every node."""
    g = torch.Generator().manual_seed(seed)
    n = max(12, int(round(n_nodes * scale)))
    e = max(n + 9, int(round(n_edges * scale)))
    counts = [max(2, int(round(f * n))) for f in OAG_TYPE_FRACTIONS]
    starts = [0]
    for c in counts:
        starts.append(starts[-1] + c)
    n = starts[-1]
    blocks = []
    typed_edges = e - n
    for r, (s, t, w) in enumerate(OAG_RELATIONS):
        m = max(1, int(round(typed_edges * w / sum(x[2] for x in OAG_RELATIONS))))
        src = torch.randint(0, counts[s], (m,), generator=g, dtype=torch.int64) + starts[s]
        dst = torch.randint(0, counts[t], (m,), generator=g, dtype=torch.int64) + starts[t]
        blocks.append((src, dst, r))
    ids = torch.arange(n, dtype=torch.int64)
    blocks.append((ids, ids.clone(), 9))
    gr = _assemble(blocks, counts, 6, 10, g, "c3-oag-cs-shaped" + ("" if scale == 1.0 else "-x%g" % scale))
    # edge_time = clip(dyear + 120, 0, 239), years ~ N(0, 8); self loops get exactly 120
    et = (torch.randn(gr.num_edges, generator=g) * 8).round().to(torch.int64) + 120
    et[-n:] = 120
    gr.edge_time = et.clamp_(0, 239)
    return gr


def make_powerlaw(n_edges, seed=5, alpha=2.0, n_types=4, n_relations=8):
    """BASELINE config 5: power-law destinations (Zipf alpha), N = E/10, 4 types / 8 relations;
    relation r goes from type (r % T) to type ((r // 2) % T)."""
    g = torch.Generator().manual_seed(seed)
    n = max(n_types * 2, n_edges // 10)
    counts = [n // n_types] * n_types
    counts[-1] += n - sum(counts)
    starts = [0]
    for c in counts:
        starts.append(starts[-1] + c)
    blocks = []
    per = n_edges // n_relations
    for r in range(n_relations):
        m = per if r < n_relations - 1 else n_edges - per * (n_relations - 1)
        s, t = r % n_types, (r // 2) % n_types
        src = torch.randint(0, counts[s], (m,), generator=g, dtype=torch.int64) + starts[s]
        dst = _zipf_ids(g, counts[t], m, alpha) + starts[t]
        blocks.append((src, dst, r))
    return _assemble(blocks, counts, n_types, n_relations, g, "c5-powerlaw-%dM" % (n_edges // 1_000_000))


def make_random(n_nodes, n_edges, num_types, num_relations, seed=0, sorted_types=False,
                isolated_frac=0.0, self_loops=0, duplicate_edges=0):
    """Unstructured test graph: any relation between any pair of types, optional isolated
    destinations, self loops and duplicated (multi-)edges."""
    g = torch.Generator().manual_seed(seed)
    node_type = torch.randint(0, num_types, (n_nodes,), generator=g, dtype=torch.int64)
    if sorted_types:
        node_type = node_type.sort().values
    n_dst = max(1, int(n_nodes * (1.0 - isolated_frac)))
    src = torch.randint(0, n_nodes, (n_edges,), generator=g, dtype=torch.int64)
    dst = torch.randint(0, n_dst, (n_edges,), generator=g, dtype=torch.int64)
    if self_loops:
        ids = torch.randint(0, n_nodes, (self_loops,), generator=g, dtype=torch.int64)
        src = torch.cat([src, ids]); dst = torch.cat([dst, ids])
    if duplicate_edges and n_edges:
        pick = torch.randint(0, n_edges, (duplicate_edges,), generator=g, dtype=torch.int64)
        src = torch.cat([src, src[pick]]); dst = torch.cat([dst, dst[pick]])
    e = src.numel()
    rel = torch.randint(0, num_relations, (e,), generator=g, dtype=torch.int64)
    if duplicate_edges and n_edges:
        rel[-duplicate_edges:] = rel[pick]
    etime = torch.randint(0, 240, (e,), generator=g, dtype=torch.int64)
    return HeteroGraph(node_type, torch.stack([src, dst]), rel, etime, num_types, num_relations, "random")
