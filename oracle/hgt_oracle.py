"""CPU restatement ("port") of pyHGT's HGTConv forward.  TEST INFRASTRUCTURE ONLY.

Two independent restatements of the same math:

* ``hgt_forward_ref_port`` — follows the reference's own algorithm step by step (per-edge gathers,
  one masked pass per <source_type, target_type, relation> triple with per-EDGE Q/K/V projections,
  PyG segment softmax, index_add aggregation, per-type update).  fp32 torch on CPU.  This is the
  function timed as ``cpu_baseline`` (kind "port") and by ``bench.py --impl reference``: it does the
  same amount of work, in the same order, as /root/reference/pyHGT/conv.py.
* ``hgt_forward_dense_fp64`` — loop-free node-level formulation in float64 that materialises a
  dense [N_dst, N_src-edge] view per destination; an independent cross-check that guards against
  a mistake shared by the port and the PyG shim.  Small graphs only.

Pinning: the reference ships NO tests/golden vectors for this path, so the port
is pinned against OUTPUTS OF THE REFERENCE ITSELF: ``oracle/make_golden.py`` runs the unmodified
/root/reference/pyHGT/conv.py (behind oracle/pyg_shim.py for the three PyG primitives) and commits
the vectors under tests/golden/; tests/test_oracle.py checks both restatements against them.

Parameters are passed as a plain ``state_dict``-style mapping with the reference's names
(conv.py:28-54): relation_pri, relation_att, relation_msg, skip, {k,q,v,a}_linears.{t}.{weight,bias},
norms.{t}.{weight,bias}, emb.emb.weight, emb.lin.{weight,bias}.
"""
import math

import torch
import torch.nn.functional as F


def _segment_softmax(scores, dst, num_nodes):
    """torch_geometric.utils.softmax as called at conv.py:108 (restated from PyG 1.3.2):
    exp(s - segment_max) / (segment_sum + 1e-16), segments keyed by destination id."""
    tail = scores.shape[1:]
    idx = dst.view(-1, *([1] * len(tail))).expand_as(scores)
    seg_max = torch.full((num_nodes, *tail), float("-inf"), dtype=scores.dtype)
    seg_max = seg_max.scatter_reduce(0, idx, scores, reduce="amax", include_self=True)
    ex = torch.exp(scores - seg_max.index_select(0, dst))
    seg_sum = torch.zeros((num_nodes, *tail), dtype=scores.dtype).index_add_(0, dst, ex)
    return ex / (seg_sum.index_select(0, dst) + 1e-16)


def _linear(params, name, t, x):
    return F.linear(x, params["%s.%d.weight" % (name, t)], params["%s.%d.bias" % (name, t)])


def _rte(params, x, dt):
    """RelTemporalEncoding.forward, conv.py:298-299: x + lin(emb(t))."""
    e = F.embedding(dt, params["emb.emb.weight"])
    return x + F.linear(e, params["emb.lin.weight"], params["emb.lin.bias"])


def hgt_forward_ref_port(params, node_inp, node_type, edge_index, edge_type, edge_time=None, *,
                         num_types, num_relations, n_heads, use_norm=True, use_RTE=True):
    """HGTConv.forward (conv.py:56-58) = propagate -> message (conv.py:60-111) -> add-aggregate
    -> update (conv.py:114-134).  Returns (out [N,d], att [E,H])."""
    n_nodes, _ = node_inp.shape
    out_dim = params["q_linears.0.weight"].shape[0]
    d_k = out_dim // n_heads                                   # conv.py:21
    src, dst = edge_index[0], edge_index[1]                     # data.py:245,254: row0 = source
    n_edges = src.numel()
    # PyG propagate: gather endpoint features and types per edge
    x_dst = node_inp.index_select(0, dst)
    x_src = node_inp.index_select(0, src)
    t_dst = node_type.index_select(0, dst)
    t_src = node_type.index_select(0, src)
    scores = torch.zeros(n_edges, n_heads, dtype=node_inp.dtype)              # conv.py:68
    msgs = torch.zeros(n_edges, n_heads, d_k, dtype=node_inp.dtype)           # conv.py:69
    for s in range(num_types):                                                # conv.py:71
        from_s = t_src == s
        for t in range(num_types):                                            # conv.py:75
            st = (t_dst == t) & from_s
            for r in range(num_relations):                                    # conv.py:78
                sel = (edge_type == r) & st
                if int(sel.sum()) == 0:                                       # conv.py:83
                    continue
                xi = x_dst[sel]
                xj = x_src[sel]
                if use_RTE:                                                   # conv.py:91-92
                    xj = _rte(params, xj, edge_time[sel])
                q = _linear(params, "q_linears", t, xi).view(-1, n_heads, d_k)          # conv.py:96
                k = _linear(params, "k_linears", s, xj).view(-1, n_heads, d_k)          # conv.py:97
                k = torch.bmm(k.transpose(1, 0), params["relation_att"][r]).transpose(1, 0)   # :98
                scores[sel] = (q * k).sum(-1) * params["relation_pri"][r] / math.sqrt(d_k)     # :99
                v = _linear(params, "v_linears", s, xj).view(-1, n_heads, d_k)          # conv.py:103
                msgs[sel] = torch.bmm(v.transpose(1, 0), params["relation_msg"][r]).transpose(1, 0)  # :104
    att = _segment_softmax(scores, dst, int(dst.max()) + 1 if n_edges else 0)  # conv.py:108
    weighted = (msgs * att.view(-1, n_heads, 1)).view(-1, out_dim)             # conv.py:109-111
    agg = torch.zeros(n_nodes, out_dim, dtype=node_inp.dtype).index_add_(0, dst, weighted)
    # update, conv.py:114-134 (eval mode: dropout is identity)
    g = F.gelu(agg)                                                            # conv.py:119
    out = torch.zeros(n_nodes, out_dim, dtype=node_inp.dtype)                  # conv.py:120
    for t in range(num_types):
        sel = node_type == t
        if int(sel.sum()) == 0:
            continue
        o = _linear(params, "a_linears", t, g[sel])                            # conv.py:125
        alpha = torch.sigmoid(params["skip"][t])                               # conv.py:129
        y = o * alpha + node_inp[sel] * (1 - alpha)                            # conv.py:131,133
        if use_norm:
            y = F.layer_norm(y, (out_dim,), params["norms.%d.weight" % t], params["norms.%d.bias" % t], 1e-5)
        out[sel] = y
    return out, att


def hgt_forward_dense_fp64(params, node_inp, node_type, edge_index, edge_type, edge_time=None, *,
                           num_types, num_relations, n_heads, use_norm=True, use_RTE=True):
    """Independent float64 restatement: node-level projections (each node projected once with its
    type's weights; the reference's per-edge projection of the same row is algebraically identical),
    per-edge relation transform, explicit per-destination softmax via sorting.  Returns (out, att)."""
    p = {k: v.double() for k, v in params.items()}
    x = node_inp.double()
    n_nodes = x.shape[0]
    d = p["q_linears.0.weight"].shape[0]
    d_k = d // n_heads
    src, dst = edge_index[0], edge_index[1]
    n_edges = src.numel()
    valid_t = (node_type >= 0) & (node_type < num_types)
    tt = node_type.clamp(0, num_types - 1)

    def typed(name, inp):
        w = torch.stack([p["%s.%d.weight" % (name, t)] for t in range(num_types)])   # [T, out, in]
        b = torch.stack([p["%s.%d.bias" % (name, t)] for t in range(num_types)])
        return torch.einsum("ni,noi->no", inp, w[tt]) + b[tt]

    q_all = typed("q_linears", x)                                    # [N, d]
    if use_RTE:
        rt = p["emb.emb.weight"] @ p["emb.lin.weight"].t() + p["emb.lin.bias"]      # [240, d]
        x_src = x[src] + rt[edge_time]
        wk = torch.stack([p["k_linears.%d.weight" % t] for t in range(num_types)])
        bk = torch.stack([p["k_linears.%d.bias" % t] for t in range(num_types)])
        wv = torch.stack([p["v_linears.%d.weight" % t] for t in range(num_types)])
        bv = torch.stack([p["v_linears.%d.bias" % t] for t in range(num_types)])
        ts = tt[src]
        k_e = torch.einsum("ei,eoi->eo", x_src, wk[ts]) + bk[ts]
        v_e = torch.einsum("ei,eoi->eo", x_src, wv[ts]) + bv[ts]
    else:
        k_e = typed("k_linears", x)[src]
        v_e = typed("v_linears", x)[src]
    ok = valid_t[src] & valid_t[dst] & (edge_type >= 0) & (edge_type < num_relations)
    rr = edge_type.clamp(0, num_relations - 1)
    k_e = torch.einsum("eha,ehab->ehb", k_e.view(-1, n_heads, d_k), p["relation_att"][rr])
    v_e = torch.einsum("eha,ehab->ehb", v_e.view(-1, n_heads, d_k), p["relation_msg"][rr])
    s = (q_all[dst].view(-1, n_heads, d_k) * k_e).sum(-1) * p["relation_pri"][rr] / math.sqrt(d_k)
    s = torch.where(ok[:, None], s, torch.zeros_like(s))             # conv.py:68: untouched edges keep 0
    v_e = torch.where(ok[:, None, None], v_e, torch.zeros_like(v_e))  # conv.py:69
    att = torch.zeros(n_edges, n_heads, dtype=torch.float64)
    agg = torch.zeros(n_nodes, n_heads, d_k, dtype=torch.float64)
    order = torch.argsort(dst, stable=True)
    counts = torch.bincount(dst, minlength=n_nodes)
    start = 0
    for n in range(n_nodes):
        c = int(counts[n])
        if c == 0:
            continue
        e = order[start:start + c]
        start += c
        sc = s[e]
        pr = torch.exp(sc - sc.max(0, keepdim=True).values)
        pr = pr / (pr.sum(0, keepdim=True) + 1e-16)
        att[e] = pr
        agg[n] = (pr[:, :, None] * v_e[e]).sum(0)
    g = F.gelu(agg.view(n_nodes, d))
    out = torch.zeros(n_nodes, d, dtype=torch.float64)
    for t in range(num_types):
        sel = node_type == t
        if int(sel.sum()) == 0:
            continue
        o = g[sel] @ p["a_linears.%d.weight" % t].t() + p["a_linears.%d.bias" % t]
        a = torch.sigmoid(p["skip"][t])
        y = o * a + x[sel] * (1 - a)
        if use_norm:
            y = F.layer_norm(y, (d,), p["norms.%d.weight" % t], p["norms.%d.bias" % t], 1e-5)
        out[sel] = y
    return out, att


def init_params(in_dim, out_dim, num_types, num_relations, n_heads, use_norm=True, use_RTE=True, seed=0):
    """Parameter inventory and initialisation of HGTConv.__init__ (conv.py:28-54) and
    RelTemporalEncoding.__init__ (conv.py:287-297), as a name->tensor mapping."""
    g = torch.Generator().manual_seed(seed)
    d_k = out_dim // n_heads
    p = {}

    def lin(prefix, fan_in, fan_out):
        bound = 1.0 / math.sqrt(fan_in)          # nn.Linear default: kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), ..)
        p[prefix + ".weight"] = (torch.rand(fan_out, fan_in, generator=g) * 2 - 1) * bound
        p[prefix + ".bias"] = (torch.rand(fan_out, generator=g) * 2 - 1) * bound

    for t in range(num_types):
        lin("k_linears.%d" % t, in_dim, out_dim)
        lin("q_linears.%d" % t, in_dim, out_dim)
        lin("v_linears.%d" % t, in_dim, out_dim)
        lin("a_linears.%d" % t, out_dim, out_dim)
        if use_norm:
            p["norms.%d.weight" % t] = torch.ones(out_dim)
            p["norms.%d.bias" % t] = torch.zeros(out_dim)
    p["relation_pri"] = torch.ones(num_relations, n_heads)
    a = math.sqrt(6.0 / (d_k + d_k))              # glorot, conv.py:53-54
    p["relation_att"] = (torch.rand(num_relations, n_heads, d_k, d_k, generator=g) * 2 - 1) * a
    p["relation_msg"] = (torch.rand(num_relations, n_heads, d_k, d_k, generator=g) * 2 - 1) * a
    p["skip"] = torch.ones(num_types)
    if use_RTE:
        p["emb.emb.weight"] = rte_sinusoid_table(in_dim)
        lin("emb.lin", in_dim, in_dim)
    return p


def rte_sinusoid_table(n_hid, max_len=240):
    """conv.py:289-294."""
    position = torch.arange(0., max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, n_hid, 2) * -(math.log(10000.0) / n_hid))
    w = torch.empty(max_len, n_hid)
    w[:, 0::2] = torch.sin(position * div_term) / math.sqrt(n_hid)
    w[:, 1::2] = torch.cos(position * div_term) / math.sqrt(n_hid)
    return w
