"""Training path of pyhgt_b200.HGTConv (reference: the same forward differentiated by autograd,
OAG/train_paper_field.py:249 ``loss.backward()``).

The graph-dependent, memory-bound part — score / softmax-by-destination / weighted aggregation — is ONE custom
autograd.Function whose forward and backward are the hand-written kernels ``hgt_edge_forward`` /
``hgt_edge_backward`` on the cached CSR plan (no per-edge intermediates are kept: backward recomputes the
softmax weights from the saved per-destination (max, sum)).  The typed linears (Q / K' / V' projection, a_linears)
run FORWARD through the same tcgen05 grouped GEMM as inference (``_TypedLinear``); their backward (dX, dW, db) uses
cuBLAS fp32 GEMMs on strided views this round.  The relation-matrix fold and the gated-skip / LayerNorm are small
differentiable torch ops, so gradients reach ``node_inp`` and every parameter of conv.py:28-54 including ``emb.*``.  The inference path
(``torch.no_grad``) does not come through here: it uses the fused tcgen05 / epilogue kernels in conv.py.
"""
import ctypes
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from . import plan as _plan


def _stream():
    return torch.cuda.current_stream().cuda_stream


class _EdgeAttention(torch.autograd.Function):
    """agg[i] = sum_{e -> i} softmax_i(<Q[i], K'[e]>) * V'[e]   (conv.py:99,108-111 + scatter-add)."""

    @staticmethod
    def forward(ctx, q, kv, kvr, plan, n_heads, want_att, variant):
        N, d = q.shape
        dev = q.device
        ws_bytes = ctypes.c_size_t()
        _lib.call("hgt_edge_workspace_bytes", plan.n_split, d, n_heads, ctypes.byref(ws_bytes))
        ws = torch.empty(ws_bytes.value, dtype=torch.uint8, device=dev)
        agg = torch.empty((N, d), dtype=torch.float32, device=dev)
        stats = torch.empty((N, 2 * n_heads), dtype=torch.float32, device=dev)
        att = torch.empty((plan.n_edges, n_heads), dtype=torch.float32, device=dev) if want_att else None
        q, kv = q.contiguous(), kv.contiguous()
        kvr = None if kvr is None else kvr.contiguous()
        _lib.call("hgt_edge_forward", q.data_ptr(), kv.data_ptr(), _lib.ptr(kvr), plan.row_ptr.data_ptr(),
                  plan.kv_row.data_ptr(), None if kvr is None else plan.rte_row.data_ptr(), plan.csr_eid.data_ptr(),
                  plan.tiles.data_ptr(), plan.n_tiles, plan.n_split, plan.hubs.data_ptr(), plan.n_hubs, N, plan.n_edges, d, n_heads, 0,
                  agg.data_ptr(), _lib.ptr(att), stats.data_ptr(), None, None, ws.data_ptr(), ws.numel(), variant,
                  _stream())
        ctx.plan, ctx.n_heads, ctx.has_kvr = plan, n_heads, kvr is not None
        ctx.save_for_backward(q, kv, kvr if kvr is not None else q.new_empty(0), agg, stats)
        ctx.mark_non_differentiable(*([att] if att is not None else []))
        return (agg, att) if att is not None else (agg, None)

    @staticmethod
    def backward(ctx, dagg, _datt=None):
        q, kv, kvr, agg, stats = ctx.saved_tensors
        plan, H = ctx.plan, ctx.n_heads
        kvr = kvr if ctx.has_kvr else None
        N, d = q.shape
        dq = torch.zeros_like(q)
        dkv = torch.zeros_like(kv)
        dkvr = torch.zeros_like(kvr) if kvr is not None else None
        ws = torch.empty(256, dtype=torch.uint8, device=q.device)
        dagg = dagg.contiguous()
        _lib.call("hgt_edge_backward", q.data_ptr(), kv.data_ptr(), _lib.ptr(kvr), agg.data_ptr(), dagg.data_ptr(),
                  stats.data_ptr(), plan.row_ptr.data_ptr(), plan.kv_row.data_ptr(),
                  None if kvr is None else plan.rte_row.data_ptr(), plan.tiles.data_ptr(), plan.n_tiles, N, d, H,
                  dq.data_ptr(), dkv.data_ptr(), _lib.ptr(dkvr), ws.data_ptr(), ws.numel(), _stream())
        return dq, dkv, dkvr, None, None, None, None


class _TypedLinear(torch.autograd.Function):
    """Grouped typed linear through the C ABI (``hgt_typed_linear``: tcgen05 split-bf16 when the shape allows).
    forward : out_flat[cblock c of group g][m, :] = A[rows_g] @ W_cat[rows of (g, c)]^T + b_cat
    backward: dA / dW_cat / db_cat with cuBLAS fp32 GEMMs on strided views of the flat gradient buffer (the
              native split-bf16 dX / dW kernels are future work)."""

    @staticmethod
    def forward(ctx, a, w_cat, b_cat, module, table, width, out_elems, impl):
        g_dev, g_host, n_g, c_dev = table
        a = a.contiguous()
        w_cat = w_cat.contiguous()
        out = torch.zeros(out_elems, dtype=torch.float32, device=a.device)
        module._typed_linear(a, a.shape[1], w_cat, b_cat, a.shape[1], width, table, out, impl, _stream())
        ctx.table, ctx.width, ctx.has_bias = table, width, b_cat is not None
        ctx.cblocks_host = table.c_host                                # host copy kept with the table (plan._pack_groups)
        ctx.save_for_backward(a, w_cat)
        return out

    @staticmethod
    def backward(ctx, dout):
        a, w_cat = ctx.saved_tensors
        _, g_host, n_g, _ = ctx.table
        width, K = ctx.width, a.shape[1]
        da = torch.zeros_like(a)
        dw = torch.zeros_like(w_cat)
        db = torch.zeros(w_cat.shape[0], dtype=torch.float32, device=a.device) if ctx.has_bias else None
        dout = dout.contiguous()
        for gi in range(n_g):
            g = g_host[gi]
            r0, m = int(g["a_row0"]), int(g["m"])
            if m == 0:
                continue
            a_g = a[r0:r0 + m]
            for c in range(int(g["n_cblocks"])):
                cb = ctx.cblocks_host[int(g["cb_first"]) + c]
                off, ld = int(cb["out_off"]), int(cb["ld"])
                d_blk = torch.as_strided(dout, (m, width), (ld, 1), off)            # [m, width] view
                w0 = int(g["w_row0"]) + c * width
                da[r0:r0 + m].addmm_(d_blk, w_cat[w0:w0 + width])
                dw[w0:w0 + width].addmm_(d_blk.t(), a_g)
                if db is not None and int(g["has_bias"]):
                    db[w0:w0 + width] += d_blk.sum(0)
        return da, dw, db, None, None, None, None, None


def _fold(w, b, rel, scale, H, dk):
    """W'[h*dk+c, :] = scale[h] * sum_a rel[h,a,c] * W[h*dk+a, :]  (and the same for the bias): the per-head
    right-multiplication of conv.py:98/104 moved into the weights."""
    d_in = w.shape[1]
    wf = torch.einsum("hac,hai->hci", rel, w.view(H, dk, d_in))
    bf = torch.einsum("hac,ha->hc", rel, b.view(H, dk))
    if scale is not None:
        wf = wf * scale.view(H, 1, 1)
        bf = bf * scale.view(H, 1)
    return wf.reshape(H * dk, d_in), bf.reshape(H * dk)


def hgt_conv_autograd(m, node_inp, node_type, edge_index, edge_type, edge_time):
    dev = node_inp.device
    d_in, d, H, T, R, dk = m.in_dim, m.out_dim, m.n_heads, m.num_types, m.num_relations, m.d_k
    plan = _plan.get_plan(node_type, edge_index, edge_type, edge_time if m.use_RTE else None, T, R)
    N, P = plan.n_nodes, plan.n_pairs
    if node_inp.shape[0] != N:
        raise ValueError("node_inp has %d rows but node_type has %d" % (node_inp.shape[0], N))
    x = node_inp if plan.sorted_types else node_inp.index_select(0, plan.perm.long())
    rows = [slice(plan.type_row0[t], plan.type_row0[t + 1]) for t in range(T)]

    # typed projections with the relation matrices folded in.  W_cat / b_cat are assembled with differentiable torch
    # ops (tiny: the fold is O(P * d * d_in * d_k)); the big grouped GEMM runs through the C ABI (_TypedLinear).
    lt = _plan.layer_tables(plan, d_in, d)
    pairs_of_type = [[] for _ in range(T)]
    for p, (s_, _) in enumerate(plan.pairs):
        pairs_of_type[s_].append(p)
    folded = {}
    for p, (s_, r) in enumerate(plan.pairs):
        kw, kb = _fold(m.k_linears[s_].weight, m.k_linears[s_].bias, m.relation_att[r],
                       m.relation_pri[r] / math.sqrt(dk), H, dk)
        vw, vb = _fold(m.v_linears[s_].weight, m.v_linears[s_].bias, m.relation_msg[r], None, H, dk)
        folded[p] = (torch.cat([kw, vw], 0), torch.cat([kb, vb], 0))            # [2d, d_in]: K' rows then V' rows
    w_parts, b_parts = [], []
    for t in range(T):                                                          # same row order as plan.layer_tables
        w_parts.append(m.q_linears[t].weight)
        b_parts.append(m.q_linears[t].bias)
        for p in pairs_of_type[t]:
            w_parts.append(folded[p][0])
            b_parts.append(folded[p][1])
    w_cat, b_cat = torch.cat(w_parts, 0), torch.cat(b_parts, 0)
    proj = _TypedLinear.apply(x, w_cat, b_cat, m, lt.proj_groups, d, lt.proj_elems, m.linear_impl)
    q = proj[lt.q_off:lt.q_off + N * d].view(N, d)
    kv = proj[lt.kv_off:lt.kv_off + (plan.kv_rows + 1) * 2 * d].view(plan.kv_rows + 1, 2 * d)
    kvr = None
    if m.use_RTE:
        rt = F.linear(m.emb.emb.weight, m.emb.lin.weight, m.emb.lin.bias)        # [240, d_in], conv.py:299
        kvr = torch.cat([F.linear(rt, folded[p][0]) for p in range(P)] + [x.new_zeros(1, 2 * d)], 0)

    tail = N - plan.type_row0[T]                                                 # nodes of unknown type (conv.py:120)
    agg, att = _EdgeAttention.apply(q, kv, kvr, plan, H, bool(m.keep_att), m.edge_variant)
    m.att = att

    g = F.gelu(agg)                                                              # conv.py:119
    wa_cat = torch.cat([l.weight for l in m.a_linears], 0)
    ba_cat = torch.cat([l.bias for l in m.a_linears], 0)
    o_all = _TypedLinear.apply(g, wa_cat, ba_cat, m, lt.upd_groups, d, N * d, m.linear_impl).view(N, d)
    outs = []
    for t in range(T):
        o = m.drop(o_all[rows[t]])                                               # conv.py:125
        alpha = torch.sigmoid(m.skip[t])                                         # conv.py:129
        y = o * alpha + x[rows[t]] * (1 - alpha)
        if m.use_norm:
            y = m.norms[t](y)
        outs.append(y)
    if tail:
        outs.append(x.new_zeros(tail, d))                                        # conv.py:120: rows stay zero
    out = torch.cat(outs, 0)
    if not plan.sorted_types:
        out = out.index_select(0, plan.rank.long())
    return out
