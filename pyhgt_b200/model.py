"""Model wrapper mirroring pyHGT/model.py:54-80 (``GNN``): per-type input adapter ``tanh(Linear_t(x))`` followed by a
stack of ``GeneralConv('hgt')`` layers — SURVEY.md §8(f) rank 1.  Parameter names match the reference
(``adapt_ws.{t}.{weight,bias}``, ``gcs.{l}.base_conv.*``) so reference checkpoints load.

The adapter is the same "per-type linear dispatch" as inside HGTConv and runs through the same C-ABI grouped GEMM
(``hgt_typed_linear``: tcgen05 when in_dim >= 64 and n_hid % 16 == 0, fp32 SIMT otherwise); every layer shares
the one cached graph plan.  Inference (no_grad) only; under autograd the adapter falls back to per-type F.linear.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from . import plan as _plan
from .conv import GeneralConv, HGTConv


class GNN(nn.Module):
    def __init__(self, in_dim, n_hid, num_types, num_relations, n_heads, n_layers, dropout=0.2, conv_name='hgt',
                 prev_norm=False, last_norm=False, use_RTE=True):
        super().__init__()
        self.gcs = nn.ModuleList()
        self.num_types = num_types
        self.in_dim = in_dim
        self.n_hid = n_hid
        self.adapt_ws = nn.ModuleList()
        self.drop = nn.Dropout(dropout)
        for _ in range(num_types):
            self.adapt_ws.append(nn.Linear(in_dim, n_hid))
        for _ in range(n_layers - 1):
            self.gcs.append(GeneralConv(conv_name, n_hid, n_hid, num_types, num_relations, n_heads, dropout,
                                        use_norm=prev_norm, use_RTE=use_RTE))
        self.gcs.append(GeneralConv(conv_name, n_hid, n_hid, num_types, num_relations, n_heads, dropout,
                                    use_norm=last_norm, use_RTE=use_RTE))
        self._ptrs = {}
        for gc in self.gcs[:-1]:                                   # every layer but the last feeds another projection
            if isinstance(gc.base_conv, HGTConv) and type(gc.base_conv) is HGTConv:
                gc.base_conv.emit_split = True

    def _adapter_cuda(self, node_feature, node_type, edge_index, edge_type, edge_time):
        conv0 = self.gcs[0].base_conv
        T = self.num_types
        plan = _plan.get_plan(node_type, edge_index, edge_type, edge_time if conv0.use_RTE else None, T,
                              conv0.num_relations)
        dev, N = node_feature.device, plan.n_nodes
        st = torch.cuda.current_stream().cuda_stream
        x = node_feature.contiguous()
        if not plan.sorted_types:
            xs = torch.empty_like(x)
            _lib.call("hgt_gather_rows", x.data_ptr(), plan.perm.data_ptr(), N, self.in_dim, xs.data_ptr(), st)
            x = xs
        key = ("adapter", self.in_dim, self.n_hid)
        table = plan._layer_tables.get(key)
        if table is None:
            groups, cblocks = [], []
            for t in range(T):
                if plan.type_count[t]:
                    groups.append((plan.type_row0[t], plan.type_count[t], t * self.n_hid, 1, len(cblocks), 1))
                    cblocks.append((plan.type_row0[t] * self.n_hid, self.n_hid))
            table = plan._layer_tables[key] = _plan._pack_groups(groups, cblocks, dev)
        w_cat = torch.empty((T * self.n_hid, self.in_dim), dtype=torch.float32, device=dev)
        b_cat = torch.empty(T * self.n_hid, dtype=torch.float32, device=dev)
        wp = conv0._ptrs("adapt_w", [l.weight for l in self.adapt_ws], dev)
        bp = conv0._ptrs("adapt_b", [l.bias for l in self.adapt_ws], dev)
        _lib.call("hgt_concat_linears", wp.data_ptr(), bp.data_ptr(), T, self.n_hid, self.in_dim, w_cat.data_ptr(),
                  b_cat.data_ptr(), st)
        res = torch.zeros((N, self.n_hid), dtype=torch.float32, device=dev)     # unknown-type rows stay 0 (model.py:70)
        conv0._typed_linear(x, self.in_dim, w_cat, b_cat, self.in_dim, self.n_hid, table, res, conv0.linear_impl, st)
        n_known = plan.type_row0[T]
        res[:n_known].tanh_()                                                    # model.py:75
        if not plan.sorted_types:
            res = res.index_select(0, plan.rank.long())
        return res

    def forward(self, node_feature, node_type, edge_time, edge_index, edge_type):
        grad = torch.is_grad_enabled() and (node_feature.requires_grad or any(p.requires_grad for p in self.parameters()))
        if node_feature.is_cuda and not grad and isinstance(self.gcs[0].base_conv, HGTConv):
            res = self._adapter_cuda(node_feature, node_type, edge_index, edge_type, edge_time)
        else:
            res = torch.zeros(node_feature.size(0), self.n_hid, device=node_feature.device)
            for t_id in range(self.num_types):
                idx = (node_type == int(t_id))
                if idx.sum() == 0:
                    continue
                res[idx] = torch.tanh(self.adapt_ws[t_id](node_feature[idx]))
        meta_xs = self.drop(res)
        del res
        for gc in self.gcs:
            meta_xs = gc(meta_xs, node_type, edge_index, edge_type, edge_time)
        return meta_xs
