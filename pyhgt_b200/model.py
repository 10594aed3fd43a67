"""Model wrapper mirroring pyHGT/model.py:54-80 (``GNN``): per-type input adapter ``tanh(Linear_t(x))`` followed by a
stack of ``GeneralConv('hgt')`` layers — SURVEY.md §8(f) rank 1.  Parameter names match the reference
(``adapt_ws.{t}.{weight,bias}``, ``gcs.{l}.base_conv.*``) so reference checkpoints load.

The adapter is the same "per-type linear dispatch" as inside HGTConv and runs through the same C-ABI grouped GEMM
(``hgt_typed_linear``: tcgen05 when in_dim >= 64 and n_hid % 16 == 0, fp32 SIMT otherwise); every layer shares
the one cached graph plan.  Under autograd the adapter uses the same GEMM with its native backward
(``autograd._TypedLinear``).
"""
import torch
import torch.nn as nn

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

    def _adapter_table(self, plan, dev):
        key = ("adapter", self.in_dim, self.n_hid)
        table = plan._layer_tables.get(key)
        if table is None:
            groups, cblocks = [], []
            for t in range(self.num_types):
                if plan.type_count[t]:
                    groups.append((plan.type_row0[t], plan.type_count[t], t * self.n_hid, 1, len(cblocks), 1))
                    cblocks.append((plan.type_row0[t] * self.n_hid, self.n_hid))
            table = plan._layer_tables[key] = _plan._pack_groups(groups, cblocks, dev)
        return table

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
        table = self._adapter_table(plan, dev)
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

    def _adapter_autograd(self, node_feature, node_type, edge_index, edge_type, edge_time):
        """Training path of the adapter: the same grouped GEMM with its native backward (autograd._TypedLinear)."""
        from .autograd import typed_linear
        conv0 = self.gcs[0].base_conv
        T = self.num_types
        plan = _plan.get_plan(node_type, edge_index, edge_type, edge_time if conv0.use_RTE else None, T,
                              conv0.num_relations)
        N = plan.n_nodes
        x = node_feature if plan.sorted_types else node_feature.index_select(0, plan.perm.long())
        table = self._adapter_table(plan, node_feature.device)
        w_cat = torch.cat([l.weight for l in self.adapt_ws], 0)
        b_cat = torch.cat([l.bias for l in self.adapt_ws], 0)
        n_known = plan.type_row0[T]
        res = typed_linear(x, w_cat, b_cat, table, self.n_hid, N * self.n_hid, conv0.linear_impl, 0,
                           ((n_known * self.n_hid, N * self.n_hid),)).view(N, self.n_hid)
        res = torch.cat([torch.tanh(res[:n_known]), res[n_known:]], 0) if n_known < N else torch.tanh(res)   # model.py:75
        if not plan.sorted_types:
            res = res.index_select(0, plan.rank.long())
        return res

    def forward(self, node_feature, node_type, edge_time, edge_index, edge_type):
        grad = torch.is_grad_enabled() and (node_feature.requires_grad or any(p.requires_grad for p in self.parameters()))
        if not node_feature.is_cuda:
            raise _lib.HgtError("pyhgt_b200.GNN runs on CUDA tensors only (got %s): there is no CPU fallback"
                                % node_feature.device)
        if grad:
            res = self._adapter_autograd(node_feature, node_type, edge_index, edge_type, edge_time)
        else:
            res = self._adapter_cuda(node_feature, node_type, edge_index, edge_type, edge_time)
        meta_xs = self.drop(res)
        del res
        for gc in self.gcs:
            meta_xs = gc(meta_xs, node_type, edge_index, edge_type, edge_time)
        return meta_xs
