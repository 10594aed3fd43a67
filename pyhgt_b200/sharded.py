"""Multi-GPU HGTConv: 1-D destination-node sharding with one all-to-all of halo source rows per layer
(the reference itself is single-GPU — pyHGT has no distributed code).

Partition (built once per graph, identically on every rank from the full COO):
  * inside each node type the nodes are cut into `world` contiguous blocks balanced on the edge-kernel cost
    2*in_degree + 1, so every rank owns a slice of every type (typed GEMMs stay balanced) and ~E/world edges;
  * rank g keeps the in-edges of its owned destinations; the sources of those edges that it does not own
    are its halo.  Local node numbering = [owned (type-sorted) | halo (by owner, then global id)].
Per layer: gather the owned rows peers asked for -> ONE all_to_all_single (NCCL over NVLink on GPUs, gloo in
the CPU tests) straight into the tail of the local feature buffer -> the ordinary single-GPU kernels run on the
local graph, with Q / a_linear / update restricted to the owned rows and K'/V' projected for owned + halo.
"""
from dataclasses import dataclass

import torch
import torch.distributed as dist


def partition_owner(node_type, edge_index, num_types, world):
    """owner[n] in [0, world): contiguous cost-balanced blocks inside every node type (deterministic, CPU)."""
    n = node_type.numel()
    deg = torch.bincount(edge_index[1], minlength=n)
    cost = 2 * deg + 1
    owner = torch.zeros(n, dtype=torch.int64)
    for t in range(num_types):
        ids = (node_type == t).nonzero(as_tuple=True)[0]           # ascending ids = stable type order
        if ids.numel() == 0:
            continue
        c = cost[ids].cumsum(0)
        total = int(c[-1])
        # node i goes to block floor(prefix_before_i * world / total)
        before = c - cost[ids]
        owner[ids] = torch.clamp((before * world) // max(total, 1), max=world - 1)
    other = (node_type < 0) | (node_type >= num_types)
    if other.any():
        ids = other.nonzero(as_tuple=True)[0]
        owner[ids] = torch.arange(ids.numel()) * world // max(ids.numel(), 1)
    return owner


@dataclass
class ShardedGraph:
    rank: int
    world: int
    device: torch.device
    n_owned: int
    n_halo: int
    n_local_edges: int
    owned_global: torch.Tensor        # [n_owned] global ids (CPU)
    halo_global: torch.Tensor         # [n_halo] global ids (CPU)
    node_type: torch.Tensor           # [n_owned + n_halo] local node types (device)
    edge_index: torch.Tensor          # [2, E_local] local ids (device)
    edge_type: torch.Tensor
    edge_time: torch.Tensor
    send_idx: torch.Tensor            # [n_send] local owned positions to send, grouped by destination rank
    send_splits: list
    recv_splits: list
    active_per_type: list             # owned nodes of each type (type T = out-of-range bucket)
    num_types: int
    num_relations: int
    group: object = None

    @staticmethod
    def build(node_type, edge_index, edge_type, edge_time, num_types, num_relations, rank, world, device,
              group=None):
        node_type, edge_index, edge_type = node_type.cpu(), edge_index.cpu(), edge_type.cpu()
        edge_time = None if edge_time is None else edge_time.cpu()
        n = node_type.numel()
        owner = partition_owner(node_type, edge_index, num_types, world)
        tkey = torch.where((node_type >= 0) & (node_type < num_types), node_type, torch.full_like(node_type, num_types))
        mine = (owner == rank).nonzero(as_tuple=True)[0]
        owned = mine[torch.argsort(tkey[mine], stable=True)]                 # type-sorted owned ids
        e_sel = (owner[edge_index[1]] == rank).nonzero(as_tuple=True)[0]    # in-edges of owned destinations
        src, dst = edge_index[0, e_sel], edge_index[1, e_sel]
        srcs = torch.unique(src)
        halo = srcs[owner[srcs] != rank]
        halo = halo[torch.argsort(owner[halo] * n + halo)]                   # by owner, then id
        local_of = torch.full((n,), -1, dtype=torch.int64)
        local_of[owned] = torch.arange(owned.numel())
        local_of[halo] = owned.numel() + torch.arange(halo.numel())
        ei_local = torch.stack([local_of[src], local_of[dst]])
        nt_local = torch.cat([node_type[owned], node_type[halo]])
        recv_splits = torch.bincount(owner[halo], minlength=world).tolist()
        # what every peer asks of me: replay the same deterministic construction for each destination rank
        send_lists = []
        dst_owner_all = owner[edge_index[1]]
        src_owner_all = owner[edge_index[0]]
        for p in range(world):
            if p == rank:
                send_lists.append(torch.zeros(0, dtype=torch.int64))
                continue
            need = (dst_owner_all == p) & (src_owner_all == rank)
            ids = torch.unique(edge_index[0, need.nonzero(as_tuple=True)[0]])   # ascending = the peer's halo order
            send_lists.append(local_of[ids])
        send_splits = [int(s.numel()) for s in send_lists]
        send_idx = torch.cat(send_lists) if send_lists else torch.zeros(0, dtype=torch.int64)
        active = torch.bincount(tkey[owned], minlength=num_types + 1).tolist()
        return ShardedGraph(rank=rank, world=world, device=device, n_owned=int(owned.numel()),
                            n_halo=int(halo.numel()), n_local_edges=int(e_sel.numel()), owned_global=owned,
                            halo_global=halo, node_type=nt_local.to(device), edge_index=ei_local.to(device),
                            edge_type=edge_type[e_sel].to(device),
                            edge_time=None if edge_time is None else edge_time[e_sel].to(device),
                            send_idx=send_idx.to(device), send_splits=send_splits, recv_splits=recv_splits,
                            active_per_type=active, num_types=num_types, num_relations=num_relations, group=group)

    # --------------------------------------------------------------------------------------------
    def exchange(self, x_own):
        """[n_owned, d] owned rows -> [n_owned + n_halo, d] local rows (owned first), one all-to-all."""
        d = x_own.shape[1]
        x_local = torch.empty((self.n_owned + self.n_halo, d), dtype=x_own.dtype, device=x_own.device)
        x_local[:self.n_owned].copy_(x_own)
        if x_own.is_cuda and self.send_idx.numel() > 0:
            from . import _lib
            send = torch.empty((self.send_idx.numel(), d), dtype=x_own.dtype, device=x_own.device)
            if getattr(self, "_send_idx32", None) is None:
                self._send_idx32 = self.send_idx.to(torch.int32)
            _lib.call("hgt_gather_rows", x_own.contiguous().data_ptr(), self._send_idx32.data_ptr(),
                      self.send_idx.numel(), d, send.data_ptr(), torch.cuda.current_stream().cuda_stream)
        else:
            send = x_own.index_select(0, self.send_idx)
        recv = x_local[self.n_owned:]
        if self.world > 1:
            dist.all_to_all_single(recv, send, self.recv_splits, self.send_splits, group=self.group)
        return x_local

    def forward(self, conv, x_own, edge_time_used=True):
        """One HGTConv layer on this rank's shard; returns the [n_owned, d] output rows (owned_global order)."""
        x_local = self.exchange(x_own)
        out, att, _ = conv._forward_impl(x_local, self.node_type, self.edge_index, self.edge_type,
                                         self.edge_time if conv.use_RTE else None, want_att=False, save=False,
                                         active_per_type=self.active_per_type)
        return out[:self.n_owned]
