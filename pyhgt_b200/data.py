"""Host-side ingest mirroring pyHGT/data.py:212-256 (``to_torch``) — SURVEY.md §8(f) rank 2.

``to_torch(feature, time, edge_list, graph)`` flattens a sampled sub-graph (the nested dicts/lists that
``sample_subgraph`` returns, data.py:87-210) into the five tensors HGTConv consumes.  The reference does it with
Python loops that append one edge at a time (data.py:240-250); this version builds every <target_type, source_type,
relation> block with one numpy conversion, so the output is IDENTICAL (same node order, same edge order, same dtypes:
FloatTensor / LongTensor, row 0 of edge_index = source) at a fraction of the host time.  With ``device=`` the tensors
are staged through pinned memory to the GPU and the destination-sorted CSR plan of the graph is built right away
(``prebuild_plan``) WITHOUT any host synchronisation: the per-type node counts and the <source type, relation> pairs are
known from the block structure of ``edge_list``, the tile counts stay on the device, and the
range checks run on the host arrays — a new sampled graph per batch (OAG/train_paper_field.py:241) costs kernel
launches only.
"""
import numpy as np
import torch

from .sampler import FrozenGraph, sample_subgraph  # noqa: F401  (data.py:87 lives next to to_torch in the reference)


def to_torch(feature, time, edge_list, graph, device=None, prebuild_plan=False, num_relations=None):
    """Returns (node_feature, node_type, edge_time, edge_index, edge_type, node_dict, edge_dict) exactly like the
    reference (data.py:256)."""
    node_dict = {}
    node_num = 0
    types = graph.get_types()
    for t in types:                                                       # data.py:228-230
        node_dict[t] = [node_num, len(node_dict)]
        node_num += len(feature[t])

    feats, times, ntypes = [], [], []
    for t in types:                                                       # data.py:232-235
        f = np.asarray(feature[t], dtype=np.float32)
        if f.ndim == 1:
            f = f.reshape(len(feature[t]), -1)
        feats.append(f)
        times.append(np.asarray(time[t], dtype=np.int64).reshape(-1))
        ntypes.append(np.full(len(feature[t]), node_dict[t][1], dtype=np.int64))
    width = max((f.shape[1] for f in feats if f.shape[0]), default=0)
    feats = [f if f.shape[0] else np.zeros((0, width), dtype=np.float32) for f in feats]
    node_feature = np.concatenate(feats, 0) if feats else np.zeros((0, 0), dtype=np.float32)
    node_time = np.concatenate(times) if times else np.zeros(0, dtype=np.int64)
    node_type = np.concatenate(ntypes) if ntypes else np.zeros(0, dtype=np.int64)

    edge_dict = {e[2]: i for i, e in enumerate(graph.get_meta_graph())}    # data.py:237-238
    edge_dict['self'] = len(edge_dict)

    src_blocks, dst_blocks, typ_blocks = [], [], []
    block_pairs = set()                                                   # <source type id, relation id> of every block
    for target_type in edge_list:                                         # data.py:240-250, same iteration order
        for source_type in edge_list[target_type]:
            for relation_type in edge_list[target_type][source_type]:
                pairs = edge_list[target_type][source_type][relation_type]
                if len(pairs) == 0:
                    continue
                block_pairs.add((node_dict[source_type][1], edge_dict[relation_type]))
                arr = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)   # [[target_ser, source_ser], ...]
                dst_blocks.append(arr[:, 0] + node_dict[target_type][0])
                src_blocks.append(arr[:, 1] + node_dict[source_type][0])
                typ_blocks.append(np.full(arr.shape[0], edge_dict[relation_type], dtype=np.int64))
    if src_blocks:
        src = np.concatenate(src_blocks)
        dst = np.concatenate(dst_blocks)
        etype = np.concatenate(typ_blocks)
        etime = node_time[dst] - node_time[src] + 120                     # data.py:250
        edge_index = np.stack([src, dst])                                 # row 0 = source (data.py:245,254)
    else:
        edge_index = np.zeros((2, 0), dtype=np.int64)
        etype = np.zeros(0, dtype=np.int64)
        etime = np.zeros(0, dtype=np.int64)

    out = [torch.from_numpy(np.ascontiguousarray(node_feature)), torch.from_numpy(node_type),
           torch.from_numpy(np.ascontiguousarray(etime)), torch.from_numpy(np.ascontiguousarray(edge_index)),
           torch.from_numpy(etype)]
    if device is not None:
        dev = torch.device(device)
        if dev.type == "cuda":
            # The edges arrive grouped in <target type, source type, relation> blocks and the nodes type by type, so the
            # host already knows the per-type counts and the <source type, relation> pairs: the plan is built from that
            # (`host_meta`) with no device read-back, and the id / time ranges are validated here on the host arrays.
            if prebuild_plan and edge_index.shape[1]:
                if edge_index.min() < 0 or edge_index.max() >= node_num:
                    raise IndexError("edge_list contains node indices outside the sampled feature lists")
                if etime.min() < 0 or etime.max() >= 240:
                    raise IndexError("edge_time contains values outside [0, 240) (RelTemporalEncoding table size)")
            out = [t.pin_memory().to(dev, non_blocking=True) for t in out]
            if prebuild_plan:
                from . import plan as _plan
                R = num_relations if num_relations is not None else len(edge_dict)
                meta = {"type_count": [len(feature[t]) for t in types] + [0], "sorted": True,
                        "pairs": sorted(block_pairs)}
                _plan.get_plan(out[1], out[3], out[4], out[2], len(types), R, host_meta=meta)
        else:
            out = [t.to(dev) for t in out]
    node_feature, node_type, edge_time, edge_index, edge_type = out
    return node_feature, node_type, edge_time, edge_index, edge_type, node_dict, edge_dict
