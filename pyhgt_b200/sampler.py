"""HGSampling — the budget sampler that feeds the hot path (reference: pyHGT/data.py:87-210 ``sample_subgraph``;
SURVEY.md §8(f) rank 4).  Host-side by nature (a sequential budget process driven by numpy's global RNG); this version
keeps the reference's semantics EXACTLY — the same draws from numpy's global RNG in the same order, the same
dict-iteration orders — so that with the same seed it returns the same ``(feature, times, edge_list, indxs, texts)``,
and ``to_torch`` then emits the same tensors.  What changes is the data structure:

  * ``FrozenGraph`` — the reference's 5-level ``edge_list[target_type][source_type][relation][target_id][source_id] =
    time`` dict-of-dicts flattened ONCE into CSR arrays per <target type, source type, relation> (neighbour ids and times
    in dict insertion order, id -> row map), the form a device-side ingest would consume;
  * the budget ``{source_id: [score, time]}`` per type becomes three flat arrays (score, time, insertion stamp) — the
    insertion stamp reproduces ``list(budget.keys())`` order, including pop-and-re-insert;
  * the final "reconstruct the sampled adjacency" triple loop (data.py:190-209: every sampled target x every neighbour,
    membership tests in Python dicts) becomes one gather + mask per <target type, source type, relation>.
  * ``add_budget`` (data.py:108-130) runs for a whole batch of newly sampled nodes in one native call
    (``hgt_sampler_add_budget``, csrc/sampler.cu, host code): the uniform draws do not depend on the budget, so numpy
    makes them first, in the reference's target-major order, and the native loop applies them; the key order of the
    budget dict comes from an append-only insertion log instead of a sort.  ``_sample_slices`` is the same process with
    one update per adjacency slice (native or numpy) — the fallback when the library is not built.

The sampled ``edge_list`` values are ``[E_block, 2]`` int64 arrays of ``[target_ser, source_ser]`` rows (the reference
builds Python lists of pairs with the same content and order); ``pyhgt_b200.data.to_torch`` and the reference's
``to_torch`` accept both.
"""
import ctypes as _c
from collections import defaultdict
from itertools import chain as _chain

import numpy as np

_NO_TIME = np.iinfo(np.int64).min          # stands for `None` in the neighbour-time arrays (data.py:125-126)


class _CBlock(_c.Structure):                 # hgt_sampler_block (include/hgt_b200.h)
    _fields_ = [("row_of", _c.c_void_p), ("n_row_of", _c.c_int64), ("ptr", _c.c_void_p), ("nbr", _c.c_void_p),
                ("time", _c.c_void_p), ("src_state", _c.c_int32), ("skip", _c.c_int32)]


class _CState(_c.Structure):                 # hgt_sampler_state
    _fields_ = [("n", _c.c_int64), ("in_layer", _c.c_void_p), ("in_budget", _c.c_void_p), ("score", _c.c_void_p),
                ("b_time", _c.c_void_p), ("stamp", _c.c_void_p), ("log", _c.c_void_p), ("log_len", _c.c_int64),
                ("layer_seq", _c.c_int64), ("budget_seq", _c.c_int64)]


class _Block:
    """One <target type, source type, relation> adjacency in CSR form, dict insertion order preserved."""
    __slots__ = ("row_of", "ptr", "nbr", "time", "has_none", "nbr_addr", "time_addr", "ptr_list", "row_list")

    def __init__(self, tesr, n_target_ids):
        adls = list(tesr.values())
        n_keys = len(adls)
        self.row_of = np.full(n_target_ids, -1, dtype=np.int64)
        counts = np.fromiter(map(len, adls), dtype=np.int64, count=n_keys)
        self.ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        if n_keys:
            self.row_of[np.fromiter(tesr.keys(), dtype=np.int64, count=n_keys)] = np.arange(n_keys, dtype=np.int64)
        total = int(self.ptr[-1])
        # one pass over the whole block (iterating a dict yields its keys = the neighbour ids, in insertion order)
        self.nbr = np.fromiter(_chain.from_iterable(adls), dtype=np.int64, count=total)
        has_none = False
        try:
            self.time = np.fromiter(_chain.from_iterable(map(dict.values, adls)), dtype=np.int64, count=total)
        except TypeError:                                  # some edge times are None (data.py:125-126)
            has_none = True
            self.time = np.fromiter((_NO_TIME if v is None else v for adl in adls for v in adl.values()),
                                    dtype=np.int64, count=total)
        self.has_none = has_none
        self.nbr_addr = self.nbr.ctypes.data               # base addresses for the native budget update
        self.time_addr = self.time.ctypes.data
        self.ptr_list = self.ptr.tolist()                  # plain ints: the per-node lookups stay out of numpy
        self.row_list = self.row_of.tolist()

    def row(self, target_id):
        if target_id < 0 or target_id >= len(self.row_list):
            return -1
        return self.row_list[target_id]


class FrozenGraph:
    """CSR snapshot of a reference ``Graph`` (pyHGT/data.py:19-84).  Keeps the dict iteration orders the sampler's
    results depend on: target types, source types per target type, relations per pair."""

    def __init__(self, graph):
        self.graph = graph
        n_ids = defaultdict(int)
        for t_t, d1 in graph.edge_list.items():
            for s_t, d2 in d1.items():
                for r, tesr in d2.items():
                    if tesr:
                        n_ids[t_t] = max(n_ids[t_t], max(tesr.keys()) + 1)
        self.blocks = {}                                  # target_type -> source_type -> relation -> _Block (ordered)
        for t_t, d1 in graph.edge_list.items():
            self.blocks[t_t] = {}
            for s_t, d2 in d1.items():
                self.blocks[t_t][s_t] = {}
                for r, tesr in d2.items():
                    blk = self.blocks[t_t][s_t][r] = _Block(tesr, n_ids.get(t_t, 0))
                    if blk.nbr.size:
                        n_ids[s_t] = max(n_ids[s_t], int(blk.nbr.max()) + 1)
        self.n_ids = dict(n_ids)
        self.types = []                                   # every type that occurs, targets first (index = native state slot)
        for t_t, d1 in graph.edge_list.items():
            for ty in [t_t] + list(d1.keys()):
                if ty not in self.types:
                    self.types.append(ty)
        self.type_idx = {ty: i for i, ty in enumerate(self.types)}
        self._cblocks = {}

    def ensure_ids(self, _type, max_id):
        if max_id + 1 > self.n_ids.get(_type, 0):
            self.n_ids[_type] = max_id + 1

    def native_blocks(self, target_type):
        """(ctypes array of hgt_sampler_block, [(source type index, _Block, skip)], ...) of one target type, blocks in the
        reference's dict order (data.py:113,115); None when the type has no in-edges."""
        ent = self._cblocks.get(target_type)
        if ent is None and target_type in self.blocks:
            lst = [(self.type_idx[s_t], blk, 1 if r == 'self' else 0)
                   for s_t, tes in self.blocks[target_type].items() for r, blk in tes.items()]
            arr = (_CBlock * max(len(lst), 1))()
            for i, (si, blk, skip) in enumerate(lst):
                arr[i].row_of, arr[i].n_row_of = blk.row_of.ctypes.data, blk.row_of.shape[0]
                arr[i].ptr, arr[i].nbr, arr[i].time = blk.ptr.ctypes.data, blk.nbr_addr, blk.time_addr
                arr[i].src_state, arr[i].skip = si, skip
            ent = self._cblocks[target_type] = (arr, lst)
        return ent


class _TypeState:
    """layer_data[type] and budget[type] of the reference as flat arrays over node ids."""

    def __init__(self, n):
        self.in_layer = np.zeros(n, dtype=bool)
        self.ser = np.full(n, -1, dtype=np.int64)
        self.layer_time = np.zeros(n, dtype=np.int64)
        self.layer_ids = []                                # insertion order == ser order
        self.in_budget = np.zeros(n, dtype=bool)
        self.score = np.zeros(n, dtype=np.float64)
        self.b_time = np.zeros(n, dtype=np.int64)
        self.stamp = np.zeros(n, dtype=np.int64)
        self.log = np.empty(n, dtype=np.int64)             # ids in budget-insertion order (written by the native path)
        self._addr()

    def _addr(self):
        self.addr = (self.in_layer.shape[0], self.in_layer.ctypes.data, self.in_budget.ctypes.data,
                     self.score.ctypes.data, self.b_time.ctypes.data, self.stamp.ctypes.data)

    def grow(self, n):
        if n <= self.in_layer.shape[0]:
            return
        def ext(a, fill):
            b = np.full(n, fill, dtype=a.dtype)
            b[:a.shape[0]] = a
            return b
        self.in_layer = ext(self.in_layer, False)
        self.ser = ext(self.ser, -1)
        self.layer_time = ext(self.layer_time, 0)
        self.in_budget = ext(self.in_budget, False)
        self.score = ext(self.score, 0.0)
        self.b_time = ext(self.b_time, 0)
        self.stamp = ext(self.stamp, 0)
        self.log = ext(self.log, 0)
        self._addr()

    def to_c(self, c):
        c.n = self.in_layer.shape[0]
        c.in_layer, c.in_budget, c.score = self.in_layer.ctypes.data, self.in_budget.ctypes.data, self.score.ctypes.data
        c.b_time, c.stamp, c.log = self.b_time.ctypes.data, self.stamp.ctypes.data, self.log.ctypes.data


_NATIVE = [None, False]      # [function, looked up]
_NATIVE_BATCH = [None, False]


def _native_batch():
    """hgt_sampler_add_budget (whole add_budget for a batch of targets) or None when the library is not built."""
    if not _NATIVE_BATCH[1]:
        _NATIVE_BATCH[1] = True
        try:
            from . import _lib
            _NATIVE_BATCH[0] = _lib.load().hgt_sampler_add_budget
        except Exception:                                   # noqa: BLE001 — optional accelerator of a host-side routine
            _NATIVE_BATCH[0] = None
    return _NATIVE_BATCH[0]


def _native_update():
    """hgt_sampler_budget_update from libhgt_b200.so (host code), or None when the library is not built: the pure numpy
    path below gives the same result."""
    if not _NATIVE[1]:
        _NATIVE[1] = True
        try:
            from . import _lib
            _NATIVE[0] = _lib.load().hgt_sampler_budget_update
        except Exception:                                   # noqa: BLE001 — optional accelerator of a host-side routine
            _NATIVE[0] = None
    return _NATIVE[0]


def sample_subgraph(graph, time_range, sampled_depth=2, sampled_number=8, inp=None, feature_extractor=None):
    """Drop-in for pyHGT/data.py:87 ``sample_subgraph``.  `graph` is a reference ``Graph`` or a ``FrozenGraph`` of it
    (freeze once, sample many batches).  Consumes numpy's global RNG exactly like the reference."""
    fg = graph if isinstance(graph, FrozenGraph) else FrozenGraph(graph)
    if _native_batch() is not None and all(t in fg.type_idx for t in inp):
        return _sample_batched(fg, time_range, sampled_depth, sampled_number, inp, feature_extractor)
    return _sample_slices(fg, time_range, sampled_depth, sampled_number, inp, feature_extractor)


def _sample_batched(fg, time_range, sampled_depth, sampled_number, inp, feature_extractor):
    """One native call per <sampling layer, node type>: the whole add_budget of the batch of newly sampled nodes runs in
    csrc/sampler.cu; numpy only makes the random draws (same order as the reference, see hgt_sampler_add_budget)."""
    nat = _native_batch()
    max_time = int(np.max(list(time_range.keys())))
    T = len(fg.types)
    cst = (_CState * max(T, 1))()
    for i in range(T):
        cst[i].layer_seq = cst[i].budget_seq = -1
    sts = [None] * T
    counters = np.zeros(3, dtype=np.int64)                # budget stamp, layer_data first touch, budget first touch

    def state(i):
        st = sts[i]
        if st is None:
            st = sts[i] = _TypeState(fg.n_ids.get(fg.types[i], 0))
            st.to_c(cst[i])
        return st

    def type_index(_type):
        i = fg.type_idx.get(_type)
        if i is None:
            raise KeyError("sample_subgraph: node type %r does not occur in graph.edge_list" % (_type,))
        return i

    def touch_layer(i):
        if cst[i].layer_seq < 0:
            cst[i].layer_seq = int(counters[1])
            counters[1] += 1

    def add_layer(i, _id, _time):
        st = state(i)
        touch_layer(i)
        if _id >= st.in_layer.shape[0]:
            fg.ensure_ids(fg.types[i], _id)
            st.grow(fg.n_ids[fg.types[i]])
            st.to_c(cst[i])
        st.ser[_id] = len(st.layer_ids)                   # re-adding an id overwrites [ser, time] (dict semantics)
        if not st.in_layer[_id]:
            st.layer_ids.append(_id)
        st.in_layer[_id] = True
        st.layer_time[_id] = _time

    def add_budget(i, ids, tms):
        ent = fg.native_blocks(fg.types[i])
        if ent is None or ids.shape[0] == 0:
            return
        arr, lst = ent
        nb = len(lst)
        need = []
        for b, (si, blk, skip) in enumerate(lst):
            state(si)
            if skip:
                continue
            inside = (ids >= 0) & (ids < blk.row_of.shape[0])
            rows = np.where(inside, blk.row_of[np.where(inside, ids, 0)], -1)
            cnt = np.where(rows >= 0, blk.ptr[rows + 1] - blk.ptr[rows], 0)
            for t in np.nonzero((cnt >= sampled_number) & (cnt > 0))[0]:
                need.append((int(t), b, int(cnt[t])))
        off_ptr = pos_ptr = None
        if need:
            need.sort()                                   # the reference draws target after target, block after block
            draw_off = np.full(ids.shape[0] * nb, -1, dtype=np.int64)
            chunks = []
            for k, (t, b, n_adl) in enumerate(need):
                # == np.random.choice(list(adl.keys()), sampled_number, replace=False): RandomState.choice draws
                # permutation(len(a))[:size] whether `a` is the population or its size, so the stream is the same
                chunks.append(np.random.choice(n_adl, sampled_number, replace=False))
                draw_off[t * nb + b] = k * sampled_number
            draw_pos = np.ascontiguousarray(np.concatenate(chunks), dtype=np.int64)
            off_ptr, pos_ptr = draw_off.ctypes.data, draw_pos.ctypes.data
        rc = nat(ids.ctypes.data, tms.ctypes.data, ids.shape[0], arr, nb, cst, T, sampled_number, off_ptr, pos_ptr,
                 _NO_TIME, max_time, counters.ctypes.data)
        if rc != 0:
            raise RuntimeError("hgt_sampler_add_budget failed (%d): neighbour id outside the frozen graph's id range" % rc)

    # first adding the sampled nodes then updating budget (data.py:135-141)
    for _type in inp:
        i = type_index(_type)
        for _id, _time in inp[_type]:
            add_layer(i, int(_id), int(_time))
    for _type in inp:
        arr_in = np.asarray(inp[_type], dtype=np.int64).reshape(-1, 2)
        add_budget(type_index(_type), np.ascontiguousarray(arr_in[:, 0]), np.ascontiguousarray(arr_in[:, 1]))

    for _layer in range(sampled_depth):                   # data.py:147
        order = sorted((i for i in range(T) if cst[i].budget_seq >= 0), key=lambda i: cst[i].budget_seq)
        for i in order:                                   # sts = list(budget.keys())
            st = sts[i]
            log = st.log[:cst[i].log_len]
            keys = log[st.in_budget[log]]                 # == list(budget[source_type].keys()): insertion order, popped ids gone
            if sampled_number > len(keys):
                sampled_ids = np.arange(len(keys))
            else:
                score = st.score[keys] ** 2
                score = score / np.sum(score)
                sampled_ids = np.random.choice(len(score), sampled_number, p=score, replace=False)
            sampled_keys = np.ascontiguousarray(keys[sampled_ids])
            tms = np.ascontiguousarray(st.b_time[sampled_keys])
            touch_layer(i)                                # data.py:166-167 (ids are new to the layer and distinct)
            st.ser[sampled_keys] = len(st.layer_ids) + np.arange(sampled_keys.shape[0])
            st.in_layer[sampled_keys] = True
            st.layer_time[sampled_keys] = tms
            st.layer_ids.extend(sampled_keys.tolist())
            add_budget(i, sampled_keys, tms)              # data.py:168-170
            st.in_budget[sampled_keys] = False            # budget[source_type].pop(k)

    states = {fg.types[i]: sts[i] for i in range(T) if sts[i] is not None}
    layer_order = [fg.types[i] for i in sorted((i for i in range(T) if cst[i].layer_seq >= 0),
                                               key=lambda i: cst[i].layer_seq)]
    return _finish(fg, states, layer_order, feature_extractor)


def _sample_slices(fg, time_range, sampled_depth, sampled_number, inp, feature_extractor):
    """The same process with one (optionally native) budget update per adjacency slice — used when the library does not
    export the batched entry point, and kept as a second implementation the tests compare with."""
    ref_graph = fg.graph
    max_time = np.max(list(time_range.keys()))
    states = {}                       # per type arrays
    layer_order = []                  # key order of the reference's `layer_data` defaultdict
    budget_order = []                 # key order of the reference's `budget` defaultdict
    stamp_counter = np.zeros(1, dtype=np.int64)
    touched = np.zeros(1, dtype=np.int32)
    upd = _native_update()
    stamp_addr, touched_addr = stamp_counter.ctypes.data, touched.ctypes.data

    def state(_type, touch_layer=False):
        st = states.get(_type)
        if st is None:
            st = states[_type] = _TypeState(fg.n_ids.get(_type, 0))
        if touch_layer and _type not in layer_seen:
            layer_seen.add(_type)
            layer_order.append(_type)
        return st

    layer_seen = set()
    budget_seen = set()

    def add_layer(_type, _id, _time):
        st = state(_type, touch_layer=True)
        if _id >= st.in_layer.shape[0]:
            fg.ensure_ids(_type, _id)
            st.grow(fg.n_ids[_type])
        st.ser[_id] = len(st.layer_ids)                   # re-adding an id overwrites [ser, time] (dict semantics)
        if not st.in_layer[_id]:
            st.layer_ids.append(_id)
        st.in_layer[_id] = True
        st.layer_time[_id] = _time

    def add_budget(target_type, target_id, target_time):
        te = fg.blocks.get(target_type)
        if te is None:
            return
        for source_type, tes in te.items():               # data.py:113
            for relation_type, blk in tes.items():        # data.py:115
                if relation_type == 'self':
                    continue
                row = blk.row(target_id)
                if row < 0:
                    continue
                a, b = blk.ptr_list[row], blk.ptr_list[row + 1]
                n_adl = b - a
                if n_adl == 0:
                    continue
                st = state(source_type)
                ids = tms = None
                if n_adl < sampled_number:                # data.py:119-122: take the whole adjacency
                    n_s = n_adl
                    ids_addr, tms_addr = blk.nbr_addr + 8 * a, blk.time_addr + 8 * a
                else:
                    # == np.random.choice(list(adl.keys()), sampled_number, replace=False): RandomState.choice draws
                    # permutation(len(a))[:size] whether `a` is the population or its size, so the stream is the same
                    pos = np.random.choice(n_adl, sampled_number, replace=False)
                    ids = np.ascontiguousarray(blk.nbr[a:b][pos])
                    tms = np.ascontiguousarray(blk.time[a:b][pos])
                    n_s = ids.shape[0]
                    ids_addr, tms_addr = ids.ctypes.data, tms.ctypes.data
                if upd is not None:
                    # one native call (csrc/sampler.cu) instead of a dozen small numpy operations
                    touched[0] = 0
                    n_arr, p_layer, p_budget, p_score, p_time, p_stamp = st.addr
                    kept = upd(ids_addr, tms_addr, n_s, int(target_time), _NO_TIME, int(max_time), n_arr, p_layer,
                               p_budget, p_score, p_time, p_stamp, stamp_addr, touched_addr)
                    if kept >= 0:
                        if touched[0]:
                            state(source_type, touch_layer=True)
                        if kept and source_type not in budget_seen:
                            budget_seen.add(source_type)
                            budget_order.append(source_type)
                        continue
                # numpy path: library not built, or an id past the arrays (grow them and redo this slice)
                if ids is None:
                    ids, tms = blk.nbr[a:b], blk.time[a:b]
                if blk.has_none:
                    tms = np.where(tms == _NO_TIME, target_time, tms)
                late = tms > max_time                     # data.py:127 (short-circuit `or`: layer_data[source_type] is
                if late.all():                            # only touched when some candidate passes the time test)
                    continue
                st = state(source_type, touch_layer=True)
                if int(ids.max()) >= st.in_layer.shape[0]:
                    fg.ensure_ids(source_type, int(ids.max()))
                    st.grow(fg.n_ids[source_type])
                keep = ~late & ~st.in_layer[ids]
                if not keep.any():
                    continue
                if source_type not in budget_seen:        # budget[source_type] is created by its first real update
                    budget_seen.add(source_type)
                    budget_order.append(source_type)
                kid, ktm = ids[keep], tms[keep]
                new = ~st.in_budget[kid]
                n_new = int(new.sum())
                if n_new:
                    st.stamp[kid[new]] = int(stamp_counter[0]) + np.arange(n_new)
                    stamp_counter[0] += n_new
                    st.in_budget[kid[new]] = True
                    st.score[kid[new]] = 0.0
                st.score[kid] += 1.0 / n_s                # data.py:129 (ids are unique inside one adjacency)
                st.b_time[kid] = ktm                      # data.py:130

    # first adding the sampled nodes then updating budget (data.py:135-141)
    for _type in inp:
        for _id, _time in inp[_type]:
            add_layer(_type, _id, _time)
    for _type in inp:
        for _id, _time in inp[_type]:
            add_budget(_type, _id, _time)

    for _layer in range(sampled_depth):                   # data.py:146
        for source_type in list(budget_order):
            st = states[source_type]
            cand = np.nonzero(st.in_budget)[0]
            keys = cand[np.argsort(st.stamp[cand], kind="stable")]          # == list(budget[source_type].keys())
            if sampled_number > len(keys):
                sampled_ids = np.arange(len(keys))
            else:
                score = st.score[keys] ** 2
                score = score / np.sum(score)
                sampled_ids = np.random.choice(len(score), sampled_number, p=score, replace=False)
            sampled_keys = keys[sampled_ids]
            for k in sampled_keys:                        # data.py:166-167
                add_layer(source_type, int(k), int(st.b_time[k]))
            for k in sampled_keys:                        # data.py:168-170
                add_budget(source_type, int(k), int(st.b_time[k]))
                st.in_budget[k] = False                   # budget[source_type].pop(k)

    return _finish(fg, states, layer_order, feature_extractor)


def _finish(fg, states, layer_order, feature_extractor):
    """layer_data -> features (data.py:174) and the sampled adjacency (data.py:181-209)."""
    ref_graph = fg.graph
    # hand the reference-shaped layer_data to the feature extractor (data.py:174)
    layer_data = defaultdict(lambda: {})
    for _type in layer_order:
        st = states[_type]
        d = layer_data[_type]
        for _id in st.layer_ids:
            d[_id] = [int(st.ser[_id]), int(st.layer_time[_id])]
    feature, times, indxs, texts = feature_extractor(layer_data, ref_graph)

    edge_list = defaultdict(lambda: defaultdict(lambda: defaultdict(lambda: [])))
    for _type in layer_order:                             # data.py:181-184 'self' loops
        st = states[_type]
        n = len(st.layer_ids)
        if n:
            sers = st.ser[np.asarray(st.layer_ids, dtype=np.int64)]
            edge_list[_type][_type]['self'] = np.stack([sers, sers], 1)
    # reconstruct the sampled adjacency (data.py:190-209), one gather + mask per block
    for target_type, te in fg.blocks.items():
        tst = states.get(target_type)
        if tst is None or not tst.layer_ids:
            continue
        tids = np.asarray(tst.layer_ids, dtype=np.int64)
        for source_type, tes in te.items():
            sst = states.get(source_type)
            if sst is None or not sst.layer_ids:
                continue
            for relation_type, blk in tes.items():
                rows = blk.row_of[tids[tids < blk.row_of.shape[0]]]
                tsel = tids[tids < blk.row_of.shape[0]][rows >= 0]
                rows = rows[rows >= 0]
                if rows.size == 0:
                    continue
                a, b = blk.ptr[rows], blk.ptr[rows + 1]
                cnt = b - a
                total = int(cnt.sum())
                if total == 0:
                    continue
                owner = np.repeat(np.arange(rows.size), cnt)
                offs = np.arange(total) - np.repeat(np.cumsum(cnt) - cnt, cnt)
                nb = blk.nbr[a[owner] + offs]
                ok = nb < sst.in_layer.shape[0]
                ok[ok] = sst.in_layer[nb[ok]]
                if not ok.any():
                    continue
                pairs = np.stack([tst.ser[tsel[owner[ok]]], sst.ser[nb[ok]]], 1)
                cur = edge_list[target_type][source_type][relation_type]
                edge_list[target_type][source_type][relation_type] = pairs if len(cur) == 0 else np.concatenate([cur, pairs])
    return feature, times, edge_list, indxs, texts
