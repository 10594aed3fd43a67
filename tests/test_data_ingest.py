"""pyhgt_b200.data.to_torch against the reference's own to_torch (pyHGT/data.py:212-256):
the golden fixture was produced by the unmodified reference on a synthetic sampled sub-graph (oracle/make_golden.py);
outputs must be IDENTICAL (node order, edge order, dtypes, dict contents)."""
import time

import pytest
import torch

from oracle import pyg_shim
from pyhgt_b200 import data as hdata
from tests.conftest import load_golden


class _GraphStub:
    """What to_torch needs from the reference's Graph: get_types() and get_meta_graph()."""

    def __init__(self, types, metas):
        self._t, self._m = list(types), [tuple(m) for m in metas]

    def get_types(self):
        return self._t

    def get_meta_graph(self):
        return self._m


def test_to_torch_is_identical_to_reference_golden():
    fx = load_golden("to_torch")
    g = _GraphStub(fx["types"], fx["meta_graph"])
    out = hdata.to_torch(fx["feature"], fx["time"], fx["edge_list"], g)
    for got, key in zip(out[:5], ("node_feature", "node_type", "edge_time", "edge_index", "edge_type")):
        ref = fx[key]
        assert got.dtype == ref.dtype and got.shape == ref.shape, key
        assert torch.equal(got, ref), key
    assert out[5] == fx["node_dict"] and out[6] == fx["edge_dict"]
    assert out[3].dtype == torch.int64 and out[0].dtype == torch.float32


def test_to_torch_empty_edge_list():
    fx = load_golden("to_torch")
    g = _GraphStub(fx["types"], fx["meta_graph"])
    out = hdata.to_torch(fx["feature"], fx["time"], {}, g)
    assert out[3].shape == (2, 0) and out[4].numel() == 0 and out[2].numel() == 0
    assert torch.equal(out[0], fx["node_feature"])


@pytest.mark.skipif(not pyg_shim.reference_available(), reason="reference tree only exists in the dev container")
def test_to_torch_matches_live_reference_and_is_faster():
    from oracle import make_golden as mg
    data, g, feature, times, edge_list = mg.synthetic_sampled_subgraph(seed=5, n_per_type=(3000, 2000, 300), n_edges=60000)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter(); ref = data.to_torch(feature, times, edge_list, g); t_ref = time.perf_counter() - t0
    t0 = time.perf_counter(); out = hdata.to_torch(feature, times, edge_list, g); t_new = time.perf_counter() - t0
    for a, b in zip(out[:5], ref[:5]):
        assert torch.equal(a, b)
    assert out[5] == ref[5] and out[6] == ref[6]
    assert t_new < t_ref, "vectorised ingest (%.3f s) should beat the reference's per-edge loop (%.3f s)" % (t_new, t_ref)
