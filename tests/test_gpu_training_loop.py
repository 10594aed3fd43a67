"""End-to-end on the GPU, the way pyHGT's own scripts use the path (OAG/train_paper_field.py:230-255): HGSampling ->
to_torch -> GNN (adapter + HGT layers) -> task head -> loss.backward() -> optimiser step, a NEW sampled graph every
batch, everything through pyhgt_b200 (sampler, ingest with the sync-free plan, native forward and backward).  The loss on
a learnable synthetic task (predict a paper's venue) must go down and every parameter must receive a finite gradient."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.conftest import load_golden          # noqa: E402
from tests.test_sampler import _GraphStub       # noqa: E402


def test_sampled_minibatch_training_reduces_the_loss():
    from pyhgt_b200 import data as hdata, sampler
    from pyhgt_b200.model import GNN
    import pyhgt_b200
    dev = torch.device("cuda:0")
    fx = load_golden("sampler")
    g = _GraphStub(fx)
    fg = sampler.FrozenGraph(g)
    types = g.get_types()
    F_in, n_hid = 32, 64
    rng = np.random.RandomState(0)
    # venue of every paper (the label) from the graph itself: edge_list['venue']['paper']['PV_Journal'][venue][paper]
    n_paper = fg.n_ids["paper"]
    venue_of = np.full(n_paper, -1, dtype=np.int64)
    for v, papers in fx["edge_list"]["venue"]["paper"]["PV_Journal"].items():
        for p in papers:
            venue_of[p] = v
    n_cls = int(venue_of.max()) + 1
    # node features: random per node, papers carry a noisy one-hot of their venue so the task is learnable
    table = {t: rng.randn(fg.n_ids.get(t, 1), F_in).astype(np.float32) * 0.1 for t in types}
    table["paper"][np.arange(n_paper), np.clip(venue_of, 0, None) % F_in] += 1.0

    def extractor(layer_data, graph):
        feature, times, indxs = {}, {}, {}
        for _type in layer_data:
            if len(layer_data[_type]) == 0:
                continue
            idxs = np.array(list(layer_data[_type].keys()))
            feature[_type] = table[_type][idxs]
            times[_type] = np.array(list(layer_data[_type].values()))[:, 1]
            indxs[_type] = idxs
        return feature, times, indxs, []

    years = {}
    for a, papers in fx["edge_list"]["paper"]["author"]["AP_write"].items():
        for _author, t in papers.items():
            years[a] = t
    labelled = np.array([p for p in range(n_paper) if venue_of[p] >= 0 and p in years])
    edge_dict = {e[2]: i for i, e in enumerate(g.get_meta_graph())}
    edge_dict["self"] = len(edge_dict)
    torch.manual_seed(0)
    gnn = GNN(F_in, n_hid, len(types), len(edge_dict), 4, 2, 0.0, "hgt", True, False, True).to(dev).train()
    head = torch.nn.Linear(n_hid, n_cls).to(dev)
    opt = torch.optim.Adam(list(gnn.parameters()) + list(head.parameters()), lr=2e-3)
    old_keep = pyhgt_b200.HGTConv.keep_att
    pyhgt_b200.HGTConv.keep_att = False
    losses = []
    try:
        for step in range(40):
            np.random.seed(step)
            batch = np.random.choice(labelled, 32, replace=False)
            inp = {"paper": np.array([[int(p), int(years[p])] for p in batch])}
            feature, times, edge_list, _, _ = sampler.sample_subgraph(fg, fx["time_range"], 3, 12, inp, extractor)
            nf, nt, etime, ei, et, node_dict, _ = hdata.to_torch(feature, times, edge_list, g, device=dev, prebuild_plan=True)
            out = gnn(nf, nt, etime, ei, et)
            p0 = node_dict["paper"][0]
            logits = head(out[p0:p0 + len(batch)])                    # the seed papers are the first papers (data.py:135-137)
            loss = torch.nn.functional.cross_entropy(logits, torch.from_numpy(venue_of[batch]).to(dev))
            opt.zero_grad()
            loss.backward()
            for name, p in gnn.named_parameters():
                assert p.grad is None or torch.isfinite(p.grad).all(), name
            opt.step()
            losses.append(float(loss))
    finally:
        pyhgt_b200.HGTConv.keep_att = old_keep
    assert np.isfinite(losses).all()
    assert np.mean(losses[-8:]) < 0.7 * np.mean(losses[:8]), losses
