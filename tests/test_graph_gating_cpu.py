"""Host logic of ResidualVQ.enable_cuda_graph: when the fused level loop may be replayed as a CUDA graph (no device
needed: `_graph_usable` only looks at module state and the process-group state)."""
import torch

from vqb200 import CodebookParams, ResidualVQ, rvq as RVQ


def _make(shared=False, sync=False):
    torch.manual_seed(0)
    m = ResidualVQ(dim=16, num_quantizers=3, codebook_params=CodebookParams(dim=16, codebook_size=32),
                   shared_codebook=shared, sync_codebook=sync).train()
    for layer in m.layers:
        layer._codebook.is_initialized = True
    return m


def test_graph_is_opt_in_and_respects_mask_grad_and_shared_codebook():
    x = torch.randn(2, 8, 16)
    m = _make()
    assert not m._graph_usable(x, None)                       # never without enable_cuda_graph()
    m.enable_cuda_graph()
    assert m._graph_usable(x, None)
    assert not m._graph_usable(x, torch.ones(2, 8, dtype=torch.bool))
    with torch.enable_grad():
        assert not m._graph_usable(x.clone().requires_grad_(True), None)
    m.enable_cuda_graph(False)
    assert not m._graph_usable(x, None) and m._graphs == {}
    shared = _make(shared=True).enable_cuda_graph()
    assert not shared._graph_usable(x, None)                  # one codebook for all levels: its refresh is read again
    un_init = _make().enable_cuda_graph()
    un_init.layers[1]._codebook.is_initialized = False
    assert not un_init._graph_usable(x, None)                 # kmeans init draws on the host


def test_graph_under_data_parallelism_needs_the_promise(monkeypatch):
    x = torch.randn(2, 8, 16)
    monkeypatch.setattr(RVQ.dist, "is_available", lambda: True)
    monkeypatch.setattr(RVQ.dist, "is_initialized", lambda: True)
    monkeypatch.setattr(RVQ.dist, "get_world_size", lambda *a, **k: 2)
    m = _make(sync=True)
    for layer in m.layers:
        layer._codebook.use_ddp = True
    m.enable_cuda_graph()
    assert not m._graph_usable(x, None)                       # a rank replaying while another captures would dead-lock
    m.enable_cuda_graph(data_parallel=True)
    assert m._graph_usable(x, None)
    local = _make(sync=False).enable_cuda_graph()             # no statistics all_reduce in the loop: nothing to promise
    for layer in local.layers:
        layer._codebook.use_ddp = False
    assert local._graph_usable(x, None)
