"""CUDA-graph replay of the fused ResidualVQ level loop (ResidualVQ.enable_cuda_graph) against the eager loop:
identical outputs, indices, losses and codebook buffers on every step, including steps on which dead codes are
replaced (the expiry stays outside the graph)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(dev, thr):
    from vqb200 import CodebookParams, ResidualVQ
    torch.manual_seed(0)
    m = ResidualVQ(dim=64, num_quantizers=4,
                   codebook_params=CodebookParams(dim=64, codebook_size=96, threshold_ema_dead_code=thr)).to(dev).train()
    g = torch.Generator().manual_seed(3)
    for li, layer in enumerate(m.layers):
        c = (torch.randn(1, 96, 64, generator=g) * (0.6 / 1.5 ** li)).to(dev)
        cb = layer._codebook
        cb.embeddings.copy_(c); cb.embed_avg.copy_(c); cb.cluster_size.fill_(1.0); cb.invalidate_cache()
    return m


@pytest.mark.parametrize("thr", [0, 2])
def test_graph_replay_equals_eager(thr):
    dev = torch.device("cuda:0")
    eager, graphed = _make(dev, thr), _make(dev, thr)
    graphed.enable_cuda_graph()
    g = torch.Generator().manual_seed(11)
    x_e = torch.empty(3, 700, 64, device=dev)
    x_g = torch.empty(3, 700, 64, device=dev)          # one address for every step: eager, capture, replay, replay
    for step in range(5):
        x = torch.randn(3, 700, 64, generator=g)
        x_e.copy_(x); x_g.copy_(x)
        torch.manual_seed(100 + step)                  # same draws for the dead-code replacement
        with torch.no_grad():
            qe, ie, le = eager(x_e)
        torch.manual_seed(100 + step)
        with torch.no_grad():
            qg, ig, lg = graphed(x_g)
        assert torch.equal(qe, qg), f"step {step}: quantized differs"
        assert torch.equal(ie, ig), f"step {step}: indices differ"
        assert torch.equal(le, lg), f"step {step}: losses differ"
        for a, b in zip(eager.layers, graphed.layers):
            for name in ("embeddings", "embed_avg", "cluster_size"):
                assert torch.equal(getattr(a._codebook, name), getattr(b._codebook, name)), (step, name)
    ent = [v for v in graphed._graphs.values() if v != "warm"]
    assert len(ent) == 1, "the loop was never captured"
