"""Fused clip + Adam step (optim.cu) against torch.nn.utils.clip_grad_norm_ + torch.optim.Adam on the same gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("wd,max_norm", [(5e-4, 5.0), (0.0, 0.0), (1e-2, 0.05)])
def test_fused_clip_adam_matches_torch(wd, max_norm):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from eeg_gnn_ssl_b200.dist import FlatGradSync
    from eeg_gnn_ssl_b200.optim import FusedClipAdam
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    shapes = [(492, 128), (128,), (492, 64), (64,), (1, 64), (1,), (7, 3)]         # odd sizes: padded slices
    ours = [torch.nn.Parameter(torch.randn(s, generator=g).to(dev)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    sync = FlatGradSync(ours, world_size=1, align=4)
    opt = FusedClipAdam(ours, lr=3e-3, weight_decay=wd, max_grad_norm=max_norm, grad_sync=sync)
    topt = torch.optim.Adam(ref, lr=3e-3, weight_decay=wd)
    for it in range(6):
        sync.zero()
        scale = 10.0 if it % 2 else 0.01                                           # clipped and unclipped steps
        for p, r in zip(ours, ref):
            gr = (scale * torch.randn(p.shape, generator=g)).to(dev)
            p.grad.copy_(gr)
            r.grad = gr.clone()
        n_ref = torch.nn.utils.clip_grad_norm_(ref, max_norm) if max_norm > 0 else \
            torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(r.grad) for r in ref]))
        topt.step()
        n_ours = opt.step()
        assert abs(float(n_ours) - float(n_ref)) <= 1e-5 * float(n_ref)
        for p, r in zip(ours, ref):
            err = float((p.detach() - r.detach()).abs().max() / r.detach().abs().max().clamp_min(1e-12))
            assert err < 5e-6, (it, tuple(p.shape), err)
            assert p.data_ptr() >= opt.flat.data_ptr()                              # still views of the flat buffer
    assert int(opt.step_count) == 6
