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


def _toy(dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(s, generator=g).to(dev)) for s in [(37, 16), (16,), (5, 3)]]


def test_fused_clip_adam_is_a_torch_optimizer_with_checkpointable_state():
    """what the reference's loop does with its optimiser: CosineAnnealingLR(optimizer), optimizer.param_groups[0]['lr']
    (train.py:224,282), optimizer.state_dict() / load_state_dict (utils.py:141,160) -- and resuming from a checkpoint
    written by torch.optim.Adam continues identically"""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from eeg_gnn_ssl_b200.optim import FusedClipAdam
    dev = torch.device("cuda:0")
    ours, ref = _toy(dev), _toy(dev)
    opt = FusedClipAdam(ours, lr=3e-3, weight_decay=5e-4, max_grad_norm=5.0)
    topt = torch.optim.Adam(ref, lr=3e-3, weight_decay=5e-4)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=4)
    tsched = torch.optim.lr_scheduler.CosineAnnealingLR(topt, T_max=4)
    g = torch.Generator().manual_seed(9)

    def one(o, ps, grads, clip):
        o.zero_grad()
        for p, gr in zip(ps, grads):
            if p.grad is None:
                p.grad = gr.clone()
            else:
                p.grad.copy_(gr)
        if clip:
            torch.nn.utils.clip_grad_norm_(ps, 5.0)
        o.step()

    for it in range(3):
        grads = [torch.randn(p.shape, generator=g).to(dev) for p in ours]
        one(opt, ours, grads, False)
        one(topt, ref, grads, True)
        sched.step(); tsched.step()
        assert abs(opt.param_groups[0]["lr"] - topt.param_groups[0]["lr"]) < 1e-12
    # state_dict has torch.optim.Adam's layout and values
    sd, tsd = opt.state_dict(), topt.state_dict()
    assert set(sd["state"].keys()) == set(tsd["state"].keys())
    for k in tsd["state"]:
        assert float(sd["state"][k]["step"]) == float(tsd["state"][k]["step"]) == 3.0
        for name in ("exp_avg", "exp_avg_sq"):
            a, b = sd["state"][k][name], tsd["state"][k][name]
            assert a.shape == b.shape and float((a - b).abs().max()) <= 5e-5 * float(b.abs().max()), (name, float((a - b).abs().max()), float(b.abs().max()))
    # resume: a fresh fused optimiser loads the TORCH optimiser's checkpoint and continues like torch does
    ours2 = [torch.nn.Parameter(r.detach().clone()) for r in ref]
    opt2 = FusedClipAdam(ours2, lr=1.0, weight_decay=5e-4, max_grad_norm=5.0)
    opt2.load_state_dict(tsd)
    assert abs(opt2.param_groups[0]["lr"] - topt.param_groups[0]["lr"]) < 1e-12 and int(opt2.step_count) == 3
    grads = [torch.randn(p.shape, generator=g).to(dev) for p in ours]
    one(opt2, ours2, grads, False)
    one(topt, ref, grads, True)
    for p, r in zip(ours2, ref):
        assert float((p.detach() - r.detach()).abs().max()) < 5e-6 * float(r.detach().abs().max())
    # and the other way round: torch.optim.Adam loads the fused optimiser's state_dict
    topt2 = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in ours2], lr=1.0)
    topt2.load_state_dict(opt2.state_dict())
    assert float(topt2.state_dict()["state"][0]["step"]) == 4.0


def test_fused_clip_adam_detects_rebound_storage():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from eeg_gnn_ssl_b200.optim import FusedClipAdam
    dev = torch.device("cuda:0")
    ps = _toy(dev)
    opt = FusedClipAdam(ps, lr=1e-3)
    ps[0].grad = None                                   # what module.zero_grad(set_to_none=True) does
    with pytest.raises(RuntimeError, match="flat gradient buffer"):
        opt.step()
    ps2 = _toy(dev)
    opt2 = FusedClipAdam(ps2, lr=1e-3)
    ps2[1].data = ps2[1].data.clone()                   # what a later model.to()/.cuda() amounts to
    with pytest.raises(RuntimeError, match="flat buffer"):
        opt2.step()


def test_grad_scale_folds_the_data_parallel_average():
    """grad_scale = 1/world inside the fused pass == scaling the summed gradient first"""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from eeg_gnn_ssl_b200.dist import FlatGradSync
    from eeg_gnn_ssl_b200.optim import FusedClipAdam
    dev = torch.device("cuda:0")
    a, b = _toy(dev, 1), _toy(dev, 1)
    sa = FlatGradSync(a, world_size=1, align=4)
    sb = FlatGradSync(b, world_size=1, align=4)
    oa = FusedClipAdam(a, lr=1e-2, max_grad_norm=0.5, grad_sync=sa)
    ob = FusedClipAdam(b, lr=1e-2, max_grad_norm=0.5, grad_sync=sb)
    g = torch.Generator().manual_seed(4)
    for _ in range(3):
        gr = torch.randn(sa.flat.shape, generator=g).to(dev)
        sa.flat.copy_(gr * 0.25)                        # averaged beforehand
        sb.flat.copy_(gr); sb.world = 4; sb._scaled = False     # summed over 4 ranks, scale left to the optimiser
        na, nb = float(oa.step()), float(ob.step())
        assert abs(na - nb) <= 1e-6 * na
        for p, q in zip(a, b):
            assert float((p.detach() - q.detach()).abs().max()) <= 2e-6 * float(p.detach().abs().max())
