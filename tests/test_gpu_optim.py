"""FlatAdam (csrc/optim.cu: fmov_grad_gather + fmov_adam_step) against torch.optim.Adam — the optimiser the reference
constructs (exp_runner.py:258-269) — including the reference's rule that only the pose optimisers of the rendered frames
are stepped (exp_runner.py:785-816): inactive groups keep parameters, moments and step counters."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _make(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [[(256, 39), (256, 1), (256,), (217, 256), (1,)], [(64, 256), (64,), (3, 64)], [(64, 256), (64,), (3, 64)]]
    return [[torch.randn(*s, generator=g).to(DEV).requires_grad_(True) for s in grp] for grp in shapes]


def test_flat_adam_equals_torch_adam_over_30_steps_with_inactive_groups():
    from fmov_pose_b200.optim import FlatAdam
    mine, ref = _make(3), _make(3)
    lrs = [5e-4, 3e-4, 1e-3]
    opt = FlatAdam([dict(params=ps, lr=lr) for ps, lr in zip(mine, lrs)])
    topt = [torch.optim.Adam(ps, lr=lr) for ps, lr in zip(ref, lrs)]      # one optimiser per group, as the reference
    g = torch.Generator().manual_seed(9)
    for it in range(30):
        active = [0] + ([1] if it % 3 != 1 else []) + ([2] if it % 4 == 0 else [])
        if it == 12:                          # LR schedule write (TrainStep._write_lr)
            opt.param_groups[0]["lr"].fill_(2e-4)
            for grp in topt[0].param_groups:
                grp["lr"] = 2e-4
        for gi in range(3):
            for a, b in zip(mine[gi], ref[gi]):
                if gi in active:
                    gr = (torch.randn(a.shape, generator=g) * (10.0 ** float(torch.randint(-4, 1, [1], generator=g)))).to(DEV)
                    a.grad, b.grad = gr.clone(), gr.clone()
                else:
                    a.grad, b.grad = None, None
        # a stale gradient on an inactive group must be ignored when the caller names the active groups
        if 2 not in active:
            for a in mine[2]:
                a.grad = torch.ones_like(a)
        opt.step(active)
        for gi in active:
            topt[gi].step()
    torch.cuda.synchronize()
    for gi in range(3):
        for a, b in zip(mine[gi], ref[gi]):
            np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().cpu().numpy(), rtol=2e-6, atol=2e-7)
            st = opt.state_of(a)
            ts = topt[gi].state[b]
            # moments: same recurrences in fp32 (one rounding per step apart from torch's mul_/addcmul_ sequence: 1.3e-5 relative
            # after 30 steps, measured); where exp_avg passes through zero the lerp's rounding is the whole value
            m_ref, v_ref = ts["exp_avg"].cpu().numpy(), ts["exp_avg_sq"].cpu().numpy()
            np.testing.assert_allclose(st["exp_avg"].cpu().numpy(), m_ref, rtol=1e-4, atol=1e-6 * float(np.abs(m_ref).max()))
            np.testing.assert_allclose(st["exp_avg_sq"].cpu().numpy(), v_ref, rtol=1e-4, atol=1e-6 * float(np.abs(v_ref).max()))
            assert float(st["step"]) == float(ts["step"])
    assert [float(s) for s in opt.step_count] == [30.0, 20.0, 8.0]


def test_flat_adam_default_active_set_and_cuda_graph_replay():
    """active_groups=None steps every group that holds a gradient; a captured step replays with new gradients / rates"""
    from fmov_pose_b200.optim import FlatAdam
    mine, ref = _make(5), _make(5)
    opt = FlatAdam([dict(params=ps, lr=1e-3) for ps in mine])
    topt = torch.optim.Adam([p for ps in ref for p in ps], lr=1e-3)
    static = [[torch.zeros_like(p) for p in ps] for ps in mine]
    for ps, gs in zip(mine[:2], static[:2]):           # group 2 never has a gradient
        for p, gr in zip(ps, gs):
            p.grad = gr
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        opt.gather()                                    # warm-up outside the capture
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        opt.step()
    g = torch.Generator().manual_seed(2)
    for it in range(5):
        for gi in range(2):
            for a, b, st in zip(mine[gi], ref[gi], static[gi]):
                gr = torch.randn(a.shape, generator=g).to(DEV) * 1e-2
                st.copy_(gr)
                b.grad = gr.clone()
        graph.replay()
        topt.step()
    torch.cuda.synchronize()
    for gi in range(3):
        for a, b in zip(mine[gi], ref[gi]):
            np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().cpu().numpy(), rtol=2e-6, atol=2e-7)
    assert [float(s) for s in opt.step_count] == [5.0, 5.0, 0.0]


def test_flat_adam_beyond_64_groups():
    """long sequences: one pose MLP per frame -> more parameter groups than the launch-constant bit mask holds"""
    from fmov_pose_b200.optim import FlatAdam
    g = torch.Generator().manual_seed(1)
    groups = [[torch.randn(7, 5, generator=g).to(DEV).requires_grad_(True)] for _ in range(70)]
    ref = [[p.detach().clone().requires_grad_(True) for p in ps] for ps in groups]
    opt = FlatAdam([dict(params=ps, lr=1e-3) for ps in groups])
    topt = [torch.optim.Adam(ps, lr=1e-3) for ps in ref]
    for it, active in enumerate(([0, 3, 69], [0, 65], [0, 3, 69])):
        for gi in active:
            gr = torch.randn(7, 5, generator=g).to(DEV)
            groups[gi][0].grad, ref[gi][0].grad = gr.clone(), gr.clone()
        opt.step(active)
        for gi in active:
            topt[gi].step()
    for gi in range(70):
        np.testing.assert_allclose(groups[gi][0].detach().cpu().numpy(), ref[gi][0].detach().cpu().numpy(), rtol=2e-6, atol=2e-7)
    st = [float(v) for v in opt.step_count]
    assert st[0] == 3.0 and st[3] == 2.0 and st[69] == 2.0 and st[65] == 1.0 and sum(st) == 8.0
