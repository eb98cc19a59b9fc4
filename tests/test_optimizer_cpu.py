"""CPU: host-side logic of FusedAdamW (model/shape_engine.py:75-136 optimizer construction; torch.optim.AdamW semantics) with the two
launch wrappers replaced by torch restatements of the kernels' arithmetic: parameter groups, per-parameter step counters, the
capturable mode's device-side scalars, and the capture / prepare_replay / mark_updated bookkeeping GraphedTrainStep relies on."""
import numpy as np
import torch


def _fake_adamw(params, grads, ms, vs, lr, b1, b2, eps, wd, bc1, bc2_sqrt):
    for p, g, m, v in zip(params, grads, ms, vs):
        p.mul_(1.0 - lr * wd)
        m.mul_(b1).add_(g, alpha=1.0 - b1)
        v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
        p.sub_((lr / bc1) * (m / (v.sqrt() / bc2_sqrt + eps)))


def _install(monkeypatch, recorded):
    from zeroshape_b200 import ops

    def multi(params, grads, ms, vs, lr, b1, b2, eps, wd, step):
        _fake_adamw(params, grads, ms, vs, lr, b1, b2, eps, wd, 1.0 - b1 ** step, float(np.sqrt(1.0 - b2 ** step)))

    def multi_dev(params, grads, ms, vs, hyper):
        def launch():
            lr, b1, b2, eps, wd, bc1, bc2 = [float(x) for x in hyper[:7]]
            _fake_adamw(params, grads, ms, vs, lr, b1, b2, eps, wd, bc1, bc2)
        if recorded is not None and recorded.get("capturing"):
            recorded["launches"].append(launch)          # a capture records, it does not execute
        else:
            launch()
    monkeypatch.setattr(ops, "adamw_step_multi", multi)
    monkeypatch.setattr(ops, "adamw_step_multi_dev", multi_dev)


def _models(seed=0, n=7):
    g = torch.Generator().manual_seed(seed)
    ws = [torch.randn(int(k), generator=g) for k in torch.randint(1, 50, (n,), generator=g)]
    return [torch.nn.Parameter(w.clone()) for w in ws], [torch.nn.Parameter(w.clone()) for w in ws], g


def _groups(ps):
    return [dict(params=ps[:4], lr=3e-3, weight_decay=0.05), dict(params=ps[4:], lr=1e-3, weight_decay=0.0)]


def test_fused_adamw_groups_and_late_parameters(monkeypatch):
    """Two groups with their own lr / weight decay, a parameter that receives its first gradient at step 3 (own bias correction),
    an lr change in between, state_dict round trip: the same numbers as torch.optim.AdamW, in both launch modes."""
    from zeroshape_b200.model.shape.implicit_train import FusedAdamW
    _install(monkeypatch, None)
    for capturable in (False, True):
        pa, pb, g = _models(1)
        oa = FusedAdamW(_groups(pa), betas=(0.9, 0.95), capturable=capturable)
        ob = torch.optim.AdamW(_groups(pb), betas=(0.9, 0.95))
        for it in range(5):
            for i, (a, b) in enumerate(zip(pa, pb)):
                if i == 2 and it < 2:
                    a.grad = b.grad = None               # joins late
                    continue
                gr = torch.randn(a.shape, generator=g)
                a.grad, b.grad = gr.clone(), gr.clone()
            if it == 3:
                oa.param_groups[1]["lr"] = ob.param_groups[1]["lr"] = 5e-4
            v0 = pa[0]._version
            oa.step(); ob.step()
            assert pa[0]._version > v0                   # packed-weight caches key on the version counter
        assert max((a - b).abs().max().item() for a, b in zip(pa, pb)) < 1e-6
        assert oa.state[pa[2]]["step"] == 3 and oa.state[pa[0]]["step"] == 5
        sd = oa.state_dict()
        oc = FusedAdamW(_groups(pa), betas=(0.9, 0.95), capturable=capturable)
        oc.load_state_dict(sd)
        assert int(oc.state[pa[0]]["step"]) == 5 and torch.equal(oc.state[pa[1]]["exp_avg"], oa.state[pa[1]]["exp_avg"])


def test_fused_adamw_capture_bookkeeping(monkeypatch):
    """What GraphedTrainStep does, without a GPU: eager steps, then a step() under 'capture' (records the launches, must NOT advance
    the step counters or touch the scalars), then per replay prepare_replay() -> recorded launches -> mark_updated()."""
    from zeroshape_b200.model.shape.implicit_train import FusedAdamW
    rec = {"capturing": False, "launches": []}
    _install(monkeypatch, rec)
    pa, pb, g = _models(2)
    oa = FusedAdamW(_groups(pa), betas=(0.9, 0.95), capturable=True)
    ob = torch.optim.AdamW(_groups(pb), betas=(0.9, 0.95))
    grads = [torch.zeros_like(p) for p in pa]            # static gradient buffers, as in a captured iteration

    def new_grads():
        for gbuf, a, b in zip(grads, pa, pb):
            gbuf.copy_(torch.randn(a.shape, generator=g))
            a.grad, b.grad = gbuf, gbuf.clone()
    for _ in range(2):                                    # warm-up iterations
        new_grads(); oa.step(); ob.step()
    # the capture
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: True)
    rec["capturing"] = True
    before = [p.detach().clone() for p in pa]
    new_grads(); oa.step()
    rec["capturing"] = False
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
    assert all(torch.equal(a, b) for a, b in zip(before, pa)) and all(oa.state[p]["step"] == 2 for p in pa)
    assert len(rec["launches"]) == 2 and len(oa._captured) == 2
    # replays
    for it in range(3):
        if it == 1:
            oa.param_groups[0]["lr"] = ob.param_groups[0]["lr"] = 1e-3       # a scheduler step between replays
        new_grads()
        v0 = pa[0]._version
        oa.prepare_replay()
        for launch in rec["launches"]:
            launch()
        oa.mark_updated()
        ob.step()
        assert pa[0]._version > v0 and all(oa.state[p]["step"] == 3 + it for p in pa)
    assert max((a - b).abs().max().item() for a, b in zip(pa, pb)) < 1e-6
