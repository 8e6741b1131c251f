"""Keyed optimizers (name-addressed state, in-place load, sharded leaves) and stage-wise learning-rate schedules."""
import math

import pytest
import torch

from torchrec_b200.optim.keyed import CombinedOptimizer, KeyedOptimizer, KeyedOptimizerWrapper, OptimizerWrapper, StateMismatch, export_state, import_state
from torchrec_b200.optim.warmup import WarmupOptimizer, WarmupPolicy, WarmupStage, lr_multiplier


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))


def _keyed(m, lr=0.1):
    return KeyedOptimizerWrapper(dict(m.named_parameters()), lambda ps: torch.optim.Adam(ps, lr=lr))


def _train(m, opt, n=3):
    for i in range(n):
        opt.zero_grad()
        m(torch.ones(2, 4) * (i + 1)).sum().backward()
        opt.step()


def test_state_is_keyed_by_name_and_loads_in_place():
    a, b = _model(), _model()
    oa, ob = _keyed(a), _keyed(b)
    _train(a, oa)
    ob.init_state()  # zero-gradient step materialises exp_avg / exp_avg_sq so that there is something to load into
    sd = oa.state_dict()
    assert set(sd["state"].keys()) == {"0.weight", "0.bias", "1.weight", "1.bias"} and "param_groups" not in sd
    exp_avg_before = ob.state[b[0].weight]["exp_avg"]
    ob.load_state_dict(sd)
    assert ob.state[b[0].weight]["exp_avg"] is exp_avg_before  # same storage, new values
    torch.testing.assert_close(exp_avg_before, oa.state[a[0].weight]["exp_avg"])
    assert float(ob.state[b[1].bias]["step"]) == 3.0


def test_param_groups_round_trip_and_mismatch_errors():
    a, b = _model(), _model()
    oa, ob = _keyed(a, lr=0.5), _keyed(b, lr=0.1)
    for o in (oa, ob):
        o.save_param_groups(True)
        o.init_state()
    ob.load_state_dict(oa.state_dict())
    assert ob.param_groups[0]["lr"] == 0.5
    sd = oa.state_dict()
    sd["state"].pop("1.bias")
    with pytest.raises(StateMismatch, match="parameter count"):
        ob.load_state_dict(sd)
    sd = oa.state_dict()
    sd["state"]["renamed"] = sd["state"].pop("1.bias")
    with pytest.raises(StateMismatch, match="1.bias not found"):
        ob.load_state_dict(sd)
    sd = oa.state_dict()
    sd["param_groups"][0]["params"] = ["0.weight"]
    with pytest.raises(StateMismatch, match="Group .* not found"):
        ob.load_state_dict(sd)
    with pytest.raises(ValueError, match="must be params"):
        KeyedOptimizer({"w": a[0].weight}, {a[1].weight: {}}, [{"params": [a[0].weight]}])


def test_state_tree_handlers():
    class Counter:
        def __init__(self, n):
            self.n = n

        def state_dict(self):
            return {"n": self.n}

        def load_state_dict(self, sd):
            self.n = sd["n"]

    cur = {"t": torch.zeros(3), "nested": {"c": Counter(1), "plain": 5}}
    new = {"t": torch.arange(3.0), "nested": {"c": {"n": 9}, "plain": [1, 2]}}
    t = cur["t"]
    assert export_state(cur)["nested"]["c"] == {"n": 1}
    import_state(cur, new, [])
    assert cur["t"] is t and t.tolist() == [0.0, 1.0, 2.0] and cur["nested"]["c"].n == 9 and cur["nested"]["plain"] == [1, 2]
    with pytest.raises(StateMismatch, match="nested/plain"):
        import_state(cur, {"t": t, "nested": {"c": {"n": 1}}}, [])


def test_combined_optimizer_prefixes_and_uniqueness():
    a, b = _model(), _model()
    combo = CombinedOptimizer([("dense", _keyed(a)), ("over", _keyed(b))])
    assert "dense.0.weight" in combo.params and "over.1.bias" in combo.params
    _train(a, combo)
    sd = combo.state_dict()
    assert "dense.0.weight" in sd["state"] and "over.0.weight" not in sd["state"]  # `b` never stepped: stateless
    with pytest.raises(ValueError, match="Duplicate param key"):
        CombinedOptimizer([_keyed(a), _keyed(a)])
    assert CombinedOptimizer.prepend_opt_key("w", "") == "w" and CombinedOptimizer.prepend_opt_key("", "k") == "k"


def test_optimizer_wrapper_aliases_wrapped_state():
    a = _model()
    inner = _keyed(a)
    w = OptimizerWrapper(inner)
    _train(a, w)
    assert w.state is inner.state and w.state_dict()["state"].keys() == inner.state_dict()["state"].keys()


def test_schedule_shapes():
    assert lr_multiplier(WarmupStage(WarmupPolicy.LINEAR, max_iters=10, value=0.0), 5) == pytest.approx(0.5)
    assert lr_multiplier(WarmupStage(WarmupPolicy.CONSTANT, max_iters=10, value=0.3, lr_scale=2.0), 7) == pytest.approx(0.6)
    assert lr_multiplier(WarmupStage(WarmupPolicy.POLY, max_iters=10, value=2.0, decay_iters=10), 5) == pytest.approx(0.25)
    assert lr_multiplier(WarmupStage(WarmupPolicy.STEP, max_iters=100, value=0.5, decay_iters=10), 25) == pytest.approx(0.25)
    assert lr_multiplier(WarmupStage(WarmupPolicy.INVSQRT, max_iters=100), 16) == pytest.approx(0.25)
    cos = WarmupStage(WarmupPolicy.COSINE_ANNEALING_WARM_RESTARTS, max_iters=100, value=0.1, sgdr_period=10)
    assert lr_multiplier(cos, 0) == pytest.approx(1.0) and lr_multiplier(cos, 5) == pytest.approx(0.1 + 0.9 * 0.5) and lr_multiplier(cos, 10) == pytest.approx(1.0)
    itp = WarmupStage(WarmupPolicy.INTERPOLATE, max_iters=30, value=1.0, start_interpolating_iters=10, end_value=0.2)
    assert lr_multiplier(itp, 10) == pytest.approx(1.0) and lr_multiplier(itp, 20) == pytest.approx(0.6) and lr_multiplier(itp, 30) == pytest.approx(0.2)


def test_warmup_optimizer_stages_and_resume():
    a = _model()
    stages = [WarmupStage(WarmupPolicy.LINEAR, max_iters=4, value=0.0), WarmupStage(WarmupPolicy.INTERPOLATE, max_iters=8, value=1.0, start_interpolating_iters=4, end_value=0.5)]
    opt = WarmupOptimizer(_keyed(a), stages, lr=1.0)
    lrs = []
    for _ in range(10):
        lrs.append(opt.param_groups[0]["lr"])
        _train(a, opt, 1)
    assert lrs[:5] == pytest.approx([0.0, 0.25, 0.5, 0.75, 1.0])
    assert lrs[5:9] == pytest.approx([0.875, 0.75, 0.625, 0.5])
    assert lrs[9] == pytest.approx(1.0)  # past the last stage
    sd = opt.state_dict()
    assert sd["state"]["__warmup"]["warmup"].tolist() == [10, 2]
    b = _model()
    opt_b = WarmupOptimizer(_keyed(b), [WarmupStage(WarmupPolicy.LINEAR, max_iters=4, value=0.0),
                                        WarmupStage(WarmupPolicy.INTERPOLATE, max_iters=8, value=1.0, start_interpolating_iters=4, end_value=0.5)], lr=1.0)
    _train(b, opt_b, 6)
    assert opt_b.param_groups[0]["lr"] == pytest.approx(0.75)
    with pytest.raises(AssertionError):
        WarmupOptimizer(_keyed(_model()), [WarmupStage(max_iters=5), WarmupStage(max_iters=5)])
    with pytest.raises(AssertionError):
        WarmupOptimizer(_keyed(_model()), [WarmupStage(WarmupPolicy.INTERPOLATE, max_iters=5)])


def _clip_rank(ctx):
    """Global-norm clipping where only rank 0 owns a sharded parameter (and rank 1 has no gradient for its replicated one on the second
    step): every rank still takes part in the sharded all-reduce, norms agree with a single-process computation."""
    import torch

    from torchrec_b200.optim.clipping import GradientClipping, GradientClippingOptimizer
    from torchrec_b200.optim.keyed import KeyedOptimizerWrapper

    me = ctx.rank
    torch.manual_seed(0)
    rep = torch.nn.Parameter(torch.ones(4))                      # replicated: same values on every rank, counted once
    params = {"rep": rep}
    if me == 0:
        sh = torch.nn.Parameter(torch.full((3,), 2.0))           # "sharded": only rank 0 holds a shard
        sh._is_sharded = True
        params["sh"] = sh
    opt = GradientClippingOptimizer(KeyedOptimizerWrapper(params, lambda p: torch.optim.SGD(p, lr=1.0)), clipping=GradientClipping.NORM, max_gradient=1.0,
                                    enable_global_grad_clip=True)
    assert opt._has_sharded_anywhere  # rank 1 learnt it from rank 0
    rep.grad = torch.full((4,), 3.0)
    if me == 0:
        sh.grad = torch.full((3,), 4.0)
    total = opt.clip_grad_norm_()
    want = (4 * 9.0 + 3 * 16.0) ** 0.5                            # ||rep||^2 + ||sh||^2, the replicated part counted once
    assert abs(float(total) - want) < 1e-4, (me, float(total), want)
    torch.testing.assert_close(rep.grad, torch.full((4,), 3.0 / want), rtol=1e-4, atol=1e-5)
    if me == 0:
        torch.testing.assert_close(sh.grad, torch.full((3,), 4.0 / want), rtol=1e-4, atol=1e-5)
    # second step: nobody has a sharded gradient, rank 1 has no gradient at all - the collective still lines up (no deadlock)
    rep.grad = torch.full((4,), 0.1) if me == 0 else None
    if me == 0:
        sh.grad = None
    total2 = opt.clip_grad_norm_()
    if me == 0:
        assert abs(float(total2) - (4 * 0.01) ** 0.5) < 1e-5


def test_global_grad_norm_clipping_is_collective_safe():
    from torchrec_b200.utils.multiprocess import run_multi_process

    run_multi_process(_clip_rank, world_size=2, backend="gloo")


def _run_semi_sync(ctx):
    """Two workers run local SGD on different data; every 2nd step the outer SGD (lr 1) applies the averaged pseudo-gradient: both workers
    land on the average of where they went, and the global copy follows."""
    from torchrec_b200.optim.keyed import KeyedOptimizerWrapper
    from torchrec_b200.optim.semi_sync import SemisyncOptimizer

    torch.manual_seed(0)
    model = torch.nn.Linear(3, 1, bias=False)
    with torch.no_grad():
        model.weight.fill_(1.0)
    named = dict(model.named_parameters())
    local = KeyedOptimizerWrapper(named, lambda ps: torch.optim.SGD(ps, lr=0.1))
    outer = KeyedOptimizerWrapper(named, lambda ps: torch.optim.SGD(ps, lr=1.0))
    opt = SemisyncOptimizer(model.parameters(), local, outer, num_local_steps=2)
    x = torch.full((1, 3), float(ctx.rank + 1))
    trace = []
    for step in range(4):
        opt.zero_grad()
        model(x).sum().backward()
        opt.step()
        trace.append(model.weight.detach().clone())
    # local steps: w -= 0.1 * (rank + 1); after two of them the workers sit at 1 - 0.2 * (rank + 1); the average over ranks 0, 1 is 1 - 0.3
    torch.testing.assert_close(trace[0], torch.full((1, 3), 1.0 - 0.1 * (ctx.rank + 1)))
    torch.testing.assert_close(trace[1], torch.full((1, 3), 0.7))
    torch.testing.assert_close(trace[2], torch.full((1, 3), 0.7 - 0.1 * (ctx.rank + 1)))
    torch.testing.assert_close(trace[3], torch.full((1, 3), 0.4))
    assert int(opt._global_step_counter) == 2 and int(opt._local_step_counter) == 4
    sd = opt.state_dict()
    opt.load_state_dict(sd)
    assert int(opt._local_step_counter) == 4


def test_semi_sync_optimizer_averages_workers_every_n_steps():
    from torchrec_b200.utils.multiprocess import run_multi_process

    run_multi_process(_run_semi_sync, world_size=2, backend="gloo")
