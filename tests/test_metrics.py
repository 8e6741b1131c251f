"""Each metric vs a straightforward reference formula (strategy of the reference's metrics/tests)."""
import math

import pytest
import torch

from torchrec_b200.metrics import (
    DefaultMetricsConfig,
    MetricsConfig,
    RecComputeMode,
    RecMetricDef,
    RecMetricEnum,
    RecMetricModule,
    RecTaskInfo,
    ThroughputDef,
    generate_metric_module,
)
from torchrec_b200.metrics import metrics_impl as M
from torchrec_b200.utils.multiprocess import run_multi_process


def _data(n=200, seed=0):
    g = torch.Generator().manual_seed(seed)
    p = torch.rand(n, generator=g)
    l = (torch.rand(n, generator=g) < p).float()
    w = torch.rand(n, generator=g) + 0.1
    return p, l, w


def _one(metric_cls, task="t", **kw):
    return metric_cls(world_size=1, my_rank=0, batch_size=64, tasks=[RecTaskInfo(name=task)], window_size=1000, **kw)


def test_ne_calibration_ctr_mse_mae_accuracy():
    p, l, w = _data()
    ne = _one(M.NEMetric, include_logloss=True)
    cal, ctr, mse, mae, acc = _one(M.CalibrationMetric), _one(M.CTRMetric), _one(M.MSEMetric), _one(M.MAEMetric), _one(M.AccuracyMetric)
    for m in (ne, cal, ctr, mse, mae, acc):
        for i in range(0, 200, 50):
            m.update(predictions={"t": p[i : i + 50]}, labels={"t": l[i : i + 50]}, weights={"t": w[i : i + 50]})
    ce = -(w * (l * torch.log2(p.clamp(1e-12)) + (1 - l) * torch.log2((1 - p).clamp(1e-12)))).sum()
    ml = (w * l).sum() / w.sum()
    base = -((w * l).sum() * math.log2(ml) + (w * (1 - l)).sum() * math.log2(1 - ml))
    r = ne.compute()
    assert r["ne-t|lifetime_ne"].item() == pytest.approx((ce / base).item(), rel=1e-4)
    assert r["ne-t|window_ne"].item() == pytest.approx((ce / base).item(), rel=1e-4)
    assert "ne-t|lifetime_logloss" in r
    assert cal.compute()["calibration-t|lifetime_calibration"].item() == pytest.approx(((w * p).sum() / (w * l).sum()).item(), rel=1e-5)
    assert ctr.compute()["ctr-t|lifetime_ctr"].item() == pytest.approx(ml.item(), rel=1e-5)
    assert mse.compute()["mse-t|lifetime_mse"].item() == pytest.approx(((w * (p - l) ** 2).sum() / w.sum()).item(), rel=1e-5)
    assert mae.compute()["mae-t|lifetime_mae"].item() == pytest.approx(((w * (p - l).abs()).sum() / w.sum()).item(), rel=1e-5)
    assert acc.compute()["accuracy-t|lifetime_accuracy"].item() == pytest.approx(((w * ((p >= 0.5).float() == l)).sum() / w.sum()).item(), rel=1e-5)


def test_auc_matches_pairwise_definition():
    p, l, w = _data(150, seed=3)
    auc = _one(M.AUCMetric)
    auc.update(predictions={"t": p}, labels={"t": l}, weights={"t": w})
    pos, neg = l == 1, l == 0
    num = ((p[pos].unsqueeze(1) > p[neg].unsqueeze(0)).float() + 0.5 * (p[pos].unsqueeze(1) == p[neg].unsqueeze(0)).float())
    ref = (num * w[pos].unsqueeze(1) * w[neg].unsqueeze(0)).sum() / (w[pos].sum() * w[neg].sum())
    assert auc.compute()["auc-t|window_auc"].item() == pytest.approx(ref.item(), rel=1e-5)


def test_window_drops_old_batches():
    m = M.CTRMetric(world_size=1, my_rank=0, batch_size=10, tasks=[RecTaskInfo(name="t")], window_size=20)
    m.update(predictions={"t": torch.zeros(10)}, labels={"t": torch.ones(10)}, weights=None)
    for _ in range(2):
        m.update(predictions={"t": torch.zeros(10)}, labels={"t": torch.zeros(10)}, weights=None)
    r = m.compute()
    assert r["ctr-t|lifetime_ctr"].item() == pytest.approx(1 / 3, rel=1e-5)
    assert r["ctr-t|window_ctr"].item() == pytest.approx(0.0, abs=1e-9)


def test_fused_tasks_and_metric_module():
    cfg = MetricsConfig(rec_tasks=[RecTaskInfo(name="a", label_name="la", prediction_name="pa", weight_name="wa"),
                                   RecTaskInfo(name="b", label_name="lb", prediction_name="pb", weight_name="wb")],
                        rec_metrics={RecMetricEnum.NE: RecMetricDef(rec_task_indices=[0, 1], window_size=1000),
                                     RecMetricEnum.AUC: RecMetricDef(rec_task_indices=[0], window_size=1000)},
                        throughput_metric=ThroughputDef(), rec_compute_mode=RecComputeMode.FUSED_TASKS_COMPUTATION, compute_interval_steps=1)
    mod = generate_metric_module(RecMetricModule, cfg, batch_size=50, world_size=1, my_rank=0, state_metrics_mapping={}, device=torch.device("cpu"))
    p, l, w = _data(50)
    out = {"la": l, "pa": p, "wa": w, "lb": 1 - l, "pb": 1 - p, "wb": w}
    mod.update(out)
    r = mod.compute()
    assert r["ne-a|lifetime_ne"].item() == pytest.approx(r["ne-b|lifetime_ne"].item(), rel=1e-6)
    assert "auc-a|window_auc" in r and "throughput-throughput|total_examples" in r


def _dist_metric(ctx):
    p, l, w = _data(100, seed=ctx.rank)
    m = M.NEMetric(world_size=ctx.world_size, my_rank=ctx.rank, batch_size=100, tasks=[RecTaskInfo(name="t")], window_size=1000, compute_on_all_ranks=True)
    m.update(predictions={"t": p}, labels={"t": l}, weights={"t": w})
    got = m.compute()["ne-t|lifetime_ne"].item()
    ps, ls, ws = zip(*[_data(100, seed=r) for r in range(ctx.world_size)])
    p, l, w = torch.cat(ps), torch.cat(ls), torch.cat(ws)
    ce = -(w * (l * torch.log2(p.clamp(1e-12)) + (1 - l) * torch.log2((1 - p).clamp(1e-12)))).sum()
    ml = (w * l).sum() / w.sum()
    base = -((w * l).sum() * math.log2(ml) + (w * (1 - l)).sum() * math.log2(1 - ml))
    assert got == pytest.approx((ce / base).item(), rel=1e-4)
    auc = M.AUCMetric(world_size=ctx.world_size, my_rank=ctx.rank, batch_size=100, tasks=[RecTaskInfo(name="t")], window_size=1000, compute_on_all_ranks=True)
    auc.update(predictions={"t": ps[ctx.rank]}, labels={"t": ls[ctx.rank]}, weights={"t": ws[ctx.rank]})
    assert auc.compute()["auc-t|window_auc"].item() == pytest.approx(M._auc_from_samples(p.double(), l.double(), w.double()).item(), rel=1e-6)
    # states of the other shapes: mean over ranks (scalar), session sums (NDCG), per-threshold sums (hindsight PR), grouped window keys (grouped AUC)
    kw = dict(world_size=ctx.world_size, my_rank=ctx.rank, batch_size=100, tasks=[RecTaskInfo(name="t")], window_size=1000, compute_on_all_ranks=True)
    sc = M.ScalarMetric(**kw)
    sc.update(predictions={"t": ps[ctx.rank]}, labels={"t": torch.full((100,), float(ctx.rank + 1))}, weights={"t": ws[ctx.rank]})
    assert sc.compute()["scalar-t|lifetime_scalar"].item() == pytest.approx(sum(range(1, ctx.world_size + 1)) / ctx.world_size)
    sess = torch.arange(100) // 5
    nd, single = M.NDCGMetric(**kw), []
    nd.update(predictions={"t": ps[ctx.rank]}, labels={"t": ls[ctx.rank]}, weights={"t": ws[ctx.rank]}, required_inputs={"session_id": sess})
    for r in range(ctx.world_size):
        one = M.NDCGMetric(world_size=1, my_rank=0, batch_size=100, tasks=[RecTaskInfo(name="t")], window_size=1000)
        one.update(predictions={"t": ps[r]}, labels={"t": ls[r]}, weights={"t": ws[r]}, required_inputs={"session_id": sess})
        single.append(one.compute()["ndcg-t|lifetime_ndcg"].item())
    assert nd.compute()["ndcg-t|lifetime_ndcg"].item() == pytest.approx(sum(single) / len(single), rel=1e-9)  # equal session counts per rank
    hp, whole = M.HindsightTargetPRMetric(**kw), M.HindsightTargetPRMetric(world_size=1, my_rank=0, batch_size=200, tasks=[RecTaskInfo(name="t")], window_size=1000)
    hp.update(predictions={"t": ps[ctx.rank]}, labels={"t": ls[ctx.rank]}, weights={"t": ws[ctx.rank]})
    whole.update(predictions={"t": p}, labels={"t": l}, weights={"t": w})
    for k, v in whole.compute().items():
        assert hp.compute()[k].item() == pytest.approx(v.item(), rel=1e-9), k
    ga = M.AUCMetric(grouped_auc=True, **kw)
    keys = [torch.arange(100) % 3 + 10 * r for r in range(ctx.world_size)]  # disjoint groups per rank
    ga.update(predictions={"t": ps[ctx.rank]}, labels={"t": ls[ctx.rank]}, weights={"t": ws[ctx.rank]}, required_inputs={"grouping_keys": keys[ctx.rank]})
    want = sum(M._auc_from_samples(ps[r][keys[r] == g].double(), ls[r][keys[r] == g].double(), ws[r][keys[r] == g].double()).item()
               for r in range(ctx.world_size) for g in keys[r].unique()) / (3 * ctx.world_size)
    assert ga.compute()["auc-t|window_grouped_auc"].item() == pytest.approx(want, rel=1e-6)


def test_metrics_sync_across_ranks():
    run_multi_process(_dist_metric, world_size=2, backend="gloo")


def test_ndcg_gauc_recall_session_and_others_run():
    from torchrec_b200.metrics.metrics_config import SessionMetricDef

    p, l, w = _data(60, seed=5)
    sess = torch.arange(60) // 6
    session_task = RecTaskInfo(name="t", session_metric_def=SessionMetricDef(session_var_name="session_ids", top_threshold=2))
    for m, extra in ((_one(M.NDCGMetric), {"required_inputs": {"session_id": sess}}), (_one(M.GAUCMetric), {"num_candidates": torch.full((10,), 6)}),
                     (M.RecallSessionMetric(world_size=1, my_rank=0, batch_size=64, tasks=[session_task], window_size=1000), {"required_inputs": {"session_ids": sess}}),
                     (M.PrecisionSessionMetric(world_size=1, my_rank=0, batch_size=64, tasks=[session_task], window_size=1000), {"required_inputs": {"session_ids": sess}})):
        m.update(predictions={"t": p}, labels={"t": l}, weights={"t": w}, **extra)
        v = list(m.compute().values())[0]
        assert 0.0 <= float(v) <= 1.5, type(m)
    for cls in (M.PrecisionMetric, M.RecallMetric, M.WeightedAvgMetric, M.NMSEMetric, M.XAUCMetric, M.RAUCMetric, M.AUPRCMetric, M.CaliFreeNEMetric,
                M.UnweightedNEMetric, M.ServingNEMetric, M.ServingCalibrationMetric, M.AverageMetric, M.HindsightTargetPRMetric, M.ScalarMetric,
                M.TowerQPSMetric):
        m = _one(cls)
        for _ in range(3):
            m.update(predictions={"t": p}, labels={"t": l}, weights={"t": w})
        r = m.compute()
        assert len(r) >= 1 and all(torch.isfinite(torch.as_tensor(v)).all() for v in r.values()), cls
    mc = _one(M.MulticlassRecallMetric, number_of_classes=4)
    mc.update(predictions={"t": torch.rand(60, 4)}, labels={"t": torch.randint(0, 4, (60,))}, weights={"t": w})
    assert list(mc.compute().values())[0].shape[-1] == 4


def test_recalibrated_and_volume_metrics():
    p, l, w = _data()
    c = 0.25
    rne, rcal = _one(M.RecalibratedNEMetric, recalibration_coefficient=c), _one(M.RecalibratedCalibrationMetric, recalibration_coefficient=c)
    ae, npos, sw, nmiss, wsp = (_one(M.ServingAELossMetric), _one(M.NumPositiveSamplesMetric), _one(M.SumWeightsMetric),
                                _one(M.NumMissingLabelsMetric), _one(M.WeightedSumPredictionsMetric))
    l_nan = l.clone()
    l_nan[::7] = float("nan")
    for m in (rne, rcal, ae, npos, sw, wsp):
        m.update(predictions={"t": p}, labels={"t": l}, weights={"t": w})
    nmiss.update(predictions={"t": p}, labels={"t": l_nan}, weights={"t": w})
    q = p.double() / (p.double() + (1 - p.double()) / c)
    plain = _one(M.NEMetric)
    plain.update(predictions={"t": q.float()}, labels={"t": l}, weights={"t": w})
    assert rne.compute()["recalibrated_ne-t|lifetime_recalibrated_ne"].item() == pytest.approx(plain.compute()["ne-t|lifetime_ne"].item(), rel=1e-5)
    assert rcal.compute()["recalibrated_calibration-t|lifetime_recalibrated_calibration"].item() == pytest.approx(
        ((w * q).sum() / (w * l).sum()).item(), rel=1e-5)
    assert ae.compute()["serving_ae_loss-t|lifetime_serving_ae_loss"].item() == pytest.approx(((w * (p - l).abs()).sum() / w.sum()).item(), rel=1e-5)
    assert npos.compute()["num_positive_samples-t|lifetime_num_positive_samples"].item() == pytest.approx((w * l).sum().item(), rel=1e-5)
    assert sw.compute()["sum_weights-t|window_sum_weights"].item() == pytest.approx(w.sum().item(), rel=1e-5)
    assert nmiss.compute()["num_missing_labels-t|lifetime_num_missing_labels"].item() == pytest.approx(w[::7].sum().item(), rel=1e-5)
    assert wsp.compute()["weighted_sum_predictions-t|lifetime_weighted_sum_predictions"].item() == pytest.approx((w * p).sum().item(), rel=1e-5)
    mod = generate_metric_module(RecMetricModule, MetricsConfig(rec_tasks=[RecTaskInfo(name="t")], rec_metrics={
        RecMetricEnum.RECALIBRATED_NE: RecMetricDef(rec_tasks=[RecTaskInfo(name="t")], window_size=1000, arguments={"recalibration_coefficient": c}),
        RecMetricEnum.SUM_WEIGHTS: RecMetricDef(rec_tasks=[RecTaskInfo(name="t")], window_size=1000)}), batch_size=200, world_size=1, my_rank=0,
        state_metrics_mapping={}, device=torch.device("cpu"))
    assert len(mod.rec_metrics.rec_metrics) == 2


def test_metric_extras_deferrable_snapshot_noop_cpu_comms():
    from concurrent.futures import Future

    from torchrec_b200.metrics import (CPUCommsRecMetricModule, DeferrableMetrics, MetricComputeJob, MetricStateSnapshot, MetricUpdateJob, NoOpMetricModule,
                                       SynchronizationMarker)
    from torchrec_b200.metrics.model_utils import is_empty_signals, parse_model_outputs, session_ids_to_tensor

    # deferrable metrics
    fut = Future()
    dm = DeferrableMetrics(fut)
    dm["local"] = 1.0
    seen = []
    dm.subscribe(seen.append)
    assert not dm.is_resolved() and seen == []
    fut.set_result({"ne": 0.5})
    assert seen == [{"ne": 0.5, "local": 1.0}] and dm["ne"] == 0.5 and len(dm) == 2 and dm.resolve() == {"ne": 0.5, "local": 1.0}
    plain = DeferrableMetrics({"a": 1})
    plain.update({"b": 2})
    assert dict(plain) == {"a": 1, "b": 2} and bool(plain) and not bool(DeferrableMetrics())

    # no-op module
    noop = NoOpMetricModule()
    noop.update({"x": torch.ones(1)})
    assert not noop.should_compute() and dict(noop.compute()) == {} and noop.async_compute().result() == {}

    # snapshot -> cpu-comms clone computes the same numbers as the live module
    p, l, w = _data()
    cfg = MetricsConfig(rec_tasks=[RecTaskInfo(name="t")], rec_metrics={RecMetricEnum.NE: RecMetricDef(rec_tasks=[RecTaskInfo(name="t")], window_size=1000),
                                                                         RecMetricEnum.AUC: RecMetricDef(rec_tasks=[RecTaskInfo(name="t")], window_size=1000)})
    live = generate_metric_module(RecMetricModule, cfg, batch_size=200, world_size=1, my_rank=0, state_metrics_mapping={}, device=torch.device("cpu"))
    live.update({"prediction": p, "label": l, "weight": w})
    snap = MetricStateSnapshot.from_metrics(live.rec_metrics)
    assert any(k.startswith("ne_") for k in snap.metric_states)
    shadow = generate_metric_module(CPUCommsRecMetricModule, cfg, batch_size=200, world_size=1, my_rank=0, state_metrics_mapping={}, device=torch.device("cpu"))
    got = shadow.compute_from_snapshot(snap)
    want = live.compute()
    for k in want:
        if "ne" in k or "auc" in k:
            assert float(got[k]) == pytest.approx(float(want[k]), rel=1e-6)
    job = MetricUpdateJob({"x": torch.ones(1)}, {}, merged_count=3)
    f2 = Future()
    assert job.merged_count == 3 and MetricComputeJob(f2, snap).metric_state_snapshot is snap and SynchronizationMarker(f2).future is f2

    # model utils
    assert session_ids_to_tensor(["a", "a", "b", "a"]).tolist() == [0, 0, 1, 2]
    lab, pred, wt = parse_model_outputs("l", "p", "w", {"l": torch.ones(4, 1), "p": torch.rand(4, 1), "w": torch.ones(4, 1)})
    assert lab.shape == pred.shape == wt.shape == (4,)
    assert is_empty_signals(torch.zeros(0), torch.zeros(0), torch.zeros(0))


def test_per_metric_modules_and_functional_helpers():
    """Every reference ``torchrec/metrics/<name>.py`` import path resolves; the stateless helpers agree with the stateful metrics."""
    import importlib

    import torch

    from torchrec_b200.metrics.metrics_config import RecMetricEnum
    from torchrec_b200.metrics.metric_module import REC_METRICS_MAPPING
    from torchrec_b200.metrics.rec_metric import RecTaskInfo

    for mod in ("accuracy", "auc", "auprc", "average", "cali_free_ne", "calibration", "calibration_with_recalibration", "cpu_offloaded_metric_module", "ctr", "gauc",
                "hindsight_target_pr", "mae", "mse", "multi_label_precision", "multiclass_recall", "ndcg", "ne", "ne_positive", "ne_with_recalibration", "nmse",
                "num_missing_labels", "num_positive_samples", "output", "precision", "precision_session", "rauc", "recall", "recall_session", "scalar", "segmented_ne",
                "serving_calibration", "serving_ne", "sum_weights", "tensor_weighted_avg", "tower_qps", "unweighted_ne", "weighted_avg", "weighted_sum_predictions", "xauc"):
        importlib.import_module(f"torchrec_b200.metrics.{mod}")
    assert RecMetricEnum.NE_POSITIVE in REC_METRICS_MAPPING

    from torchrec_b200.metrics.ne import compute_ne, get_ne_states
    from torchrec_b200.metrics.ne_positive import NEPositiveMetric, compute_ne_positive, get_ne_positive_states
    from torchrec_b200.metrics.mse import compute_mse, compute_r_squared, compute_rmse, get_mse_states
    from torchrec_b200.metrics.rauc import compute_rauc, count_reverse_pairs_divide_and_conquer
    from torchrec_b200.metrics.xauc import compute_xauc, get_xauc_states

    g = torch.Generator().manual_seed(3)
    p = torch.rand(1, 64, generator=g)
    y = (torch.rand(1, 64, generator=g) > 0.6).float()
    w = torch.rand(1, 64, generator=g)
    task = RecTaskInfo(name="t", label_name="l", prediction_name="p", weight_name="w")
    m = NEPositiveMetric(world_size=1, my_rank=0, batch_size=64, tasks=[task])
    m.update(predictions={"t": p[0]}, labels={"t": y[0]}, weights={"t": w[0]})
    got = m.compute()["ne_positive-t|lifetime_ne_positive"]
    st = get_ne_positive_states(y, p, w, 1e-12)
    want = compute_ne_positive(st["cross_entropy_positive_sum"], st["weighted_num_samples"], st["pos_labels"], st["neg_labels"], 1e-12)
    torch.testing.assert_close(got.double().reshape(-1), want.reshape(-1))
    ne = get_ne_states(y, p, w, 1e-12)
    assert float(compute_ne(ne["cross_entropy_sum"], ne["weighted_num_samples"], ne["pos_labels"], ne["neg_labels"])) > 0
    ms = get_mse_states(y, p, w)
    torch.testing.assert_close(compute_rmse(ms["error_sum"], ms["weighted_num_samples"]) ** 2, compute_mse(ms["error_sum"], ms["weighted_num_samples"]))
    assert float(compute_r_squared(ms["error_sum"], ms["weighted_num_samples"], ms["label_sum"], ms["label_squared_sum"])) <= 1.0
    assert count_reverse_pairs_divide_and_conquer([4, 3, 2, 1]) == 1.0 and count_reverse_pairs_divide_and_conquer([1, 3, 2, 4]) == 1 / 6
    torch.testing.assert_close(compute_rauc(1, torch.tensor([[0.1, 0.2, 0.3]]), torch.tensor([[1.0, 2.0, 3.0]]), torch.ones(1, 3)), torch.ones(1, dtype=torch.double))
    xs = get_xauc_states(torch.tensor([[1.0, 2.0, 3.0]]), torch.tensor([[0.1, 0.3, 0.2]]), torch.ones(1, 3))
    torch.testing.assert_close(compute_xauc(xs["error_sum"], xs["weighted_num_pairs"]), torch.tensor([2.0 / 3.0], dtype=torch.double))


def test_mock_rec_metric_records_calls_and_exposes_states():
    from torchrec_b200.metrics.rec_metric import RecComputeMode
    from torchrec_b200.metrics.test_utils import (MockRecMetric, assert_tensor_dict_equals, create_metric_states_dict, create_tensor_list_states, create_tensor_states)

    tasks = [RecTaskInfo(name="a"), RecTaskInfo(name="b")]
    init = create_tensor_states(["s1", "s2"])
    m = MockRecMetric(world_size=1, my_rank=0, batch_size=4, tasks=tasks, initial_states={k: v.clone() for k, v in init.items()})
    assert not m.update_called() and not m.compute_called() and m.verify_sync_disabled()
    p = {"a": torch.rand(4), "b": torch.rand(4)}
    m.update(predictions=p, labels=p, weights=None)
    m.update(predictions=p, labels=p, weights=p)
    assert m.update_called_count == 2 and m.predictions_update_calls[0] is p and m.weights_update_calls == [None, p] and m.compute() == {} and m.compute_called()
    assert_tensor_dict_equals(m.get_computation_states(), init)
    m.add_to_computation_states({"s1": torch.ones(1)})
    assert_tensor_dict_equals(m.get_computation_states(), {"s1": init["s1"] + 1, "s2": init["s2"]})
    m.set_computation_states({"s2": torch.zeros(1)})
    assert float(m.get_computation_states()["s2"]) == 0.0 and len(m._metrics_computations) == 2
    m.reset()
    assert not m.update_called() and not m.compute_called()
    with pytest.raises(AssertionError):
        assert_tensor_dict_equals(m.get_computation_states(), init)
    lists = create_tensor_list_states(["preds", "labels"])
    fused = MockRecMetric(world_size=1, my_rank=0, batch_size=4, tasks=tasks, compute_mode=RecComputeMode.FUSED_TASKS_COMPUTATION, initial_states=lists, is_tensor_list=True,
                          reduction_fn="cat")
    assert len(fused._metrics_computations) == 1
    extra = torch.rand(1, 3)
    fused.append_to_computation_states({"preds": extra})
    assert_tensor_dict_equals(fused.get_computation_states(), {"preds": lists["preds"] + [extra], "labels": lists["labels"]})
    assert create_metric_states_dict("a", "MockRecMetricComputation", {"s1": 1}) == {"a_MockRecMetricComputation_s1": 1}
