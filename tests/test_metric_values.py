"""Every metric against an independent plain-python evaluation of its definition (the reference's metric semantics: report names,
additive lifetime / window states, per-batch sessions), on random data fed in several batches."""
import math
import random

import pytest
import torch

from torchrec_b200.metrics import metrics_impl as M
from torchrec_b200.metrics.metrics_config import SessionMetricDef
from torchrec_b200.metrics.rec_metric import RecComputeMode, RecMetricException, RecTaskInfo

N_BATCH, B = 3, 40


def _batches(seed, binary=True):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(N_BATCH):
        p = torch.rand(B, generator=g)
        l = (torch.rand(B, generator=g) < p).float() if binary else torch.rand(B, generator=g) * 3
        w = torch.rand(B, generator=g) + 0.1
        out.append((p, l, w))
    return out


def _metric(cls, task=None, **kw):
    return cls(world_size=1, my_rank=0, batch_size=B, tasks=[task or RecTaskInfo(name="t")], window_size=10_000, **kw)


def _feed(m, batches, **extra_per_batch):
    for i, (p, l, w) in enumerate(batches):
        kw = {k: v[i] for k, v in extra_per_batch.items()}
        m.update(predictions={"t": p}, labels={"t": l}, weights={"t": w}, **kw)
    return {k: v for k, v in m.compute().items()}


def _cat(batches):
    return [torch.cat([b[i] for b in batches]).double().tolist() for i in range(3)]


def _close(got, want, tol=1e-6):
    assert float(got) == pytest.approx(want, rel=tol, abs=tol), (float(got), want)


def test_sum_state_metrics_match_definitions():
    bs = _batches(0)
    p, l, w = _cat(bs)
    W = sum(w)
    r = _feed(_metric(M.AverageMetric), bs)
    for scope in ("lifetime", "window"):
        _close(r[f"average-t|{scope}_label_average"], sum(a * b for a, b in zip(l, w)) / W)
        _close(r[f"average-t|{scope}_prediction_average"], sum(a * b for a, b in zip(p, w)) / W)
    r = _feed(_metric(M.MSEMetric, include_r_squared=True), bs)
    mse = sum(wi * (pi - li) ** 2 for pi, li, wi in zip(p, l, w)) / W
    mean_l = sum(a * b for a, b in zip(l, w)) / W
    ss_tot = sum(wi * (li - mean_l) ** 2 for li, wi in zip(l, w))
    _close(r["mse-t|lifetime_mse"], mse)
    _close(r["mse-t|window_rmse"], math.sqrt(mse))
    _close(r["mse-t|lifetime_r_squared"], 1 - mse * W / ss_tot)
    assert "mse-t|lifetime_r_squared" not in _feed(_metric(M.MSEMetric), bs)
    r = _feed(_metric(M.NMSEMetric), bs)
    base = sum(wi * (1 - li) ** 2 for li, wi in zip(l, w)) / W
    _close(r["nmse-t|lifetime_nmse"], mse / base)
    _close(r["nmse-t|lifetime_nrmse"], math.sqrt(mse / base))
    # precision / recall / accuracy at the 0.5 threshold
    tp = sum(wi for pi, li, wi in zip(p, l, w) if pi >= 0.5 and li == 1)
    fp = sum(wi for pi, li, wi in zip(p, l, w) if pi >= 0.5 and li == 0)
    fn = sum(wi for pi, li, wi in zip(p, l, w) if pi < 0.5 and li == 1)
    tn = W - tp - fp - fn
    _close(_feed(_metric(M.PrecisionMetric), bs)["precision-t|lifetime_precision"], tp / (tp + fp))
    _close(_feed(_metric(M.RecallMetric), bs)["recall-t|lifetime_recall"], tp / (tp + fn))
    _close(_feed(_metric(M.AccuracyMetric), bs)["accuracy-t|lifetime_accuracy"], (tp + tn) / W)
    # NE family
    ce = sum(-wi * (li * math.log2(pi) + (1 - li) * math.log2(1 - pi)) for pi, li, wi in zip(p, l, w))
    pos = sum(a * b for a, b in zip(l, w))
    base_ce = -(pos * math.log2(pos / W) + (W - pos) * math.log2(1 - pos / W))
    r = _feed(_metric(M.NEMetric, include_logloss=True), bs)
    _close(r["ne-t|lifetime_ne"], ce / base_ce, 1e-5)
    _close(r["ne-t|lifetime_logloss"], ce / W * math.log(2), 1e-5)
    mean_p = sum(a * b for a, b in zip(p, w)) / W
    _close(_feed(_metric(M.CaliFreeNEMetric), bs)["cali_free_ne-t|lifetime_cali_free_ne"],
           (ce / base_ce) / (-pos * math.log2(mean_p) - (W - pos) * math.log2(1 - mean_p)), 1e-5)
    n = len(p)
    ce_u = sum(-(li * math.log2(pi) + (1 - li) * math.log2(1 - pi)) for pi, li in zip(p, l))
    pos_u = sum(l)
    _close(_feed(_metric(M.UnweightedNEMetric), bs)["unweighted_ne-t|lifetime_unweighted_ne"],
           ce_u / -(pos_u * math.log2(pos_u / n) + (n - pos_u) * math.log2(1 - pos_u / n)), 1e-5)
    # serving metrics: NE / calibration keys + the number of examples with a non-zero weight
    bz = [(bp, bl, torch.where(torch.arange(B) % 4 == 0, torch.zeros(B), bw)) for bp, bl, bw in bs]
    r = _feed(_metric(M.ServingNEMetric), bz)
    assert int(r["serving_ne-t|total_examples"]) == N_BATCH * (B - B // 4)
    pz, lz, wz = _cat(bz)
    cez = sum(-wi * (li * math.log2(pi) + (1 - li) * math.log2(1 - pi)) for pi, li, wi in zip(pz, lz, wz))
    posz, Wz = sum(a * b for a, b in zip(lz, wz)), sum(wz)
    _close(r["serving_ne-t|lifetime_ne"], cez / -(posz * math.log2(posz / Wz) + (Wz - posz) * math.log2(1 - posz / Wz)), 1e-5)
    assert "serving_ne-t|window_ne" in r
    r = _feed(_metric(M.ServingCalibrationMetric), bz)
    _close(r["serving_calibration-t|lifetime_calibration"], sum(a * b for a, b in zip(pz, wz)) / posz, 1e-5)
    assert int(r["serving_calibration-t|total_examples"]) == N_BATCH * (B - B // 4)


def test_xauc_rauc_pairs():
    bs = _batches(1, binary=False)
    r = _feed(_metric(M.XAUCMetric), bs)
    num = den = 0.0
    for p, l, w in bs:  # pairs inside a batch only
        p, l, w = p.double().tolist(), l.double().tolist(), w.double().tolist()
        for i in range(B):
            for j in range(i + 1, B):
                den += w[i] * w[j]
                if (p[i] - p[j]) * (l[i] - l[j]) > 0 or (p[i] == p[j] and l[i] == l[j]):
                    num += w[i] * w[j]
    _close(r["xauc-t|lifetime_xauc"], num / den)
    _close(r["xauc-t|window_xauc"], num / den)
    # RAUC over the whole window: share of the pairs ordered the right way; grouped: mean over the groups
    g = torch.Generator().manual_seed(5)
    keys = [torch.randint(0, 3, (B,), generator=g) for _ in range(N_BATCH)]
    m = _metric(M.RAUCMetric, grouped_rauc=True)
    assert m.get_required_inputs() == {"grouping_keys"}
    r = _feed(m, bs, required_inputs=[{"grouping_keys": k} for k in keys])
    p, l, _ = _cat(bs)
    k = torch.cat(keys).tolist()

    def rauc(idx):
        good = tot = 0
        for a in range(len(idx)):
            for b in range(a + 1, len(idx)):
                i, j = idx[a], idx[b]
                tot += 1
                good += (p[i] - p[j]) * (l[i] - l[j]) > 0 or p[i] == p[j] or l[i] == l[j]
        return good / tot

    _close(r["rauc-t|window_rauc"], rauc(list(range(len(p)))))
    _close(r["rauc-t|window_grouped_rauc"], sum(rauc([i for i in range(len(p)) if k[i] == grp]) for grp in range(3)) / 3)
    with pytest.raises(RecMetricException):
        _metric(M.RAUCMetric, grouped_rauc=True, fused_update_limit=2)


def _auc(p, l, w):
    num = den = 0.0
    for i in range(len(p)):
        for j in range(len(p)):
            if l[i] == 1 and l[j] == 0:
                den += w[i] * w[j]
                num += w[i] * w[j] * (1.0 if p[i] > p[j] else 0.5 if p[i] == p[j] else 0.0)
    return num / den if den else 0.5


def test_auc_grouped_auc_auprc_and_gauc():
    bs = _batches(2)
    g = torch.Generator().manual_seed(6)
    keys = [torch.randint(0, 4, (B,), generator=g) for _ in range(N_BATCH)]
    req = [{"grouping_keys": k} for k in keys]
    m = _metric(M.AUCMetric, grouped_auc=True)
    assert m.get_required_inputs() == {"grouping_keys"}
    r = _feed(m, bs, required_inputs=req)
    p, l, w = _cat(bs)
    k = torch.cat(keys).tolist()
    _close(r["auc-t|window_auc"], _auc(p, l, w))
    per_group = []
    for grp in range(4):
        idx = [i for i in range(len(p)) if k[i] == grp]
        per_group.append(_auc([p[i] for i in idx], [l[i] for i in idx], [w[i] for i in idx]))
    _close(r["auc-t|window_grouped_auc"], sum(per_group) / 4)
    assert "auc-t|window_grouped_auc" not in _feed(_metric(M.AUCMetric), bs)
    with pytest.raises(RecMetricException):
        _feed(_metric(M.AUCMetric, grouped_auc=True), bs)  # the grouping keys are required
    # soft labels binarised at 0.039
    soft = [(bp, bl * 0.5 + 0.01, bw) for bp, bl, bw in bs]
    _close(_feed(_metric(M.AUCMetric, apply_bin=True), soft)["auc-t|window_auc"], _auc(p, l, w))
    # AUPRC: sum over the distinct score thresholds of precision * recall increment
    r = _feed(_metric(M.AUPRCMetric, grouped_auprc=True), bs, required_inputs=req)

    def auprc(idx):
        order = sorted(idx, key=lambda i: -p[i])
        tot_pos = sum(w[i] * l[i] for i in idx)
        if tot_pos == 0:
            return 0.0
        tp = fp = area = prev_rec = 0.0
        for n_, i in enumerate(order):
            tp += w[i] * l[i]
            fp += w[i] * (1 - l[i])
            if n_ + 1 == len(order) or p[order[n_ + 1]] != p[i]:
                area += (tp / tot_pos - prev_rec) * tp / (tp + fp)
                prev_rec = tp / tot_pos
        return area

    _close(r["auprc-t|window_auprc"], auprc(list(range(len(p)))), 1e-5)
    _close(r["auprc-t|window_grouped_auprc"], sum(auprc([i for i in range(len(p)) if k[i] == grp]) for grp in range(4)) / 4, 1e-5)
    # GAUC: sessions laid out back to back, `num_candidates` examples each; single-class sessions and constant predictions do not count
    rng = random.Random(3)
    m = _metric(M.GAUCMetric)
    auc_sum = n_eff = 0
    for bp, bl, bw in bs:
        lens = []
        while sum(lens) < B:
            lens.append(min(rng.randint(1, 9), B - sum(lens)))
        bl = bl.clone()
        bl[: lens[0]] = 1.0  # a single-class session
        bp = bp.clone()
        bp[lens[0] : lens[0] + lens[1]] = 0.25  # a session with identical predictions
        m.update(predictions={"t": bp}, labels={"t": bl}, weights={"t": bw}, num_candidates=torch.tensor(lens))
        o = 0
        for n_ in lens:
            sp, sl, sw = bp[o : o + n_].tolist(), bl[o : o + n_].tolist(), bw[o : o + n_].tolist()
            o += n_
            if min(sl) != max(sl) and min(sp) != max(sp):
                auc_sum += _auc(sp, sl, sw)
                n_eff += 1
    r = m.compute()
    _close(r["gauc-t|lifetime_gauc"], auc_sum / n_eff, 1e-5)
    _close(r["gauc-t|window_gauc"], auc_sum / n_eff, 1e-5)
    assert float(r["gauc-t|lifetime_gauc_num_samples"]) == n_eff
    empty = _metric(M.GAUCMetric)
    empty.update(predictions={"t": torch.rand(4)}, labels={"t": torch.ones(4)}, weights={"t": torch.ones(4)}, num_candidates=torch.tensor([4]))
    _close(empty.compute()["gauc-t|lifetime_gauc"], 0.5)


def _ndcg(p, l, k=None, exp=False):
    order = sorted(range(len(p)), key=lambda i: -p[i])
    gain = (lambda x: 2 ** x - 1) if exp else (lambda x: x)
    k = len(p) if k is None else min(k, len(p))
    dcg = sum(gain(l[i]) / math.log2(r + 2) for r, i in enumerate(order[:k]))
    idcg = sum(gain(x) / math.log2(r + 2) for r, x in enumerate(sorted(l, reverse=True)[:k]))
    return dcg / (idcg if idcg else 1e-6)


def test_ndcg_options():
    g = torch.Generator().manual_seed(8)
    bs = []
    for _ in range(N_BATCH):
        p = torch.rand(B, generator=g)
        l = torch.randint(0, 4, (B,), generator=g).float()
        w = torch.rand(B, generator=g) + 0.5
        bs.append((p, l, w))
    sess = [torch.randint(0, 7, (B,), generator=g) for _ in range(N_BATCH)]
    sess[0][0] = 99  # a session of length 1
    req = [{"session_id": s} for s in sess]

    def expected(decreasing=True, exp=False, k=None, drop_single=False, scale=False):
        total = count = 0.0
        for (p, l, w), s in zip(bs, sess):
            for sid in set(s.tolist()):
                idx = [i for i in range(B) if int(s[i]) == sid]
                if drop_single and len(idx) == 1:
                    continue
                pp, ll, ww = [float(p[i]) for i in idx], [float(l[i]) for i in idx], [float(w[i]) for i in idx]
                if scale:
                    pp, ll = [a * b for a, b in zip(pp, ww)], [a * b for a, b in zip(ll, ww)]
                kk = None
                if k is not None:  # the cut-off is min(k, longest session of the batch)
                    longest = max(sum(1 for x in s.tolist() if x == q) for q in set(s.tolist()))
                    kk = min(k, longest)
                v = _ndcg(pp, ll, kk, exp)
                v = 1 - v if decreasing else v
                total += v if scale else v * max(ww)
                count += 1
        return total / count

    m = _metric(M.NDCGMetric)
    assert m.get_required_inputs() == {"session_id"}
    _close(_feed(m, bs, required_inputs=req)["ndcg-t|lifetime_ndcg"], expected(), 1e-6)
    _close(_feed(_metric(M.NDCGMetric, report_ndcg_as_decreasing_curve=False, exponential_gain=True), bs, required_inputs=req)["ndcg-t|window_ndcg"],
           expected(decreasing=False, exp=True), 1e-6)
    _close(_feed(_metric(M.NDCGMetric, k=2, remove_single_length_sessions=True), bs, required_inputs=req)["ndcg-t|lifetime_ndcg"], expected(k=2, drop_single=True), 1e-6)
    _close(_feed(_metric(M.NDCGMetric, scale_by_weights_tensor=True), bs, required_inputs=req)["ndcg-t|lifetime_ndcg"], expected(scale=True), 1e-6)
    other = _metric(M.NDCGMetric, session_key="sid")
    assert other.get_required_inputs() == {"sid"}
    with pytest.raises(RecMetricException):
        _feed(other, bs, required_inputs=req)
    # negative task: predictions and labels are flipped first
    nb = [(p, (l > 1).float(), w) for p, l, w in bs]
    flipped = [(1 - p, 1 - l, w) for p, l, w in nb]
    a = _feed(_metric(M.NDCGMetric, is_negative_task_mask=[True]), nb, required_inputs=req)["ndcg-t|lifetime_ndcg"]
    b = _feed(_metric(M.NDCGMetric), flipped, required_inputs=req)["ndcg-t|lifetime_ndcg"]
    _close(a, float(b), 1e-9)


def test_session_recall_precision():
    bs = _batches(4)
    g = torch.Generator().manual_seed(9)
    sess = [torch.randint(0, 8, (B,), generator=g) for _ in range(N_BATCH)]
    bs = [(torch.round(p * 10) / 10, l, w) for p, l, w in bs]  # ties inside sessions
    req = [{"sess": s} for s in sess]
    for top, rank_labels in ((1, False), (2, False), (2, True)):
        task = RecTaskInfo(name="t", session_metric_def=SessionMetricDef(session_var_name="sess", top_threshold=top, run_ranking_of_labels=rank_labels))
        tp = fn = fp = 0.0
        for (p, l, w), s in zip(bs, sess):
            for i in range(B):
                peers = [j for j in range(B) if int(s[j]) == int(s[i])]
                pred_pos = sum(1 for j in peers if float(p[j]) > float(p[i])) < top
                lab = float(l[i])
                if rank_labels:
                    lab = 1.0 if sum(1 for j in peers if float(l[j]) > float(l[i])) < top else 0.0
                tp += float(w[i]) * lab * pred_pos
                fn += float(w[i]) * lab * (not pred_pos)
                fp += float(w[i]) * (1 - lab) * pred_pos
        rm = _metric(M.RecallSessionMetric, task=task)
        assert rm.get_required_inputs() == {"sess"}
        r = _feed(rm, bs, required_inputs=req)
        _close(r["recall_session_level-t|lifetime_recall_session_level"], tp / (tp + fn))
        _close(r["recall_session_level-t|window_recall_session_level"], tp / (tp + fn))
        r = _feed(_metric(M.PrecisionSessionMetric, task=task), bs, required_inputs=req)
        _close(r["precision_session_level-t|lifetime_precision_session_level"], tp / (tp + fp))
    with pytest.raises(RecMetricException):
        _metric(M.RecallSessionMetric)  # no session definition
    with pytest.raises(RecMetricException):
        _metric(M.RecallSessionMetric, task=RecTaskInfo(name="t", session_metric_def=SessionMetricDef(session_var_name="s")))  # no threshold
    with pytest.raises(RecMetricException):
        M.RecallSessionMetric(world_size=1, my_rank=0, batch_size=B, tasks=[task], compute_mode=RecComputeMode.FUSED_TASKS_COMPUTATION)
    plain = RecTaskInfo(name="t", session_metric_def=SessionMetricDef(session_var_name="sess", top_threshold=1))
    nothing = _metric(M.RecallSessionMetric, task=plain)
    nothing.update(predictions={"t": torch.rand(4)}, labels={"t": torch.zeros(4)}, weights={"t": torch.ones(4)}, required_inputs={"sess": torch.zeros(4, dtype=torch.long)})
    assert math.isnan(float(nothing.compute()["recall_session_level-t|lifetime_recall_session_level"]))


def test_hindsight_target_pr_and_segmented_ne():
    bs = _batches(5)
    p, l, w = _cat(bs)
    r = _feed(_metric(M.HindsightTargetPRMetric, target_precision=0.7), bs)
    th = torch.linspace(0, 1, 1000, dtype=torch.double).tolist()
    idx = None
    for i, t in enumerate(th):
        tp = sum(wi for pi, li, wi in zip(p, l, w) if pi >= t and li == 1)
        fp = sum(wi for pi, li, wi in zip(p, l, w) if pi >= t and li == 0)
        if tp + fp > 0 and tp / (tp + fp) >= 0.7:
            idx = i
            break
    assert idx is not None
    fn = sum(wi for pi, li, wi in zip(p, l, w) if pi < th[idx] and li == 1)
    for scope in ("lifetime", "window"):
        assert int(r[f"hindsight_target_pr-t|{scope}_hindsight_target_pr"]) == idx
        _close(r[f"hindsight_target_pr-t|{scope}_hindsight_target_precision"], tp / (tp + fp))
        _close(r[f"hindsight_target_pr-t|{scope}_hindsight_target_recall"], tp / (tp + fn))
    # segmented NE: two segmentations at once, log loss per group
    g = torch.Generator().manual_seed(10)
    ka = [torch.randint(0, 2, (B,), generator=g) for _ in range(N_BATCH)]
    kb = [torch.randint(0, 3, (B,), generator=g).float() for _ in range(N_BATCH)]
    m = _metric(M.SegmentedNEMetric, include_logloss=True, grouping_keys=[{"name": "ka", "num_groups": 2}, {"name": "kb", "num_groups": 3, "cast_keys_to_int": True}])
    assert m.get_required_inputs() == {"ka", "kb"}
    r = _feed(m, bs, required_inputs=[{"ka": a, "kb": b} for a, b in zip(ka, kb)])

    def ne_of(idx):
        ce = sum(-w[i] * (l[i] * math.log2(p[i]) + (1 - l[i]) * math.log2(1 - p[i])) for i in idx)
        W, pos = sum(w[i] for i in idx), sum(w[i] * l[i] for i in idx)
        return ce / -(pos * math.log2(pos / W) + (W - pos) * math.log2(1 - pos / W)), ce / W * math.log(2)

    for name, keys, n_groups in (("ka", torch.cat(ka).tolist(), 2), ("kb", torch.cat(kb).tolist(), 3)):
        for grp in range(n_groups):
            ne, ll = ne_of([i for i in range(len(p)) if int(keys[i]) == grp])
            _close(r[f"segmented_ne-t|lifetime_segmented_ne_{grp}@{name}"], ne, 1e-5)
            _close(r[f"segmented_ne-t|window_segmented_ne_{grp}@{name}"], ne, 1e-5)
            _close(r[f"segmented_ne-t|lifetime_logloss_{grp}@{name}"], ll, 1e-5)
    # the single default key keeps the short description
    r = _feed(_metric(M.SegmentedNEMetric, num_groups=2), bs, required_inputs=[{"grouping_keys": a} for a in ka])
    _close(r["segmented_ne-t|lifetime_segmented_ne_1"], ne_of([i for i in range(len(p)) if int(torch.cat(ka)[i]) == 1])[0], 1e-5)
    with pytest.raises(RecMetricException):  # float keys need the cast flag
        _feed(_metric(M.SegmentedNEMetric, num_groups=3), bs, required_inputs=[{"grouping_keys": b} for b in kb])


def test_tensor_weighted_avg_and_tower_qps():
    bs = _batches(6)
    g = torch.Generator().manual_seed(11)
    xs = [torch.rand(B, generator=g) for _ in range(N_BATCH)]
    ys = [torch.rand(B, generator=g) for _ in range(N_BATCH)]
    ta, tb = RecTaskInfo(name="a", tensor_name="x", weighted=True), RecTaskInfo(name="b", tensor_name="y", weighted=False)
    wts = torch.cat([b[2] for b in bs]).double()
    want_a = float((torch.cat(xs).double() * wts).sum() / wts.sum())
    want_b = float(torch.cat(ys).double().mean())
    for mode in (RecComputeMode.UNFUSED_TASKS_COMPUTATION, RecComputeMode.FUSED_TASKS_COMPUTATION):
        m = M.TensorWeightedAvgMetric(world_size=1, my_rank=0, batch_size=B, tasks=[ta, tb], window_size=10_000, compute_mode=mode)
        assert m.get_required_inputs() == {"x", "y"}
        for (p, l, w), x, y in zip(bs, xs, ys):
            m.update(predictions={"a": p, "b": p}, labels={"a": l, "b": l}, weights={"a": w, "b": w}, required_inputs={"x": x, "y": y})
        r = m.compute()
        _close(r["weighted_avg-a|lifetime_weighted_avg"], want_a)
        _close(r["weighted_avg-b|window_weighted_avg"], want_b)
    with pytest.raises(RecMetricException):
        M.TensorWeightedAvgMetric(world_size=1, my_rank=0, batch_size=B, tasks=[RecTaskInfo(name="a")], window_size=10_000)  # no tensor name
    with pytest.raises(RecMetricException):
        M.TensorWeightedAvgMetric(world_size=1, my_rank=0, batch_size=B, tasks=[ta, RecTaskInfo(name="c", tensor_name="x", weighted=False)], window_size=10_000,
                                  compute_mode=RecComputeMode.FUSED_TASKS_COMPUTATION)  # one tensor, two weightings
    # tower QPS: warm-up examples are counted in the total but not rated
    import time

    q = _metric(M.TowerQPSMetric, warmup_steps=2)
    for i in range(5):
        q.update(predictions={"t": bs[0][0]}, labels={"t": bs[0][1]}, weights={"t": bs[0][2]})
        time.sleep(0.01)
    r = q.compute()
    assert int(r["qps-t|total_examples"]) == 5 * B
    comp = q._metrics_computations[0]
    assert int(comp.warmup_examples) == 2 * B and float(comp.time_lapse) >= 0.03
    _close(r["qps-t|lifetime_qps"], 3 * B / float(comp.time_lapse))
    _close(r["qps-t|window_qps"], 3 * B / float(comp.time_lapse))


def test_scalar_output_and_multi_label_precision():
    # scalar: the latest value and the window mean of the per-batch values
    m = _metric(M.ScalarMetric)
    vals = [0.5, 2.0, 3.5]
    for v in vals:
        m.update(predictions={"t": torch.zeros(4)}, labels={"t": torch.full((4,), v)}, weights={"t": torch.ones(4)})
    r = m.compute()
    _close(r["scalar-t|lifetime_scalar"], 3.5)
    _close(r["scalar-t|window_scalar"], sum(vals) / 3)
    # output: batch means of two named model outputs, no lifetime / window prefix
    m = _metric(M.OutputMetric)
    assert m.get_required_inputs() == {"latest_imp", "total_latest_imp"}
    with pytest.raises(RecMetricException):
        m.update(predictions={"t": torch.zeros(4)}, labels={"t": torch.zeros(4)}, weights={"t": torch.ones(4)})
    m.update(predictions={"t": torch.zeros(4)}, labels={"t": torch.zeros(4)}, weights={"t": torch.ones(4)},
             required_inputs={"latest_imp": torch.tensor([1.0, 2.0, 3.0, 6.0]), "total_latest_imp": torch.tensor([10.0, 10.0, 20.0, 20.0])})
    r = m.compute()
    _close(r["output-t|output_latest_imp"], 3.0)
    _close(r["output-t|output_total_latest_imp"], 15.0)
    # multi-label precision: label sets encoded as bits (LSB first)
    m = _metric(M.MultiLabelPrecisionMetric, num_labels=3, label_names=["cat", "dog", "horse"])
    rng = random.Random(0)
    tp, fp = [0.0] * 3, [0.0] * 3
    for _ in range(3):
        p = [rng.randrange(8) for _ in range(16)]
        l = [rng.randrange(8) for _ in range(16)]
        w = [rng.random() + 0.1 for _ in range(16)]
        m.update(predictions={"t": torch.tensor(p)}, labels={"t": torch.tensor(l)}, weights={"t": torch.tensor(w)})
        for pi, li, wi in zip(p, l, w):
            for bit in range(3):
                if (pi >> bit) & 1:
                    if (li >> bit) & 1:
                        tp[bit] += wi
                    else:
                        fp[bit] += wi
    r = m.compute()
    for bit, name in enumerate(["cat", "dog", "horse"]):
        _close(r[f"multi_label_precision-t|lifetime_multi_label_precision{name}"], tp[bit] / (tp[bit] + fp[bit]), 1e-5)
        _close(r[f"multi_label_precision-t|window_multi_label_precision{name}"], tp[bit] / (tp[bit] + fp[bit]), 1e-5)


def test_fused_tasks_equal_unfused_tasks():
    """Stacking the tasks into one computation (FUSED_TASKS_COMPUTATION) gives the same report as one computation per task."""
    tasks = [RecTaskInfo(name=f"t{i}") for i in range(3)]
    g = torch.Generator().manual_seed(21)
    batches = []
    for _ in range(3):
        p = {t.name: torch.rand(B, generator=g) for t in tasks}
        l = {t.name: (torch.rand(B, generator=g) < 0.4).float() for t in tasks}
        w = {t.name: torch.rand(B, generator=g) + 0.1 for t in tasks}
        batches.append((p, l, w))
    sess = [torch.randint(0, 6, (B,), generator=g) for _ in range(3)]
    lens = [torch.tensor([10, 7, 13, 10]) for _ in range(3)]
    keys = [torch.randint(0, 3, (B,), generator=g) for _ in range(3)]
    cases = [(M.NEMetric, {}, None), (M.CalibrationMetric, {}, None), (M.CTRMetric, {}, None), (M.MSEMetric, {"include_r_squared": True}, None), (M.MAEMetric, {}, None),
             (M.AccuracyMetric, {}, None), (M.PrecisionMetric, {}, None), (M.RecallMetric, {}, None), (M.NMSEMetric, {}, None), (M.XAUCMetric, {}, None),
             (M.AverageMetric, {}, None), (M.WeightedAvgMetric, {}, None), (M.CaliFreeNEMetric, {}, None), (M.UnweightedNEMetric, {}, None),
             (M.HindsightTargetPRMetric, {}, None), (M.AUCMetric, {}, None), (M.AUPRCMetric, {}, None), (M.RAUCMetric, {}, None),
             (M.NDCGMetric, {}, lambda i: {"required_inputs": {"session_id": sess[i]}}), (M.GAUCMetric, {}, lambda i: {"num_candidates": lens[i]}),
             (M.SegmentedNEMetric, {"num_groups": 3}, lambda i: {"required_inputs": {"grouping_keys": keys[i]}}), (M.ServingNEMetric, {}, None)]
    for cls, kw, extra in cases:
        out = {}
        for mode in (RecComputeMode.UNFUSED_TASKS_COMPUTATION, RecComputeMode.FUSED_TASKS_COMPUTATION):
            m = cls(world_size=1, my_rank=0, batch_size=B, tasks=tasks, window_size=10_000, compute_mode=mode, **kw)
            for i, (p, l, w) in enumerate(batches):
                m.update(predictions=p, labels=l, weights=w, **(extra(i) if extra else {}))
            out[mode] = m.compute()
        a, b = out[RecComputeMode.UNFUSED_TASKS_COMPUTATION], out[RecComputeMode.FUSED_TASKS_COMPUTATION]
        assert set(a) == set(b) and len(a) >= 3, (cls.__name__, sorted(a), sorted(b))
        for k in a:
            torch.testing.assert_close(torch.as_tensor(a[k]).double().reshape(-1), torch.as_tensor(b[k]).double().reshape(-1), rtol=1e-9, atol=1e-12, msg=lambda m_: f"{cls.__name__} {k}: {m_}")
