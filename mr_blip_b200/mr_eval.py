"""Moment-retrieval metrics for the results MomentRetrievalTask collects from BLIP2_MR.generate: R1@IoU, mIoU, mAP@IoU
(0.5 ... 0.95) and the invalid-prediction count, as `lavis/tasks/moment_retrieval.py:115-152` (_report_metrics) computes
them through `lavis/tasks/mr_eval.py:26-218,330-417` and `lavis/tasks/mr_utils.py:16-171` (the Moment-DETR evaluation).

SURVEY.md §8f rank 4 ("next" row): host-side numpy after decoding, single process (the reference fans the per-query AP
out to an mp.Pool of 8).  Numbers are pinned to the reference's own functions run on synthetic predictions
(tests/golden/mr_eval_golden.json, generator tests/golden/make_golden_mr_eval.py).

Conventions kept from the reference: the reference calls eval_submission(results, results), i.e. predictions and targets
travel in the same records; R1 uses the FIRST predicted window against the target window it overlaps most; its IoU
divides by the hull of the two windows (max end - min start), not by the true union; AP matches predictions to targets
greedily in the order they were predicted (no scores) and integrates the VOC-2011 interpolated precision envelope."""
import json
from collections import OrderedDict

import numpy as np

from .mr_utils import moment_str_to_list

IOU_THRESHOLDS = tuple(float("%.2f" % t) for t in np.linspace(0.5, 0.95, 10))


def iou_hull(pred, gt):
    """[N,2] x [N,2] -> [N]: intersection / (max end - min start), 0 where that span is empty (mr_utils.py:16-37)."""
    pred, gt = np.asarray(pred, dtype=float), np.asarray(gt, dtype=float)
    inter = np.maximum(0, np.minimum(pred[:, 1], gt[:, 1]) - np.maximum(pred[:, 0], gt[:, 0]))
    hull = np.maximum(pred[:, 1], gt[:, 1]) - np.minimum(pred[:, 0], gt[:, 0])
    return np.divide(inter, hull, out=np.zeros_like(inter), where=hull != 0)


def iou_cross(a, b):
    """[N,2] x [M,2] -> [N,M] true temporal IoU (mr_utils.py:40-67)."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    inter = np.clip(np.minimum(a[:, None, 1], b[None, :, 1]) - np.maximum(a[:, None, 0], b[None, :, 0]), 0, None)
    union = (a[:, 1] - a[:, 0])[:, None] + (b[:, 1] - b[:, 0])[None, :] - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        return inter / union


def _interpolated_ap(precision, recall):
    """Area under the monotone precision envelope (mr_utils.py:70-86)."""
    p = np.concatenate([[0.0], precision, [0.0]])
    r = np.concatenate([[0.0], recall, [1.0]])
    p = np.maximum.accumulate(p[::-1])[::-1]
    step = np.nonzero(r[1:] != r[:-1])[0] + 1
    return float(np.sum((r[step] - r[step - 1]) * p[step]))


def average_precision(gt_windows, pred_windows, thresholds=IOU_THRESHOLDS):
    """Detection AP of ONE query at each IoU threshold (mr_utils.py:89-171): predictions are visited in the given order;
    each claims the not-yet-claimed target it overlaps most, provided that overlap reaches the threshold."""
    nt, ng, npred = len(thresholds), len(gt_windows), len(pred_windows)
    ap = np.zeros(nt)
    if npred == 0:
        return ap
    tp = np.zeros((nt, npred))
    claimed = np.zeros((nt, ng), dtype=bool)
    ious = iou_cross(pred_windows, gt_windows) if ng else np.zeros((npred, 0))
    for i in range(npred):
        order = np.argsort(ious[i])[::-1]
        for t, thd in enumerate(thresholds):
            for j in order:
                if not ious[i, j] >= thd:          # sorted by overlap: nothing further can match (NaN counts as a miss)
                    break
                if not claimed[t, j]:
                    claimed[t, j] = True
                    tp[t, i] = 1
                    break
    tp_cum = np.cumsum(tp, axis=1)
    fp_cum = np.cumsum(1 - tp, axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        recall = tp_cum / float(ng)
        precision = tp_cum / (tp_cum + fp_cum)
    for t in range(nt):
        ap[t] = _interpolated_ap(precision[t], recall[t])
    return ap


def moment_retrieval_metrics(records, thresholds=IOU_THRESHOLDS):
    """records: [{"qid", "pred_relevant_windows": [[s, e], ...], "relevant_windows": [[s, e], ...]}]
    -> {"MR-mAP": {thd: %, "average": %}, "MR-R1": {thd: %}, "MR-R1-avg", "MR-mIoU", "MR-invalid_pred_num"}
    (mr_eval.py:26-140; the last record wins when a qid repeats, as in the reference's dicts)."""
    by_qid = OrderedDict((r["qid"], r) for r in records)
    # ---- mAP: every predicted window of a query against all of its target windows
    aps = np.array([average_precision([w[:2] for w in r["relevant_windows"]], [w[:2] for w in r["pred_relevant_windows"]],
                                      thresholds)
                    for r in by_qid.values() if len(r["pred_relevant_windows"]) > 0])
    ap_thd = aps.mean(0)
    m_ap = {str(t): float("%.2f" % (100 * v)) for t, v in zip(thresholds, ap_thd)}
    m_ap["average"] = float("%.2f" % (100 * np.mean(ap_thd)))
    # ---- R1 / mIoU: first predicted window vs the target window it overlaps most
    first = np.array([r["pred_relevant_windows"][0][:2] for r in by_qid.values()], dtype=float)
    best = []
    for p, r in zip(first, by_qid.values()):
        gts = r["relevant_windows"]
        k = int(np.argmax(iou_cross(p[None], np.array(gts, dtype=float))[0])) if len(gts) > 0 else 0
        best.append(gts[k][:2])
    iou = iou_hull(first, np.array(best, dtype=float))
    r1 = {str(t): float("%.2f" % (np.mean(iou >= t) * 100)) for t in thresholds}
    return {"MR-mAP": m_ap, "MR-R1": r1, "MR-R1-avg": float(np.mean(list(r1.values()))), "MR-mIoU": float(np.mean(iou)),
            "MR-invalid_pred_num": int(sum(1 for p in first if -1 in p))}


def report_metrics(results):
    """MomentRetrievalTask._report_metrics (moment_retrieval.py:115-152) on the list valid_step builds
    ({"qid", "prediction", "target", ...} with window strings) or on the path of its JSON dump."""
    if isinstance(results, str):
        results = json.load(open(results))
    records = [{"qid": r["qid"], "pred_relevant_windows": moment_str_to_list(r["prediction"]),
                "relevant_windows": moment_str_to_list(r["target"])} for r in results]
    m = moment_retrieval_metrics(records)
    return {"agg_metrics": m["MR-R1-avg"], "r1": m["MR-R1"], "mAP": m["MR-mAP"], "mIoU": m["MR-mIoU"],
            "invalid_predictions": m["MR-invalid_pred_num"] / len(results), "total": len(results)}
