"""Golden vectors for the moment-retrieval metrics, produced by the REFERENCE's own evaluation code
(lavis/tasks/mr_eval.py eval_submission + lavis/tasks/mr_utils.py, loaded by file path) on synthetic predictions.
Run in the build container (needs /root/reference): python tests/golden/make_golden_mr_eval.py
-> tests/golden/mr_eval_golden.json"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402


def synth_records(n, seed):
    rng = np.random.RandomState(seed)
    recs = []
    for i in range(n):
        dur = float(rng.randint(30, 151))
        ngt = int(rng.randint(1, 4))
        gts = []
        for _ in range(ngt):
            s = float(rng.randint(0, int(dur) - 4))
            gts.append([s, float(min(dur, s + rng.randint(2, 40)))])
        kind = rng.rand()
        if kind < 0.1:
            preds = [[-1, -1]]                                   # unparsable prediction (utils.py:300-341)
        else:
            preds = []
            for k in range(int(rng.randint(1, 4))):
                if rng.rand() < 0.6:                             # near a target
                    g = gts[int(rng.randint(0, ngt))]
                    s = max(0.0, g[0] + float(rng.randint(-6, 7)))
                    e = max(s + 1.0, g[1] + float(rng.randint(-6, 7)))
                else:
                    s = float(rng.randint(0, int(dur) - 2))
                    e = s + float(rng.randint(1, 30))
                preds.append([s, e])
        recs.append({"qid": "q%d_%d" % (i, i % 7), "pred_relevant_windows": preds, "relevant_windows": gts})
    return recs


def main():
    ref_shim._install_shims()
    ref_shim._stub("lavis.tasks")
    mu = ref_shim._load("lavis.tasks.mr_utils", "lavis/tasks/mr_utils.py")
    me = ref_shim._load("lavis.tasks.mr_eval", "lavis/tasks/mr_eval.py")
    out = {"cases": []}
    for n, seed in ((40, 0), (7, 1), (120, 2)):
        recs = synth_records(n, seed)
        res = me.eval_submission(recs, recs, verbose=False)
        full = res["full"]
        out["cases"].append({"records": recs, "MR-mAP": full["MR-mAP"], "MR-R1": full["MR-R1"],
                             "MR-R1-avg": float(full["MR-R1-avg"]), "MR-mIoU": float(full["MR-mIoU"]),
                             "MR-invalid_pred_num": int(full["MR-invalid_pred_num"])})
    # a single-query AP table for the matching rule
    gt = [{"video-id": "v", "t-start": 10.0, "t-end": 20.0}, {"video-id": "v", "t-start": 30.0, "t-end": 50.0}]
    pr = [{"video-id": "v", "t-start": 11.0, "t-end": 19.0}, {"video-id": "v", "t-start": 9.0, "t-end": 21.0},
          {"video-id": "v", "t-start": 28.0, "t-end": 45.0}, {"video-id": "v", "t-start": 100.0, "t-end": 110.0}]
    thds = [float("%.2f" % t) for t in np.linspace(0.5, 0.95, 10)]
    out["single_ap"] = {"gt": [[g["t-start"], g["t-end"]] for g in gt], "pred": [[p["t-start"], p["t-end"]] for p in pr],
                        "ap": mu.compute_average_precision_detection(gt, pr, tiou_thresholds=thds).tolist()}
    json.dump(out, open(os.path.join(HERE, "mr_eval_golden.json"), "w"))
    print({k: out["cases"][0][k] for k in ("MR-mAP", "MR-R1", "MR-R1-avg", "MR-mIoU", "MR-invalid_pred_num")})
    print(out["single_ap"]["ap"])


if __name__ == "__main__":
    main()
