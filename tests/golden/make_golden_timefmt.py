"""Golden vectors for the other input_time_format helpers of the reference
(lavis/models/blip2_mr_models/utils.py:242-297 convert_to_absolute_time, :437-512 relative_integers / seconds_floats /
relative_floats), produced by the REFERENCE's own functions through ref_shim.  Run in the build container (needs
/root/reference): python tests/golden/make_golden_timefmt.py  ->  tests/golden/time_formats_golden.json"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402


def main():
    ru = ref_shim.load_reference_utils()
    ts = [torch.tensor([1.25, 3.75, 6.5, 149.6]), torch.tensor([0.4, 10.5, 11.5, 29.9]), torch.tensor([0.0, 33.333, 66.667, 99.99])]
    du = torch.tensor([150.0, 30.2, 100.0])
    out = {"timestamps": [t.tolist() for t in ts], "durations": du.tolist()}
    for name in ("relative_integers", "seconds_floats", "relative_floats"):
        t, d, p = getattr(ru, "get_timestamps_as_" + name)(ts, du, {})
        out[name] = {"ts": [x.tolist() for x in t], "ts_str": [[str(v.item()) for v in x] for x in t],
                     "dur": [float(x) for x in d], "prompt": p}
    preds = ["[[10, 25], [50, 100]]", "[[-1, -1]]", "[[0, 7]]"]
    fpreds = ["[[0.1, 0.25], [0.5, 1.0]]", "[[-1, -1]]", "[[0.0, 0.07]]"]
    out["absolute"] = {"relative_integers": [preds, ru.convert_to_absolute_time(preds, du.tolist(), "relative_integers")],
                       "relative_floats": [fpreds, ru.convert_to_absolute_time(fpreds, du.tolist(), "relative_floats")]}
    json.dump(out, open(os.path.join(HERE, "time_formats_golden.json"), "w"), indent=1)
    print(json.dumps(out, indent=1)[:1500])


if __name__ == "__main__":
    main()
