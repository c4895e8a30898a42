"""Top warp-stall-sampled SASS instructions of every kernel in an ncu report (needs --set full / source counters).
python tools/ncu_hot.py report.ncu-rep [N=25] [kernel substring]"""
import csv
import io
import subprocess
import sys


def main():
    rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
    only = sys.argv[3] if len(sys.argv) > 3 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    kernels, cur = [], None
    for row in csv.reader(io.StringIO(txt)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "hdr": None, "rows": []}
            kernels.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = row
        elif cur is not None and row:
            cur["rows"].append(row)
    for k in kernels:
        if only and only not in k["name"]:
            continue
        h = {c: i for i, c in enumerate(k["hdr"])}
        si, ii = h["# Samples"], h["Instructions Executed"]
        tot = sum(int(r[si] or 0) for r in k["rows"]) or 1
        print("==== %s\n     %d instructions, %d samples" % (k["name"][:150], len(k["rows"]), tot))
        order = sorted(range(len(k["rows"])), key=lambda i: -int(k["rows"][i][si] or 0))[:n]
        for i in sorted(order):
            r = k["rows"][i]
            print("%5d %6.2f%% %9s  %s" % (i, 100.0 * int(r[si] or 0) / tot, r[ii], r[h["Source"]].strip()[:110]))


if __name__ == "__main__":
    main()
