"""ncu launch list (--csv --metrics gpu__time_duration.sum) -> per-kernel summary csv.  python tools/summarize_launches.py in.csv out.csv"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        v = float(r[idx["Metric Value"]].replace(",", ""))
        u = r[idx["Metric Unit"]]
        v = v / 1e3 if u == "us" else v / 1e6 if u == "ns" else v * 1e3 if u == "s" else v
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    with open(sys.argv[2], "w") as f:
        f.write("# ncu launch list of ONE graph-replayed QVH training step (B=4, T=60): every kernel node of the captured step\n")
        f.write("# command: ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv python tools/profile_one_step.py\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write("kernel,launches,total_ms,share\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write("%s,%d,%.3f,%.4f\n" % (k.replace(",", ";"), cnt[k], v, v / total))
        f.write("TOTAL,%d,%.3f,1.0\n" % (sum(cnt.values()), total))
    print("total %.1f ms in %d launches" % (total, sum(cnt.values())))


if __name__ == "__main__":
    main()
