"""Is the single-thread tcgen05.mma issuer of a kernel waiting or busy?  From an `ncu --set full --import-source on` report: the
warp-state samples that fall into the SASS region around the UTCHMMA instructions (the MMA warp's loop), split into samples on
mbarrier waits (SYNCS ... TRYWAIT and the branch that follows it) and samples on everything else (descriptor arithmetic, ELECT /
R2UR moves, the UTCHMMA / UTCBAR instructions themselves), plus the instruction count between the first and the last UTCHMMA.
python tools/ncu_issue_warp.py report.ncu-rep [kernel substring] [lead=150] [tail=60]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    only = sys.argv[2] if len(sys.argv) > 2 else None
    lead = int(sys.argv[3]) if len(sys.argv) > 3 else 150
    tail = int(sys.argv[4]) if len(sys.argv) > 4 else 60
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
    seen = set()
    for k in kernels:
        if (only and only not in k["name"]) or k["name"] in seen:
            continue
        seen.add(k["name"])
        h = {c: i for i, c in enumerate(k["hdr"])}
        si, ii, so = h["# Samples"], h["Instructions Executed"], h["Source"]
        rows = k["rows"]
        um = [i for i, r in enumerate(rows) if "UTCHMMA" in r[so]]
        if not um:
            continue
        tot = sum(int(r[si] or 0) for r in rows) or 1
        lo, hi = max(0, um[0] - lead), min(len(rows), um[-1] + tail)
        region = rows[lo:hi]
        wait = 0
        for j, r in enumerate(region):
            s = r[so]
            prev = region[j - 1][so] if j else ""
            if "TRYWAIT" in s or ("BRA" in s and "TRYWAIT" in prev):
                wait += int(r[si] or 0)
        reg = sum(int(r[si] or 0) for r in region)
        execs = max(int(rows[i][ii] or 0) for i in um)
        print("==== %s" % k["name"][:140])
        print("     %d SASS instructions, %d samples; UTCHMMA: %d sites, rows %d..%d, hottest site executed %d times" %
              (len(rows), tot, len(um), um[0], um[-1], execs))
        print("     issue region rows [%d, %d): %d samples = %.1f %% of the kernel's; on mbarrier waits %d (%.0f %% of the region), "
              "on everything else %d (%.0f %%)" % (lo, hi, reg, 100.0 * reg / tot, wait, 100.0 * wait / max(reg, 1), reg - wait,
                                                 100.0 * (reg - wait) / max(reg, 1)))
        n_between = um[-1] - um[0] + 1
        print("     %d instructions between the first and the last UTCHMMA = %.1f per tcgen05.mma" % (n_between, n_between / len(um)))


if __name__ == "__main__":
    main()
