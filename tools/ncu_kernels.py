"""Key metrics (time, occupancy, pipe utilisation, DRAM bytes, warp-stall samples) of every kernel in an `ncu --set full` report.
python tools/ncu_kernels.py report.ncu-rep out.md"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__cluster_dim_x",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes.sum.per_second", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    with open(out, "w") as f:
        f.write("# %s (ncu --set full --clock-control none; one launch each)\n" % rep.split("/")[-1])
        for r in rows[2:]:
            f.write("\n## %s\n\n| metric | value | unit |\n|---|---|---|\n" % r[idx["Kernel Name"]][:140].replace("|", "/"))
            for k in KEYS:
                if k in idx:
                    f.write("| %s | %s | %s |\n" % (k, r[idx[k]], units[idx[k]]))
            vals = []
            for h in stalls:
                try:
                    vals.append((float(r[idx[h]].replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
            tot = sum(v for v, _ in vals) or 1.0
            f.write("\nwarp-stall samples (all warps, share of %d): " % tot)
            f.write(", ".join("%s %.1f%%" % (n, 100 * v / tot) for v, n in sorted(vals, reverse=True)[:8]) + "\n")
    print(open(out).read())


if __name__ == "__main__":
    main()
