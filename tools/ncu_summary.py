"""ncu --page raw --csv export of ONE kernel -> the handful of metrics the roofline discussion uses.  python tools/ncu_summary.py raw.csv out.csv"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "gpc__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__cluster_dim_x",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, vals = rows[0], rows[1], rows[2]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(sys.argv[2], "w") as f:
        f.write("# %s\n" % vals[idx["Kernel Name"]][:160])
        f.write("metric,unit,value\n")
        for k in KEYS:
            if k in idx:
                f.write("%s,%s,%s\n" % (k, units[idx[k]], vals[idx[k]]))
    print(open(sys.argv[2]).read())


if __name__ == "__main__":
    main()
