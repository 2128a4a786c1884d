"""Summarise an .ncu-rep (ncu --set full) into the per-kernel metric table kept under profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--stalls] > profiles/rNN_ncu_x.txt
"""
import csv
import io
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers"]


def main():
    rep = sys.argv[1]
    stalls = "--stalls" in sys.argv
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("%-78s %s %s" % (w, r[i][:70], units[i] if w != "Kernel Name" else ""))
        if stalls:
            st = []
            for i, h in enumerate(hdr):
                if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") \
                        or h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio"):
                    try:
                        st.append((float(r[i]), h))
                    except ValueError:
                        pass
            for v, h in sorted(st, reverse=True)[:6]:
                print("%-78s %.3f" % (h, v))
        print()


if __name__ == "__main__":
    main()
