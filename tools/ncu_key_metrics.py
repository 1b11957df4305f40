"""Development aid: key metrics of one kernel launch from an .ncu-rep -> JSON (numbers in base
units: bytes, seconds, percent).

    python tools/ncu_key_metrics.py gpurun_out/r02_eas_v13.ncu-rep <events_in_launch> \\
        "<command the capture came from>" > profiles/r02_eas_v13_ncu_key_metrics.json
"""
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.per_cycle_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "smsp__cycles_active.avg",
]
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3,
         "s": 1.0, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}


def main():
    rep, events = sys.argv[1], float(sys.argv[2])
    cmd = sys.argv[3] if len(sys.argv) > 3 else ""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {}
    for i, h in enumerate(hdr):
        stall = "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "average" in h
        if h in KEYS or stall or h == "Kernel Name":
            v = vals[i]
            try:
                v = float(v.replace(",", "")) * SCALE.get(units[i], 1.0)
            except ValueError:
                pass
            out[h] = v
    out["events_in_launch"] = events
    out["warp_instructions_per_event"] = out["smsp__inst_executed.sum"] / events
    out["dram_bytes_per_event"] = (out["dram__bytes_read.sum"] + out["dram__bytes_write.sum"]) / events
    out["_units"] = "bytes, seconds, percent, counts"
    out["_provenance"] = cmd + f" ({rep})"
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
