"""Summarise an ncu capture of the column-mode stage kernels against their algorithmic bytes.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \\
        --clock-control none --csv --log-file gpurun_out/column_mode.csv \\
        -k regex:'geom_kernel|tau_kernel|alt_dec_kernel|lat_long_kernel|compact|tile_|mc_pass' \\
        python tools/column_mode_probe.py 2097152
    python tools/column_mode_report.py gpurun_out/column_mode.csv 2097152 <n_valid> \\
        > profiles/r02_column_mode.json

Per kernel (LAST launch of each name = the warm second compute() call): duration, DRAM bytes
read + written, the algorithmic bytes of the stage (arrays it must read and write once, float64
columns as the reference's), achieved algorithmic GB/s and its fraction of the measured HBM
copy bandwidth (MEASURED_PEAKS.json: 6535.1 GB/s).  ncu serialises launches and starts each
from a cold L2, so durations are upper bounds of the in-pipeline times.
"""
import csv
import json
import os
import sys

HBM_GBS = 6535.1
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
        HBM_GBS = float(json.load(fh).get("hbm_gbs", HBM_GBS))
except Exception:
    pass


def algorithmic_bytes(name, N, V):
    """(bytes read, bytes written) the stage must move, per launch."""
    if "geom_kernel" in name:       # u (4,N) f64 in; 13 f64 arrays + 1 byte mask out
        return 32 * N, (13 * 8 + 1) * N
    if "tile_count" in name:        # mask in; one count per 4096-event tile out
        return N, 8 * (N // 4096 + 1)
    if "tile_scan" in name:
        return 8 * (N // 4096 + 1), 8 * (N // 4096 + 1)
    if "tile_pos" in name:          # mask in; int64 position per thrown event out
        return N, 8 * N
    if "compact_columns" in name:   # mask + pos + 11 columns in; 11 compacted columns out
        return (1 + 8 + 11 * 8) * N, 11 * 8 * V
    if "compact_kernel" in name:    # mask + pos + 1 column in; 1 compacted column out
        return (1 + 8 + 8) * N, 8 * V
    if "lat_long" in name:          # 7 columns in, 2 out (valid events)
        return 7 * 8 * V, 2 * 8 * V
    if "tau_kernel" in name:        # beta, log_e_nu, u in; 5 columns out
        return 3 * 8 * V, 5 * 8 * V
    if "alt_dec" in name:           # beta, tauBeta, tauLorentz, u in; altDec, lenDec out
        return 4 * 8 * V, 2 * 8 * V
    if "mc_pass" in name:           # 6 columns in; per-block partials out
        return 6 * 8 * V, 0
    return None


def main():
    path, N, V = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    rows = []
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    rd = csv.DictReader(lines)
    per = {}
    for r in rd:
        key = (r["ID"], r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6,
                 "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3,
                 "second": 1.0}.get(unit, 1)
        per.setdefault(key, {})[r["Metric Name"]] = v * scale
    last = {}
    for (kid, name), m in per.items():
        short = name.split("(")[0]
        if short not in last or int(kid) > last[short][0]:
            last[short] = (int(kid), m)
    out = {}
    for short, (kid, m) in sorted(last.items(), key=lambda kv: kv[1][0]):
        ab = algorithmic_bytes(short, N, V)
        t = m.get("gpu__time_duration.sum")
        dram = m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        e = dict(launch_id=kid, duration_us=t * 1e6, dram_bytes_read=m.get("dram__bytes_read.sum"),
                 dram_bytes_write=m.get("dram__bytes_write.sum"), dram_bytes=dram)
        if ab is not None:
            alg = float(ab[0] + ab[1])
            e.update(algorithmic_bytes_read=ab[0], algorithmic_bytes_write=ab[1],
                     algorithmic_bytes=alg, traffic_over_algorithmic=dram / alg if alg else None,
                     achieved_gbs=alg / t * 1e-9, hbm_peak_gbs=HBM_GBS,
                     frac_of_hbm=alg / t * 1e-9 / HBM_GBS)
        out[short] = e
    print(json.dumps(dict(n_thrown=N, n_valid=V, hbm_peak_gbs=HBM_GBS,
                          peak_source="MEASURED_PEAKS.json hbm_gbs (driver-measured copy bandwidth)",
                          kernels=out), indent=1))


if __name__ == "__main__":
    main()
