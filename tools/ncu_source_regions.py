"""Development aid: executed warp-instructions of eas_walk_kernel per source region.

    ncu -i gpurun_out/<rep>.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python tools/ncu_source_regions.py src.csv [events_in_launch] [nss_eas_core.cuh as profiled]

The capture must come from a build with -lineinfo and `--import-source on`.  Lines of
nss_eas_core.cuh are grouped into the regions of the walk by the markers in REGIONS (first line
of each region is found by text search in the source column, so the grouping follows the code
as it was profiled)."""
import csv
import json
import sys

REGIONS = [  # (name, text that starts the region in nss_eas_core.cuh)
    ("event setup", "if (E.n_raw <= 0) return;"),
    ("altitude step geometry + grammage + ozone", "const int k = base + lane;"),
    ("prefix sum, AirN, Greisen/GH profile, keep", "const float gincl = gcarry"),
    ("tail-skip vote and light path", "if (!Ops::any(keep && !(s > 1.0f && may_stop"),
    ("per-step: E0, threshold, thetaC, Tfrac, D", "if (s >= 0.4f) {"),
    ("per-step: aerosol + wavelength sum", "if (z < 30.0f) {"),
    ("per-step: spsum, tanC, bin count", "float spsum = (PY * RN) * (acc0 + acc1);"),
    ("per-step: energy-decade constants, bin limits, cold decades", "const float inv_e2 = fast_rcp(e2hill);"),
    ("radial bins, quantised zone (steep decades)", "const float mk1 = "),
    ("radial bins, closed forms + edge corrections", "const float h2 = invD * invD"),
    ("per-step accumulation", "a_phot += (double)inner * (double)spsum;"),
    ("event totals", "const int nv = Ops::sum_i(n_val);"),
]


def main():
    path = sys.argv[1]
    events = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = list(csv.reader(open(path)))
    # `--print-source cuda,sass`: per file, one row per source line followed by the SASS rows
    # attributed to it.  An instruction inlined through several levels is listed under every
    # level, so instructions are keyed by ADDRESS and counted once: under the deepest line of
    # the walk body that lists them (the call site), else under the helper line.
    addr = {}   # address -> dict(n=executed, mufu=bool, lines=[(file, line, src)])
    cur, hdr, line = None, None, None
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur, hdr = r[1], None
            continue
        if cur is None or not r:
            continue
        if r[0] == "Line No":
            hdr = r
            i_exec = hdr.index("Instructions Executed")
            continue
        if hdr is None:
            continue
        if r[0].isdigit():
            line = (cur, int(r[0]), r[1])
            continue
        if r[0] == "" and len(r) > i_exec and r[2].startswith("0x"):
            try:
                n = float(r[i_exec])
            except ValueError:
                continue
            a = addr.setdefault(r[2], dict(n=n, sass=r[3].strip(), lines=[]))
            a["lines"].append(line)
    total = sum(a["n"] for a in addr.values())
    core_lines = {}
    for a in addr.values():
        for f, l, src in a["lines"]:
            if f.endswith("nss_eas_core.cuh"):
                core_lines[l] = src
    starts = []
    src_lines = None
    if len(sys.argv) > 3:  # the header as it was profiled: markers -> line numbers
        src_lines = open(sys.argv[3]).read().split("\n")
    for name, marker in REGIONS:
        key = marker.strip()
        if src_lines is not None:
            ln = next((i + 1 for i, t in enumerate(src_lines) if key in t), None)
        else:
            ln = next((l for l in sorted(core_lines) if key in core_lines[l]), None)
        if ln is not None:
            starts.append((ln, name))
    starts.sort()
    first = starts[0][0] if starts else 10**9
    reg = {name: dict(warp_instructions=0.0, mufu=0.0) for _, name in starts}
    other = dict(warp_instructions=0.0, mufu=0.0)
    helper_of_region = {}
    # The listing names only the innermost inlined function of an instruction.  Instructions of
    # helpers (math wrappers, decade constants, closed-form pieces ...) are therefore given to
    # the region of the nearest preceding instruction, in address order, that carries a line of
    # the walk body -- the code is laid out in program order, so this follows the call sites.
    last = None
    for key in sorted(addr, key=lambda x: int(x, 16)):
        a = addr[key]
        walk = [l for f, l, _ in a["lines"] if f.endswith("nss_eas_core.cuh") and l >= first]
        if walk:
            last = [nm for s0, nm in starts if s0 <= max(walk)][-1]
        tgt = reg[last] if last is not None else other
        tgt["warp_instructions"] += a["n"]
        if "MUFU" in a["sass"]:
            tgt["mufu"] += a["n"]
    for v in list(reg.values()) + [other]:
        v["share"] = v["warp_instructions"] / total
    out = dict(total_warp_instructions=total, regions=reg,
               kernel_wrapper_and_unattributed=other,
               mufu_total=sum(a["n"] for a in addr.values() if "MUFU" in a["sass"]))
    if events:
        out["events_in_launch"] = events
        out["warp_instructions_per_event"] = total / events
        for v in reg.values():
            v["per_event"] = v["warp_instructions"] / events
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
