#!/usr/bin/env python
"""Summarise one .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md / bench.py cite.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r01_x   -> r01_x_summary.json, r01_x_details.txt, r01_x_raw.csv
"""
import csv
import io
import json
import subprocess
import sys

KEEP = {
    "duration_ms": ("gpu__time_duration.sum", 1e-6, "ns"),
    "sm_mhz": ("smsp__cycles_elapsed.avg.per_second", 1e-6, None),
    "dram_read_bytes": ("dram__bytes_read.sum", None, None),
    "dram_write_bytes": ("dram__bytes_write.sum", None, None),
    "pipe_fma_active_pct": ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", 1, None),
    "pipe_fmaheavy_active_pct": ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", 1, None),
    "pipe_alu_active_pct": ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", 1, None),
    "pipe_fp64_active_pct": ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 1, None),
    "pipe_xu_inst_pct": ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1, None),
    "pipe_tensor_inst_pct": ("sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", 1, None),
    "pipe_lsu_inst_pct": ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1, None),
    "issue_slots_busy_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1, None),
    "sm_throughput_pct": ("sm__throughput.avg.pct_of_peak_sustained_elapsed", 1, None),
    "achieved_occupancy_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1, None),
    "registers_per_thread": ("launch__registers_per_thread", 1, None),
    "stall_math_pipe_throttle": ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", 1, None),
    "stall_wait": ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", 1, None),
    "stall_barrier": ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", 1, None),
    "stall_mio_throttle": ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", 1, None),
    "stall_dispatch": ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", 1, None),
    "inst_executed": ("smsp__inst_executed.sum", 1, None),
}
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9,
              "usecond": 1e3, "msecond": 1e6, "nsecond": 1.0, "second": 1e9}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(out + "_raw.csv", "w").write(raw)
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    open(out + "_details.txt", "w").write(det)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kernels = []
    for vals in rows[2:]:
        rec = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        s = {"kernel": rec.get("Kernel Name"), "grid": rec.get("Grid Size"), "block": rec.get("Block Size")}
        for k, (name, _scale, _) in KEEP.items():
            if name not in rec or rec[name] in ("", "n/a"):
                continue
            v = float(rec[name].replace(",", ""))
            unit = u.get(name, "")
            if k.endswith("_bytes"):
                v *= UNIT_SCALE.get(unit, 1.0)
            elif k == "duration_ms":
                v = v * UNIT_SCALE.get(unit, 1.0) / 1e6
            elif k == "sm_mhz":
                v = v * {"Ghz": 1e3, "Mhz": 1.0, "hz": 1e-6, "Khz": 1e-3}.get(unit, 1.0)
            s[k] = round(v, 4)
        if "dram_read_bytes" in s:
            s["dram_traffic_bytes"] = s["dram_read_bytes"] + s.get("dram_write_bytes", 0.0)
        kernels.append(s)
    json.dump({"report": rep, "kernels": kernels}, open(out + "_summary.json", "w"), indent=1)
    print(json.dumps(kernels, indent=1))


if __name__ == "__main__":
    main()
