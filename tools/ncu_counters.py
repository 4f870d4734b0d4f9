#!/usr/bin/env python
"""profiles/r02_ncu_counters.json from `ncu --set full` captures (gpurun_out/*.ncu-rep): per kernel and workload the DRAM
bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum), the tensor / IMMA / FP64 pipe activity, issue activity and
the duration under ncu.  bench.py reads `traffic` from this file (never a typed-in constant).
usage: python tools/ncu_counters.py <workload>:<rep> [...]   (appends / replaces records keyed by (kernel, workload))"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r02_ncu_counters.json")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
        "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def val(hdr, units, row, name, scale=True):
    for i, h in enumerate(hdr):
        if h == name or h.endswith("." + name):
            try:
                v = float(row[i].replace(",", ""))
            except ValueError:
                return None
            return v * UNIT.get(units[i], 1.0) if scale else v
    return None


def main():
    db = {"kernels": []}
    if os.path.exists(OUT):
        db = json.load(open(OUT))
    try:
        head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
    except Exception:
        head = None
    for arg in sys.argv[1:]:
        workload, rep = arg.split(":", 1)
        hdr, units, rows = raw(rep)
        # a capture may hold several launches of one kernel (set-up launches of a few microseconds next to the screen's): keep
        # the longest launch of every kernel name
        dur_i = next(i for i, h in enumerate(hdr) if h == "gpu__time_duration.sum" or h.endswith(".gpu__time_duration.sum"))
        best = {}
        for row in rows:
            k = row[hdr.index("Kernel Name")]
            try:
                dv = float(row[dur_i].replace(",", "")) * UNIT.get(units[dur_i], 1.0)
            except ValueError:
                continue
            if k not in best or dv > best[k][0]:
                best[k] = (dv, row)
        for _, row in best.values():
            kname = row[hdr.index("Kernel Name")]
            rd, wr = val(hdr, units, row, "dram__bytes_read.sum"), val(hdr, units, row, "dram__bytes_write.sum")
            rec = {
                "kernel": kname, "workload": workload, "rep": os.path.basename(rep), "head_at_capture": head,
                "source": f"ncu --set full --clock-control none, one launch ({os.path.basename(rep)}), dram__bytes_read.sum + dram__bytes_write.sum",
                "dram_bytes_read": rd, "dram_bytes_write": wr,
                "dram_bytes_per_launch": (rd + wr) if rd is not None and wr is not None else None,
                "duration_ms_under_ncu": val(hdr, units, row, "gpu__time_duration.sum"),
                "tensor_pipe_active_pct": val(hdr, units, row, "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", False),
                "imma_pipe_active_pct": val(hdr, units, row, "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active", False),
                "imma_cycles_active_realtime_avg": val(hdr, units, row, "sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg", False),
                "fp64_pipe_active_pct": val(hdr, units, row, "sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", False),
                "fp64_inst_pct": val(hdr, units, row, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", False),
                "dfma_per_cycle": val(hdr, units, row, "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", False),
                "dadd_per_cycle": val(hdr, units, row, "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed", False),
                "dmul_per_cycle": val(hdr, units, row, "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed", False),
                "warp_inst_executed": val(hdr, units, row, "smsp__inst_executed.sum", False),
                "l2_hit_rate_pct": val(hdr, units, row, "lts__t_sector_hit_rate.pct", False),
                "issue_active_pct": val(hdr, units, row, "sm__issue_active.avg.pct_of_peak_sustained_elapsed", False),
                "dram_throughput_pct": val(hdr, units, row, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False),
                "sm_cycles_elapsed_avg": val(hdr, units, row, "sm__cycles_elapsed.avg", False),
                "registers_per_thread": val(hdr, units, row, "launch__registers_per_thread", False),
                "grid_size": val(hdr, units, row, "launch__grid_size", False),
            }
            cyc = rec["sm_cycles_elapsed_avg"]
            if cyc and None not in (rec["dfma_per_cycle"], rec["dadd_per_cycle"], rec["dmul_per_cycle"]):
                # FP64 flops the launch EXECUTED (thread-level DADD + DMUL + 2 DFMA; DMNMX / DSETP compares are not flops)
                rec["executed_fp64_flop"] = (rec["dadd_per_cycle"] + rec["dmul_per_cycle"] + 2.0 * rec["dfma_per_cycle"]) * cyc
            # every tensor-related counter the report holds, for the record
            rec["tensor_counters"] = {h: row[i] for i, h in enumerate(hdr) if ("tensor" in h or "imma" in h) and "peak_sustained" not in h.split(".")[-1]
                                      and row[i] not in ("0", "", "n/a")}
            db["kernels"] = [r for r in db["kernels"] if not (r["kernel"] == kname and r["workload"] == workload)] + [rec]
            print(kname[:70], workload, "dram", rec["dram_bytes_per_launch"], "tensor%", rec["tensor_pipe_active_pct"], "ms", rec["duration_ms_under_ncu"])
    json.dump(db, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
