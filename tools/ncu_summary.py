#!/usr/bin/env python
"""Summarise ncu outputs: `launches <csv>` (per-kernel totals / shares) or `stalls <ncu-rep>` (stall samples per SASS opcode)
and `metrics <ncu-rep>` (a few headline counters)."""
import collections
import csv
import subprocess
import sys


def launches(path, last_n=None):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[h]
    ix = {k: j for j, k in enumerate(hdr)}
    recs = []
    for r in rows[h + 1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[ix["Metric Unit"]], 1.0)
        recs.append((r[ix["Kernel Name"]].split("(")[0], v))
    if last_n:
        recs = recs[-last_n:]
    tot, cnt = collections.Counter(), collections.Counter()
    for n, v in recs:
        tot[n] += v
        cnt[n] += 1
    T = sum(tot.values())
    print(f"{len(recs)} launches, {T/1e6:.3f} ms total")
    for k, v in tot.most_common(20):
        print(f"{k[:58]:58s} n={cnt[k]:5d} total={v/1e6:9.3f} ms share={100*v/T:5.1f}% avg={v/cnt[k]/1e3:10.1f} us")


def _ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def stalls(rep):
    rows = list(csv.reader(_ncu(rep, "source").splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

    def I(x):
        try:
            return int(x)
        except ValueError:
            return 0
    tot, ex = collections.Counter(), collections.Counter()
    per = collections.defaultdict(collections.Counter)
    n = 0
    for r in rows[2:]:
        if len(r) < len(hdr) or r[0] == "Address":
            continue
        toks = r[ix["Source"]].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        base = op.split(".")[0]
        if base in ("LDS", "LDGSTS", "STS", "BAR", "LDG", "STG", "SYNCS", "UTMALDG"):
            base = op
        s = I(r[ix["# Samples"]])
        n += s
        tot[base] += s
        ex[base] += I(r[ix["Instructions Executed"]])
        for st in names:
            per[base][st] += I(r[ix[st]])
    print("total samples", n)
    for op, s in tot.most_common(14):
        top = {k.replace("stall_", ""): v for k, v in per[op].items() if v > 0.04 * s}
        print(f"{op:24s} {100*s/n:5.1f}% exec {ex[op]:12d} {top}")


def metrics(rep):
    rows = list(csv.reader(_ncu(rep, "raw").splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "sm__pipe_tensor_cycles_active_realtime.avg.pct",
            "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "dram__bytes_read.sum.per_second", "lts__t_bytes.sum ", "sm__pipe_fp64_cycles_active_realtime.avg.pct",
            "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum "]
    for i, h in enumerate(hdr):
        if any((h + " ").startswith(w) or h == w.strip() for w in want):
            print(f"{h} [{units[i]}] = {' | '.join(r[i] for r in rows[2:])}")


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "launches":
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else None)
    elif cmd == "stalls":
        stalls(sys.argv[2])
    else:
        metrics(sys.argv[2])
