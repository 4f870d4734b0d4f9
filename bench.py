#!/usr/bin/env python
"""Headline benchmark: acquisition evaluations per second on BASELINE.json config 3
(ZDT1 30-D, 2 objectives, qNEHVI, N=2000, q=4, 512 MC samples, 16384 raw-sample q-batches), and ask() latency.

    python bench.py --gpus 1 --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W  # CPU float64 port of the reference path (oracle/)
    torchrun ... bench.py --gpus N ...                     # one rank per GPU: weak scaling (16384 q-batches per GPU) is the
                                                           # headline line, strong scaling (16384 in total) rides along

One "step" = one raw-sample screen: AcquisitionFunction.forward over `raw_samples` q-batches resident in HBM (`value`) or
handed over as HOST buffers through bo_acqf_forward_host (`e2e`).  1 eval = one q-batch scored with all S MC samples.
Prints ONE JSON line (rank 0).  Keys the driver keeps are flat scalars inside `config`, `roofline`, `cpu_baseline`, `e2e`;
the nested objects (`ask_latency`, `setup_s`, `roofline.step_time_share`, ...) repeat the detail for a reader of the line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "acqf_evals_per_sec"
UNIT = "evals/s"
ASK_MAXITER = 200


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="zdt1", choices=["zdt1", "dtlz2", "himmelblau", "mixed"])
    ap.add_argument("--raw-samples", type=int, default=None, help="override the q-batches per step (per GPU)")
    ap.add_argument("--cpu-sample", type=int, default=None, help="q-batches per CPU baseline measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ask", action="store_true", help="skip the ask() latency measurement")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="which of the two multi-GPU measurements becomes `value` (the other one is reported beside it)")
    return ap.parse_args()


def make_problem(args):
    from everest_b200 import configs as Cf

    kw = {}
    if args.raw_samples:
        kw["raw"] = args.raw_samples
    if args.workload == "zdt1":
        return Cf.zdt1_qnehvi(**kw)
    if args.workload == "dtlz2":
        return Cf.dtlz2_qnehvi(**kw)
    if args.workload == "himmelblau":
        return Cf.himmelblau_qlogei(**kw)
    if args.raw_samples:
        kw = {"n_choices": args.raw_samples}
    return Cf.mixed_tanimoto_qlogei(**kw)


def workload_config(p, n_gpus):
    """The SAME keys and values in both arms (the driver compares them): only what defines the workload."""
    return {
        "workload": p["name"], "N_train": int(p["X"].shape[0]), "d": int(p["d"]), "outputs": len(p["outputs"]),
        "q": int(p["q"]), "mc_samples": int(p["S"]), "raw_samples_per_gpu": int(p["raw_samples"]),
        "global_q_batches_per_step": int(p["raw_samples"]) * n_gpus, "acqf": p["acqf"],
        "parallelism": f"q-batches sharded over {n_gpus} GPU(s), model state replicated",
        "eval_definition": "1 eval = one q-batch scored with all MC samples",
        "l2": "per-step working set (digit planes / K(X*,X) of the q-batches, >= 1 GB) >> 126 MB L2; no flush needed",
    }


# ----------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port of the BoTorch op sequence, all host threads)
# ----------------------------------------------------------------------------------------------------
def cpu_reference_setup(p, acq_d=None):
    """Oracle acquisition function.  With the device acquisition function at hand its pruned baseline (and, for more
    than two objectives, its cell lists) are reused: same points, minutes of pure-Python set-up saved."""
    from tests import problems as P

    torch.set_num_threads(os.cpu_count() or 1)
    gp = P.oracle_gp(p)
    if p["acqf"] == "qnehvi" and acq_d is not None and getattr(acq_d, "prune_idx", None) is not None:
        return P.oracle_qnehvi_on_device_baseline(p, gp, acq_d, inject_cells=len(p["ref_point"]) > 2)
    return P.oracle_acqf(p, gp)


def cpu_time_forward(acq, X, chunk):
    """evals/s of the oracle on X[n, q, d], scored `chunk` q-batches per call like BoFire's batch_limit."""
    t0 = time.perf_counter()
    if hasattr(acq, "forward_reference_style"):
        acq.forward_reference_style(X, batch_limit=chunk)
    else:
        for i in range(0, X.shape[0], chunk):
            acq.forward(X[i:i + chunk])
    dt = time.perf_counter() - t0
    return X.shape[0] / dt, dt


def cpu_ask_latency(p, acq_o, screen_sample=512, maxiter=ASK_MAXITER):
    """ask() on the CPU port: acqf build is timed by the caller; the raw-sample screen is timed on `screen_sample`
    q-batches in chunks of 8 (as BoFire drives BoTorch) and extrapolated to raw_samples; the refinement is the SAME
    gen_candidates_scipy code as the GPU arm, with the gradient from torch autograd through the oracle (what BoTorch does)."""
    from everest_b200 import configs as Cf
    from everest_b200 import optim
    from tests import problems as P

    if p.get("bounds") is None:
        return None
    bounds = torch.as_tensor(p["bounds"])
    X = Cf.candidates(p, max(screen_sample, 8))[:screen_sample]
    t0 = time.perf_counter()
    vals = torch.cat([acq_o.forward(X[i:i + 8]) for i in range(0, X.shape[0], 8)])
    t_sample = time.perf_counter() - t0
    screen_s = t_sample * p["raw_samples"] / screen_sample
    X_ic, _ = optim.initialize_q_batch(X, vals, n=p["num_restarts"])
    adapter = P.OracleAcqfAdapter(acq_o, p["d"])
    t0 = time.perf_counter()
    _, Yref, info = optim.gen_candidates_scipy(X_ic, adapter, bounds[0], bounds[1], options={"maxiter": maxiter})
    refine_s = time.perf_counter() - t0
    return {"screen_s": screen_s, "screen_s_note": f"{screen_sample} q-batches timed ({t_sample:.2f} s), scaled to {p['raw_samples']}",
            "refine_s": refine_s, "refine_maxiter": maxiter, "refine_iterations": info["nit"],
            "refine_acqf_evals": info["n_acqf_evals"], "best_refined": float(Yref.max())}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from everest_b200 import configs as Cf

    p = make_problem(args)
    t_setup = time.perf_counter()
    acq = cpu_reference_setup(p)
    t_setup = time.perf_counter() - t_setup
    n = args.cpu_sample or (1024 if p["acqf"] == "qnehvi" else 2048)  # ~1.5 s of CPU work per step
    X = Cf.candidates(p, max(n, 8))[:n]
    for _ in range(max(args.warmup, 1)):
        cpu_time_forward(acq, X[: max(8, n // 4)], 8)
    times = []
    for _ in range(args.steps):
        _, dt = cpu_time_forward(acq, X, 8)
        times.append(dt)
    total = sum(times)
    value = n * args.steps / total
    cores = torch.get_num_threads()
    ask = None
    if not args.no_ask and len(p.get("ref_point", [0, 0])) <= 2:
        try:
            ask = cpu_ask_latency(p, acq)
            if ask is not None:
                ask["acqf_build_s"] = t_setup
                ask["total_s"] = t_setup + ask["screen_s"] + ask["refine_s"]
        except Exception as exc:  # the latency leg must not take the throughput number down with it
            ask = {"failed": repr(exc)}
    sample = (f"{n} of {p['raw_samples']} q-batches per step, scored in chunks of 8 (BoFire batch_limit=num_restarts, "
              f"botorch.py:108-128) with the per-call recompute of the n_b={getattr(acq, 'nb', 0)} baseline rows; "
              f"set-up {t_setup:.1f} s not timed")
    e2e = {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if ask and "total_s" in ask:
        e2e.update(ask_total_s=ask["total_s"], ask_build_s=ask["acqf_build_s"], ask_screen_s=ask["screen_s"], ask_refine_s=ask["refine_s"])
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(p, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "sample_q_batches": n},
        "e2e": e2e, "gpu_launches": 0, "ask_latency": ask, "setup_s": {"acqf_build": t_setup},
    }
    emit(line)


# ----------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    """`nvidia-smi` polled every 50 ms for the whole life of the process (it needs a few hundred ms to deliver its first line,
    which a 90 ms timed region does not wait for); `mark_begin` / `mark_end` bracket the timed region and `stop` reports the
    samples that fall inside it -- or, when the region slipped between two polls, the samples nearest to it (flagged)."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None
        self.t_begin = self.t_end = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.idx)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def stop(self):
        import datetime

        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        recs = []
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 10:
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    rec = (ts, float(f[2]), float(f[3]), float(f[4]))
                except ValueError:
                    continue
                rs = [name for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[6:10])
                      if v.lower().startswith("active")]
                recs.append(rec + (rs,))
            os.unlink(self.path)
        except Exception:
            pass
        if not recs:
            return out
        t0 = self.t_begin if self.t_begin is not None else recs[0][0]
        t1 = self.t_end if self.t_end is not None else recs[-1][0]
        inside = [r for r in recs if t0 - 0.02 <= r[0] <= t1 + 0.02]
        nearest = False
        if not inside:
            # the region fell between two polls: the two samples closest to it (the GPU is under the same load during the
            # warm-up steps right before and the end-to-end steps right after)
            mid = 0.5 * (t0 + t1)
            inside = sorted(recs, key=lambda r: abs(r[0] - mid))[:2]
            nearest = True
        sm = [r[1] for r in inside]
        out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(r[2] for r in inside),
                   reasons=sorted({x for r in inside for x in r[4]}), samples=len(inside), sm_mhz_min=min(sm),
                   power_w_max=max(r[3] for r in inside), timed_region_s=round(t1 - t0, 4),
                   samples_in_process=len(recs))
        if nearest:
            out["nearest_samples_only"] = True
            out["nearest_sample_offset_s"] = round(min(abs(r[0] - 0.5 * (t0 + t1)) for r in inside), 3)
        return out


# ----------------------------------------------------------------------------------------------------
# peaks measured in this run (MEASURED_PEAKS.json holds HBM and bf16 only)
# ----------------------------------------------------------------------------------------------------
def _best_of(fn, reps, device):
    fn()
    torch.cuda.synchronize(device)
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def measure_fp64_peak(device):
    """cuBLAS DGEMM 8192^3 through torch.matmul, best of 5 (CUDA events): the fp64 roofline denominator."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.double, device=device)
    b = torch.randn(n, n, dtype=torch.double, device=device)
    ms = _best_of(lambda: torch.matmul(a, b), 5, device)
    return 2.0 * n**3 / (ms * 1e-3) / 1e12


def measure_int8_peak(device):
    """cuBLASLt INT8 x INT8 -> INT32 GEMM 8192^3 through torch._int_mm: (burst TOP/s best of 10, sustained TOP/s over ~1 s
    back to back) -- the same protocol MEASURED_PEAKS.json uses for bf16.  None if this torch build has no such kernel."""
    try:
        n = 8192
        a = torch.randint(-128, 127, (n, n), dtype=torch.int8, device=device)
        b = torch.randint(-128, 127, (n, n), dtype=torch.int8, device=device)
        ms = _best_of(lambda: torch._int_mm(a, b), 10, device)
        burst = 2.0 * n**3 / (ms * 1e-3) / 1e12
        reps = max(10, int(1000.0 / ms))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch._int_mm(a, b)
        e1.record()
        e1.synchronize()
        sustained = 2.0 * n**3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
        return burst, sustained
    except Exception:
        return None


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


def committed_ncu(kernel_substr, workload):
    """Per-launch DRAM traffic and tensor-pipe counters of a kernel from the committed ncu summary of THIS round
    (profiles/r02_ncu_counters.json, written by tools/ncu_counters.py from a `ncu --set full` capture); None when the file
    holds nothing for this kernel on this workload -- never a constant typed into this script."""
    try:
        db = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_counters.json")))
    except Exception:
        return None
    for rec in db.get("kernels", []):
        if kernel_substr in rec.get("kernel", "") and rec.get("workload") == workload:
            return rec
    return None


# ----------------------------------------------------------------------------------------------------
# per-kernel rooflines: algorithmic work of one step / measured time of the kernel family in that step
# ----------------------------------------------------------------------------------------------------
def kernel_rooflines(p, st, acq, b, avg, peaks, fp64_peak, int8_peak, used_int8, workload):
    """avg: {family: (ms per step summed over its launches, launches per step)}.  Returns {family: roofline dict}."""
    N, nb, M, q, S = st.N, int(acq.nb), st.M, int(p["q"]), int(p["S"])
    rows = b * q
    ldk = ((N + 15) // 16) * 16
    hbm = float(peaks["hbm_gbs"]) if peaks else 6650.0
    hbm_src = "hbm_gbs of MEASURED_PEAKS.json (measured)" if peaks else "6650 GB/s (fallback of B200_PROFILING.md)"
    out = {}

    def entry(family, kernel, ncu_name, bound, work, unit_scale, unit, peak, peak_src, note):
        ms, cnt = avg.get(family, (0.0, 0))
        if cnt <= 0 or ms <= 0:
            return
        achieved = work / (ms * 1e-3) / unit_scale
        rec = committed_ncu(ncu_name, workload)
        out[family] = {"bound": bound, "kernel": kernel, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
                       "traffic": rec.get("dram_bytes_per_launch") if rec else None,
                       "traffic_source": (rec.get("source") if rec else "no ncu --set full capture of this kernel on this workload "
                                          "committed this round (profiles/r02_ncu_counters.json): null, not a typed-in constant"),
                       "peak_source": peak_src, "algorithmic_work_per_launch": work / cnt, "launch_ms": ms / cnt,
                       "launches_per_step": cnt, "algorithmic_work_note": note}
        if rec:
            for k in ("tensor_pipe_active_pct", "imma_pipe_active_pct", "fp64_pipe_active_pct", "issue_active_pct", "dram_throughput_pct"):
                if rec.get(k) is not None:
                    out[family]["ncu_" + k] = rec[k]

    # ---- posterior GEMM: V = K(X*,X) LinvExt^T never stored, Gram / W / mean reduced on the fly ----
    flops = M * rows * (float(N) * N + 2.0 * N * (nb + 1))   # triangular-solve equivalent N^2 per point and output + extras
    if used_int8:
        i8_peak = 2.0 * float(peaks["bf16_tflops"]) if peaks else 4500.0
        src = ("2 x bf16_tflops (burst) of MEASURED_PEAKS.json: INT8 runs on the same tensor pipe at half the operand width"
               if peaks else "nominal 4.5 POPS dense INT8 (MEASURED_PEAKS.json unavailable)")
        entry("posterior_gemm", "ozaki_gemm2p_kernel (tcgen05.mma.kind::i8 M128xN128xK32, TMEM accumulators, two passes over K: "
              "28 exact INT8 digit-plane products = one FP64-accurate GEMM)", "ozaki_gemm2p_kernel", "tensor", 28.0 * flops, 1e12,
              "TFLOP/s", i8_peak, src,
              "achieved = 28 digit-plane products x algorithmic FP64 flops M b q (N^2 + 2 N (1 + n_b)) / kernel time (INT8 TOP/s)")
        r = out.get("posterior_gemm")
        if r:
            r["fp64_equivalent_tflops"] = r["achieved"] / 28.0
            r["fp64_pipe_peak_tflops"] = fp64_peak
            r["fp64_equivalent_over_fp64_pipe"] = r["achieved"] / 28.0 / fp64_peak
            if int8_peak:
                r["int8_peak_measured_in_run_burst"] = int8_peak[0]
                r["int8_peak_measured_in_run_sustained"] = int8_peak[1]
                r["frac_vs_measured_int8"] = r["achieved"] / int8_peak[0]
                r["frac_vs_measured_int8_sustained"] = r["achieved"] / int8_peak[1]
                r["int8_peak_source"] = "torch._int_mm 8192^3 (cuBLASLt IGEMM) in this run: best of 10 / back to back for ~1 s"
            if peaks and peaks.get("bf16_tflops_sustained"):
                r["frac_vs_2x_bf16_sustained"] = r["achieved"] / (2.0 * float(peaks["bf16_tflops_sustained"]))
    else:
        entry("posterior_gemm", "posterior_gemm_tma_kernel (FP64 DMMA m8n8k4, TMA + mbarrier ring)", "posterior_gemm", "tensor", flops,
              1e12, "TFLOP/s", fp64_peak,
              "torch.matmul f64 8192^3 (cuBLAS DGEMM) best of 5 in this run; MEASURED_PEAKS.json has no fp64 figure",
              "achieved = algorithmic FP64 flops M b q (N^2 + 2 N (1 + n_b)) / kernel time")
    # ---- K(X*,X): writes 7 digit planes (INT8 path) or the FP64 matrix; one exp per element ----
    bytes_x = M * rows * float(ldk) * (7.0 if used_int8 else 8.0)
    from everest_b200 import kernels as Kn

    tani_words = 0
    for o in p["outputs"]:
        for lf in Kn.flatten(o["kernel"]).leaves:          # parameter-free Tanimoto leaves are shared: counted once
            if isinstance(lf, Kn.TanimotoKernel):
                tani_words += (len(lf.active_dims) + 63) // 64
    if tani_words:
        # fingerprint inner products <x, x'> = popc(x & x'): 2 x 32-bit POPC per 64-bit word and pair
        popc = 2.0 * rows * float(N) * tani_words
        clk = 1.965e9
        entry("crosscov", "crosscov_kernel2 (kernel trees: Tanimoto by AND + POPC over bit-packed fingerprints, Hamming on codes, "
              "continuous leaves on FP64 DMMA, kernel functions, sum of products)", "crosscov_kernel2", "alu", popc, 1e9, "GPOPC/s",
              148 * 16 * clk / 1e9, "derived, not measured: 16 POPC lanes per SM and clock (POPC issues at 1/8 of the warp rate, ncu "
              "profiles/r01_s3_ncu_summary.txt) x 148 SMs x 1.965 GHz", "algorithmic 32-bit POPCs = 2 b q N W (W = 64-bit words per fingerprint)")
        if "crosscov" in out:
            ms, cnt = avg["crosscov"]
            out["crosscov"]["hbm_write_gbs"] = bytes_x / (ms * 1e-3) / 1e9
            out["crosscov"]["hbm_write_frac"] = out["crosscov"]["hbm_write_gbs"] / hbm
    else:
        entry("crosscov", "crosscov_fast_kernel (K(X*,X): distances on FP64 DMMA, kernel function, digit slicing)",
              "crosscov_fast", "hbm", bytes_x, 1e9, "GB/s", hbm, hbm_src,
              "algorithmic bytes = M b q ldk x (7 digit planes | 8 B FP64) written once; reads (candidates, training rows) are L2-resident")
    # ---- MC value: samples -> objectives -> inclusion-exclusion over the cells ----
    if p["acqf"] == "qnehvi":
        Mo = len(p["ref_point"])
        try:
            cbar = float(st.debug_get("ncells", dtype=torch.int32).double().mean())
        except Exception:
            cbar = float(getattr(acq, "max_cells", 0))
        f_hvi = float(b) * S * cbar * (2 ** q - 1) * 3.0 * Mo + 2.0 * float(b) * S * M * q * (nb + q)
        entry("mc_acqf", "mc_hvi_tiled_kernel / mc_hvi_chunked_kernel (MC samples, objectives, inclusion-exclusion HVI)", "mc_hvi",
              "tensor", f_hvi, 1e12, "TFLOP/s", fp64_peak,
              "FP64 ALU (DFMA) rate = the FP64 tensor-pipe rate on this part (tools/fp64_peak: 37.0 TFLOP/s both); denominator = "
              "cuBLAS DGEMM measured in this run",
              f"algorithmic flops = b S C (2^q - 1) 3 M_o + 2 b S M q (n_b + q) (SURVEY.md 8d), C = mean cells per sample = {cbar:.1f}; "
              "the kernel skips cells and subsets that cannot overlap (their contribution is exactly 0), so it EXECUTES far fewer "
              "flops than this count: a fraction above 1 is work avoided, not pipe throughput -- ncu_fp64_pipe_active_pct (committed "
              "ncu capture) is how busy the FP64 pipe really is")
    else:
        f_sc = float(b) * S * q * (M * q + 60.0)
        entry("mc_acqf", "mc_scalar_kernel (MC samples, objective, log-space EI)", "mc_scalar", "tensor", f_sc, 1e12, "TFLOP/s",
              fp64_peak, "FP64 ALU: cuBLAS DGEMM measured in this run as the FP64 rate", "algorithmic flops = b S q (M q + ~60)")
    # ---- conditional root and sample GEMM ----
    entry("cond_root", "cond_root_kernel (sample_cached_cholesky: bl = Sqb Lb^-T, br = chol(Sqq - bl bl^T))", "cond_root", "tensor",
          float(b) * M * (q * float(nb) * nb + q ** 3 / 3.0 + 2.0 * q * q * nb), 1e12, "TFLOP/s", fp64_peak,
          "FP64 ALU: cuBLAS DGEMM measured in this run", "algorithmic flops = b M (q n_b^2 + q^3 / 3 + 2 q^2 n_b); latency-bound kernel")
    if nb > 0:
        entry("sample_gemm", "gemm_nt_kernel (baseline part of every MC sample, F' = bl z_b^T)", "gemm_nt", "hbm",
              M * rows * float(S) * 8.0, 1e9, "GB/s", hbm, hbm_src, "algorithmic bytes = M b q S x 8 B written (the operands are small)")
    return out


# ----------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist

    from everest_b200 import configs as Cf

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # nvidia-smi needs a few hundred ms to deliver its first line, which a 90 ms timed region does not wait for: the poller runs
    # from here on (set-up takes seconds) and the timed region is bracketed by mark_begin / mark_end
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    p = make_problem(args)
    p_global = dict(p)                                   # strong scaling: ONE candidate set, sliced over the ranks
    p["cand_seed"] = p["cand_seed"] + rank               # weak scaling: every rank screens its own raw samples
    t0 = time.perf_counter()
    st = Cf.build_state(p, device=device)
    torch.cuda.synchronize(device)
    t_factor = time.perf_counter() - t0
    t0 = time.perf_counter()
    acq = Cf.build_acqf(p, st)
    torch.cuda.synchronize(device)
    t_prepare = time.perf_counter() - t0
    X_host = Cf.candidates(p).contiguous()
    X = X_host.to(device)
    b, q, d = X.shape
    stream = torch.cuda.current_stream(device)
    pair_buf = torch.empty(2, dtype=torch.double, device=device)
    gather_buf = torch.empty(2 * world, dtype=torch.double, device=device) if world > 1 else None

    def exchange(vals, offset):
        # the only exchange of the path: best (value, global index) per rank, 16 bytes, one all-gather into a preallocated buffer
        v, i = torch.max(vals, dim=0)
        pair_buf[0] = v
        pair_buf[1] = (i + offset).to(torch.double)
        dist.all_gather_into_tensor(gather_buf, pair_buf)

    def timed(step_fn, steps, warm):
        for _ in range(warm):
            step_fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step_fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.double, device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0])

    def weak_step():
        vals = acq(X)
        if world > 1:
            exchange(vals, rank * b)
        return vals

    warm = max(args.warmup, 3)
    for _ in range(warm):
        vals = weak_step()
    barrier()
    launches0 = st.launch_count()
    sampler.mark_begin()
    ms_weak = timed(weak_step, args.steps, 0)
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    launches = st.launch_count() - launches0
    vals = acq(X)
    value_weak = world * b * args.steps / (ms_weak * 1e-3)

    # ---- strong scaling: BASELINE config 3 is 16384 raw samples IN TOTAL; every rank scores its 1/N slice ----
    strong = None
    if world > 1:
        from everest_b200 import distributed as D

        Xg = Cf.candidates(p_global)
        lo, hi = D.shard_bounds(Xg.shape[0], rank, world)
        Xs = Xg[lo:hi].contiguous().to(device)

        def strong_step():
            v = acq(Xs)
            exchange(v, lo)

        ms_strong = timed(strong_step, args.steps, warm)
        strong = {"value": Xg.shape[0] * args.steps / (ms_strong * 1e-3), "unit": UNIT, "ms_per_step": ms_strong / args.steps,
                  "q_batches_total": int(Xg.shape[0]), "q_batches_per_gpu": int(hi - lo),
                  "note": "fixed total work: the 16384 raw samples of the config split over the ranks, same 16-byte exchange"}

    # ---- end to end through the host-buffer C-ABI call (pinned staging, H2D, launches, D2H) -------
    Xh = X_host.numpy()
    for _ in range(2):
        out_h = acq.forward_host(Xh)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_h = acq.forward_host(Xh)
    torch.cuda.synchronize(device)
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.double, device=device)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * b * args.steps / float(t_e2e[0])
    # the host path pipelines chunks of q-batches (copy / compute overlap); the chunk size changes the Gram partial-sum
    # grouping and may select the other GEMM variant (FP64 DMMA below the size threshold, INT8 digit planes above), so the
    # two paths agree to rounding (amplified by the inclusion-exclusion sums for q = 8), not bit for bit
    ref_v = vals.cpu().numpy()
    assert np.allclose(out_h, ref_v, rtol=1e-8, atol=1e-10 * float(np.abs(ref_v).max())), \
        "host-buffer path disagrees with the device-pointer path"
    # wire format: fingerprint columns (read by Tanimoto leaves only) cross PCIe as bits; a caller that keeps its choice set
    # packed (the discrete branch re-scores the same set after every tell) skips the host pass over the float64 rows
    dense_cols, bit_cols = acq.pack_layout()
    packs = len(bit_cols) >= 256 and os.environ.get("EVEREST_HOST_PACK", "1") != "0"
    h2d_bytes = int(b * q * (len(dense_cols) + (len(bit_cols) + 63) // 64) * 8) if packs else int(b * q * d * 8)
    e2e_prepacked = None
    if packs:
        pk_dense, pk_bits = acq.pack_rows(Xh)
        for _ in range(2):
            out_p = acq.forward_host_packed(pk_dense, pk_bits, q=q)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out_p = acq.forward_host_packed(pk_dense, pk_bits, q=q)
        torch.cuda.synchronize(device)
        t_pk = torch.tensor([time.perf_counter() - t0], dtype=torch.double, device=device)
        if world > 1:
            dist.all_reduce(t_pk, op=dist.ReduceOp.MAX)
        e2e_prepacked = world * b * args.steps / float(t_pk[0])
        assert np.allclose(out_p, ref_v, rtol=1e-8, atol=1e-10 * float(np.abs(ref_v).max())), "packed host path disagrees"

    # ---- rooflines: every kernel family timed with CUDA events on its stream inside the step --------
    roofline = None
    if rank == 0:
        st.set_timing(True)
        per = {}
        fams = ("prep", "crosscov", "ozaki_slice", "posterior_gemm", "ozaki_guard", "ozaki_redo", "cond_root", "sample_gemm", "mc_acqf")
        for _ in range(3):
            acq(X)
            torch.cuda.synchronize(device)
            for name in fams:
                t, cnt = st.last_timing(name)
                per.setdefault(name, []).append((t, cnt))
        st.set_timing(False)
        avg = {k: (sum(t for t, _ in v) / len(v), v[0][1]) for k, v in per.items()}
        total_ms = sum(t for t, _ in avg.values())
        share = {k: round(t / total_ms, 4) for k, (t, _) in avg.items() if t > 0}
        chk = st.debug_get("ozaki_check", capacity=16).cpu().tolist()
        used_int8 = avg.get("ozaki_slice", (0, 0))[1] > 0 and int(chk[0]) == 1
        fp64_peak = measure_fp64_peak(device)
        int8_peak = measure_int8_peak(device) if used_int8 else None
        peaks = measured_peaks()
        rl = kernel_rooflines(p, st, acq, b, avg, peaks, fp64_peak, int8_peak, used_int8, p["name"])
        dominant = max((k for k in share if k in rl), key=lambda k: share[k], default=None)
        if dominant:
            roofline = dict(rl[dominant])
            roofline["dominant_kernel_family"] = dominant
            roofline["dominant_share_of_step"] = share[dominant]
            for k, v in share.items():
                roofline["share_" + k] = v                  # flat copies: the driver keeps scalars
            roofline["step_time_share"] = share
            roofline["other_kernels"] = {k: {kk: vv for kk, vv in v.items() if kk in ("bound", "achieved", "peak", "unit", "frac", "launch_ms")}
                                         for k, v in rl.items() if k != dominant}
            roofline["int8_guard"] = {
                "state": {0: "not run (problem below the INT8 threshold)", 1: "INT8 path with per-row guard",
                          -1: "guard flagged most rows -> FP64 DMMA kernel"}[int(chk[0])],
                "q_batches_redone_in_fp64_last_step": int(chk[1]), "q_batches_last_step": int(chk[2]),
                "q_batches_redone_since_prepare": int(chk[3]), "q_batches_since_prepare": int(chk[4]),
                "rule": f"2 sqrt(G_ii) eps + N eps^2 <= {chk[6]:g} (k** - G_ii), eps = {chk[5]:g} x 2^-56 sqrt(N) sA max sB "
                        "(csrc/ozaki.cu); flagged q-batches are recomputed by the FP64 kernel"}
            roofline["int8_q_batches_redone_in_fp64_per_step"] = int(chk[1])

    # ---- ask() latency: acqf construction + screen + refinement of the restarts -----------------------
    ask = None
    if rank == 0 and world == 1 and p.get("bounds") is not None and not args.no_ask:
        from everest_b200 import optim

        # measured BEFORE the CPU baseline: its thread pools keep the host cores busy for a while, and the refinement loop
        # (20 small launches per step) is sensitive to that
        bnds = torch.as_tensor(p["bounds"])
        first_call = None
        for rep in range(2):
            # the whole sequence twice: the first pass pays the one-time costs of a process (lazy loading of the adjoint
            # kernels, first allocation of their workspaces: 5 ... 500 ms, very noisy on this platform) and is reported as
            # `first_call`; the second pass is what every later ask() of a BO loop costs
            torch.cuda.synchronize(device)
            t0 = time.perf_counter()
            acq2 = Cf.build_acqf(p, st)
            torch.cuda.synchronize(device)
            t1 = time.perf_counter()
            Xic, Yic, _, _ = optim.gen_batch_initial_conditions(acq2, bnds, p["q"], p["num_restarts"], p["raw_samples"], seed=0)
            torch.cuda.synchronize(device)
            t2 = time.perf_counter()
            # refinement of the restarts: the on-device batched L-BFGS (bo_acqf_optimize) is the product path; the host-driven
            # scipy L-BFGS-B over the same device gradients (the round-1 path) is timed after it for comparison
            _, Yref, info = optim.gen_candidates_device(Xic, acq2, bnds[0], bnds[1], options={"maxiter": ASK_MAXITER})
            torch.cuda.synchronize(device)
            t3 = time.perf_counter()
            if rep == 0:
                first_call = {"acqf_build_s": t1 - t0, "screen_s": t2 - t1, "refine_s": t3 - t2, "total_s": t3 - t0}
        _, Yref_s, info_s = optim.gen_candidates_scipy(Xic, acq2, bnds[0], bnds[1], options={"maxiter": ASK_MAXITER})
        torch.cuda.synchronize(device)
        t3s = time.perf_counter()
        # one forward+backward of the restarts alone (device time of the adjoint path, CUDA events)
        Xr = Xic.to(device)
        acq2.forward_backward(Xr)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(10):
            acq2.forward_backward(Xr)
        ev1.record()
        torch.cuda.synchronize(device)
        ask = {"acqf_build_s": t1 - t0, "screen_s": t2 - t1, "refine_s": t3 - t2, "refine_maxiter": ASK_MAXITER,
               "refine_iterations": info["nit"], "refine_acqf_evals": info["n_acqf_evals"],
               "refine_optimizer": info.get("optimizer", "scipy L-BFGS-B, analytic device gradient"),
               "refine_converged_restarts": info.get("n_converged"), "refine_device_steps": info.get("n_steps"),
               "refine_s_host_scipy_lbfgsb": t3s - t3, "best_refined_host_scipy_lbfgsb": float(torch.maximum(Yref_s, Yic).max()),
               "forward_backward_ms": ev0.elapsed_time(ev1) / 10.0,
               "best_screened": float(Yic.max()), "best_refined": float(torch.maximum(Yref, Yic).max()),
               "total_s": t3 - t0, "first_call_in_process": first_call}

    # ---- CPU baseline: oracle port on the host cores, bounded sample ----------------------------------
    cpu = None
    acq_cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            t0 = time.perf_counter()
            acq_cpu = cpu_reference_setup(p, acq_d=acq)
            t_cpu_setup = time.perf_counter() - t0
            many = p["acqf"] == "qnehvi" and len(p["ref_point"]) > 2
            n = args.cpu_sample or ((16 if many else 2048) if p["acqf"] == "qnehvi" else 4096)  # ~10 s of CPU work
            Xc = X_host[:n]
            cpu_time_forward(acq_cpu, Xc[:8], 8)
            v8, dt8 = cpu_time_forward(acq_cpu, Xc, 8)
            vall, dtall = cpu_time_forward(acq_cpu, Xc, n)
            ref_vals = acq_cpu.forward(Xc)
            err = float((vals[:n].cpu() - ref_vals).abs().max() / ref_vals.abs().max())
            cpu = {"value": max(v8, vall), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"{n} of {b} q-batches; as BoFire calls it (chunks of 8): {v8:.1f} evals/s in {dt8:.1f} s; "
                             f"best case (one call over the sample): {vall:.1f} evals/s in {dtall:.1f} s; value = the faster",
                   "sample_q_batches": n, "max_rel_err_gpu_vs_cpu_on_sample": err, "setup_s": t_cpu_setup}
        except Exception as exc:  # the baseline must never take the headline number down with it
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {exc!r}"}

    if ask is not None and acq_cpu is not None and len(p.get("ref_point", [0, 0])) <= 2:
        try:
            torch.set_num_threads(os.cpu_count() or 1)
            cpu_ask = cpu_ask_latency(p, acq_cpu)
            cpu_ask["acqf_build_s"] = cpu["setup_s"] if cpu else None
            cpu_ask["total_s"] = (cpu_ask["acqf_build_s"] or 0.0) + cpu_ask["screen_s"] + cpu_ask["refine_s"]
            ask["cpu_port"] = cpu_ask
            ask["speedup_vs_cpu_port"] = cpu_ask["total_s"] / ask["total_s"]
        except Exception as exc:
            ask["cpu_port"] = {"failed": repr(exc)}

    if world > 1 and p.get("bounds") is not None and not args.no_ask:
        # ask() over the GPUs of the box, STRONG scaling: the config's raw samples and restarts sharded over the ranks
        from everest_b200 import distributed as D

        # measured twice like the single-GPU sequence: the first call of a process also pays for NCCL's lazy set-up of the
        # collectives it uses (all-gather of the screen values, broadcast of the winner) and is reported separately
        first = None
        for rep in range(2):
            barrier()
            t0 = time.perf_counter()
            acq2 = Cf.build_acqf(p_global, st)
            barrier()
            t1 = time.perf_counter()
            cand, val = D.sharded_optimize_acqf(acq2, torch.as_tensor(p["bounds"]), p["q"], p["num_restarts"],
                                                p["raw_samples"], options={"maxiter": ASK_MAXITER}, seed=0)
            barrier()
            t2 = time.perf_counter()
            if rep == 0:
                first = {"acqf_build_s": t1 - t0, "screen_and_refine_s": t2 - t1, "total_s": t2 - t0}
        ask = {"acqf_build_s": t1 - t0, "screen_and_refine_s": t2 - t1, "total_s": t2 - t0, "refine_maxiter": ASK_MAXITER,
               "first_call_in_process": first,
               "raw_samples_total": int(p["raw_samples"]), "best_refined": float(val),
               "note": "sharded_optimize_acqf, strong scaling: the config's raw samples and restarts split over the ranks, "
                       "one all-gather of the screen values + one (value, rank) arg-max exchange + one broadcast"}

    if rank == 0:
        use_strong = args.scaling == "strong" and strong is not None
        value = strong["value"] if use_strong else value_weak
        ms_step = strong["ms_per_step"] if use_strong else ms_weak / args.steps
        e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(b * 8),
               "host_input_bytes_per_step": int(b * q * d * 8)}
        if e2e_prepacked is not None:
            e2e["value_prepacked_host_buffers"] = e2e_prepacked
            e2e["wire_format"] = (f"{len(bit_cols)} fingerprint columns as bits + {len(dense_cols)} float64 columns per candidate; `value` packs "
                                  "the caller's float64 rows inside the timed call, `value_prepacked_host_buffers` starts from packed host buffers")
        if ask:
            e2e["ask_total_s"] = ask["total_s"]
            e2e["ask_build_s"] = ask["acqf_build_s"]
            for k_src, k_dst in (("screen_s", "ask_screen_s"), ("refine_s", "ask_refine_s"), ("screen_and_refine_s", "ask_screen_and_refine_s")):
                if k_src in ask:
                    e2e[k_dst] = ask[k_src]
            if isinstance(ask.get("cpu_port"), dict) and "total_s" in ask["cpu_port"]:
                e2e["ask_total_s_cpu_port"] = ask["cpu_port"]["total_s"]
        if strong:
            e2e["strong_scaling_value"] = strong["value"]
            e2e["strong_scaling_ms_per_step"] = strong["ms_per_step"]
            e2e["weak_scaling_value"] = value_weak
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if use_strong else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(p, world),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "weak_scaling": {"value": value_weak, "ms_per_step": ms_weak / args.steps, "q_batches_per_gpu": int(b)},
            "strong_scaling": strong,
            "setup_s": {"factorize": t_factor, "acqf_prepare": t_prepare},
            "ask_latency": ask,
            "details": {"n_baseline_after_pruning": int(acq.nb), "max_cells_per_sample": int(getattr(acq, "max_cells", 0)),
                        "qbatch_x_mc_samples_per_sec": value * p["S"]},
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def _quiet_stdout():
    """Library chatter on fd 1 (e.g. the 'NCCL version ...' banner) goes to stderr, so that stdout carries
    exactly ONE line: the JSON result."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse_args()
    _quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
