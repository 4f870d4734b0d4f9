#!/usr/bin/env python
"""Headline benchmark: acquisition evaluations per second on BASELINE.json config 3
(ZDT1 30-D, 2 objectives, qNEHVI, N=2000, q=4, 512 MC samples, 16384 raw-sample q-batches).

    python bench.py --gpus 1 --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W  # CPU float64 port of the reference path (oracle/)
    torchrun ... bench.py --gpus N ...                     # one rank per GPU, weak scaling over q-batches

One "step" = one raw-sample screen: AcquisitionFunction.forward over `raw_samples` q-batches resident in
HBM (`value`) or handed over as HOST buffers through bo_acqf_forward_host (`e2e`).  1 eval = one q-batch
scored with all S MC samples.  Prints ONE JSON line (rank 0).
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


def workload_config(p, n_gpus, extra=None):
    cfg = {
        "workload": p["name"], "N_train": int(p["X"].shape[0]), "d": int(p["d"]), "outputs": len(p["outputs"]),
        "q": int(p["q"]), "mc_samples": int(p["S"]), "raw_samples_per_gpu": int(p["raw_samples"]),
        "global_q_batches_per_step": int(p["raw_samples"]) * n_gpus, "acqf": p["acqf"],
        "parallelism": f"q-batches sharded over {n_gpus} GPU(s), model state replicated",
        "eval_definition": "1 eval = one q-batch scored with all MC samples",
    }
    if extra:
        cfg.update(extra)
    return cfg


# ----------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port of the BoTorch op sequence, all host threads)
# ----------------------------------------------------------------------------------------------------
def cpu_reference_setup(p, baseline_idx=None, cell_bounds=None):
    from tests import problems as P
    from oracle import bo_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    gp = P.oracle_gp(p)
    if p["acqf"] == "qnehvi":
        if baseline_idx is None:
            acq = P.oracle_acqf(p, gp, prune_baseline=True)
        else:  # reuse the pruned baseline found by the device path (same points; saves CPU set-up time)
            acq = O.QNEHVIOracle(gp, p["ref_point"], torch.as_tensor(p["X"])[baseline_idx],
                                 [P.op_to_oracle(o) for o in p["objective"].ops], mc_samples=p["S"],
                                 seed=p["sampler_seed"], prune_baseline=False, cell_bounds=cell_bounds)
    else:
        acq = P.oracle_acqf(p, gp)
    return acq


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
    sample = (f"{n} of {p['raw_samples']} q-batches per step, scored in chunks of 8 (BoFire batch_limit=num_restarts, "
              f"botorch.py:108-128) with the per-call recompute of the n_b={getattr(acq, 'nb', 0)} baseline rows; "
              f"set-up {t_setup:.1f} s not timed")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(p, 1, {"cpu_sample_q_batches": n}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------
def measure_fp64_peak(device):
    """cuBLAS DGEMM 8192^3 through torch.matmul, best of 5 (CUDA events): the fp64 roofline denominator
    (MEASURED_PEAKS.json only holds HBM and bf16 figures)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.double, device=device)
    b = torch.randn(n, n, dtype=torch.double, device=device)
    torch.matmul(a, b)
    torch.cuda.synchronize(device)
    best = float("inf")
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0 * n**3 / (best * 1e-3) / 1e12


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

    p = make_problem(args)
    p["cand_seed"] = p["cand_seed"] + rank  # every rank screens its own raw samples (weak scaling)
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

    def step():
        vals = acq(X)
        if world > 1:  # the only exchange of the path: best (value, global index) per rank, 16 bytes
            i = torch.argmax(vals)
            pair = torch.stack([vals[i], (i + rank * b).to(torch.double)])
            allp = [torch.empty_like(pair) for _ in range(world)]
            dist.all_gather(allp, pair)
        return vals

    for _ in range(max(args.warmup, 3)):
        vals = step()
    barrier()
    launches0 = st.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        vals = step()
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.double, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    launches = st.launch_count() - launches0
    ms_total = float(ms[0])
    value = world * b * args.steps / (ms_total * 1e-3)

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

    # ---- roofline of the dominant kernel (posterior GEMM), timed with CUDA events on its stream ----
    roofline = None
    kernel_share = None
    if rank == 0:
        st.set_timing(True)
        per = {}
        reps = 3
        for _ in range(reps):
            acq(X)
            torch.cuda.synchronize(device)
            for name in ("prep", "crosscov", "ozaki_slice", "posterior_gemm", "cond_root", "sample_gemm", "mc_acqf"):
                t, cnt = st.last_timing(name)
                per.setdefault(name, []).append((t, cnt))
        st.set_timing(False)
        avg = {k: (sum(t for t, _ in v) / len(v), v[0][1]) for k, v in per.items()}
        total_ms = sum(t for t, _ in avg.values())
        kernel_share = {k: round(t / total_ms, 4) for k, (t, _) in avg.items()}
        g_ms, g_cnt = avg["posterior_gemm"]  # ms per step summed over its launches, launches per step
        N, nb, M = st.N, acq.nb, st.M
        rows = b * q
        # triangular solve-equivalent N^2 flops per point and output + the 1 + n_b dense extra rows
        flops_per_step = M * rows * (float(N) * N + 2.0 * N * (nb + 1))
        flops_per_launch = flops_per_step / g_cnt
        achieved = flops_per_step / (g_ms * 1e-3) / 1e12
        peak = measure_fp64_peak(device)
        # DRAM bytes of one launch from the committed ncu --set full capture (profiles/r01_ncu_summary.txt); only valid
        # for the default headline workload, null otherwise
        traffic = 2.349e9 if (args.workload == "zdt1" and not args.raw_samples) else None  # profiles/r01_s2_ncu_gemm_dram_groups8.csv
        oz_ms, oz_cnt = st.last_timing("ozaki_slice")
        roofline_fp64 = None
        chk = st.debug_get("ozaki_check", capacity=16).cpu().tolist()
        int8_check = {"state": {0: "not run (problem below the INT8 threshold)", 1: "INT8 path with per-row guard",
                                -1: "guard flagged most rows -> FP64 DMMA kernel"}[int(chk[0])],
                      "q_batches_redone_in_fp64_last_step": int(chk[1]), "q_batches_last_step": int(chk[2]),
                      "q_batches_redone_since_prepare": int(chk[3]), "q_batches_since_prepare": int(chk[4]),
                      "guard": f"2 sqrt(G_ii) eps + N eps^2 <= {chk[6]:g} (k** - G_ii), eps = {chk[5]:g} x 2^-56 sqrt(N) sA max sB "
                               "(csrc/ozaki.cu); flagged q-batches are recomputed by the FP64 kernel"}
        roofline = {"bound": "tensor", "kernel": "posterior_gemm_tma_kernel (FP64 DMMA m8n8k4, TMA + mbarrier ring)",
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                    "traffic_unit": "bytes per launch, dram__bytes_read.sum + dram__bytes_write.sum (ncu, profiles/r01_s2_ncu_gemm_dram_groups8.csv; "
                                    "16.2 GB before the L2-sharing column groups, algorithmic 2.2 GB)",
                    "peak_source": "torch.matmul f64 8192^3 (cuBLAS DGEMM) best of 5 measured in this run; "
                                   "MEASURED_PEAKS.json holds no fp64 figure; tools/fp64_peak measured 37.0 TFLOP/s "
                                   "for raw DMMA and DFMA issue on this pool",
                    "algorithmic_flops_per_launch": flops_per_launch, "launch_ms": g_ms / g_cnt,
                    "launches_per_step": g_cnt, "step_time_share": kernel_share, "int8_self_check": int8_check}
        if oz_cnt > 0:
            # the GEMM ran as 28 exact INT8 digit-plane products on tcgen05 (csrc/ozaki.cu): the roofline that bounds it is
            # the INT8 tensor pipe = 2 x the dense bf16 rate of MEASURED_PEAKS.json (same pipe, half the operand width)
            import json as _json
            try:
                mp = _json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
                i8_peak, src = 2.0 * float(mp["bf16_tflops"]), "2 x bf16_tflops (burst) of MEASURED_PEAKS.json"
            except Exception:
                i8_peak, src = 4500.0, "nominal 4.5 POPS dense INT8 (MEASURED_PEAKS.json unavailable)"
            i8_ops = 28.0 * flops_per_step / (g_ms * 1e-3) / 1e12
            roofline.update({
                "kernel": "ozaki_gemm2p_kernel (tcgen05.mma.kind::i8 M128xN128xK32, TMEM accumulators, two passes over K: "
                          "28 exact INT8 digit-plane products = one FP64-accurate GEMM)",
                "achieved": i8_ops, "peak": i8_peak, "unit": "TFLOP/s", "frac": i8_ops / i8_peak,
                "peak_source": src + "; tools/i8_mma_probe: 4.6 POPS (N >= 128) / 3.07 POPS (N = 64, this kernel's tile) "
                                     "single-SM issue rate x 148 at 1.9 GHz",
                "algorithmic_ops_note": "achieved = 28 digit-plane products x algorithmic FP64 flops / kernel time (INT8 TOP/s)",
                "fp64_equivalent_tflops": achieved, "fp64_pipe_peak_tflops": peak, "fp64_equivalent_over_fp64_pipe": achieved / peak,
                # one ncu --set full capture of the kernel on this workload (profiles/r01_s3_ncu_summary.txt): 3.50 GB read +
                # 1.13 GB written per launch; algorithmic 1.9 GB (7 digit planes of K(X*,X) and LinvExt, both outputs) -- the
                # rest is the second pass over planes 4..6 and the write-back of the 128 KB-per-SM integer scratch slab
                "traffic": 4.63e9 if (args.workload == "zdt1" and not args.raw_samples) else None,
                "traffic_unit": "bytes per launch, dram__bytes_read.sum + dram__bytes_write.sum (ncu --set full, "
                                "profiles/r01_s3_ncu_summary.txt); algorithmic 1.9 GB; 0.8 TB/s = 12 % of HBM, not the bound"})

    # ---- CPU baseline: oracle port on the host cores, bounded sample ----------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            idx = acq.prune_idx.cpu() if getattr(acq, "prune_idx", None) is not None else None
            cells = None
            if p["acqf"] == "qnehvi" and len(p["ref_point"]) > 2:
                # the oracle's pure-Python Lacour decomposition of 512 fronts takes minutes: time its forward pass
                # on the device's cell list (padded with empty cells), which the parity tests pin at small sizes
                lo, up, nc = acq.cell_bounds()
                C = int(nc.max())
                ref = torch.tensor(p["ref_point"], dtype=torch.double)
                lo, up = lo[:, :C].clone(), up[:, :C].clone()
                pad = torch.arange(C).unsqueeze(0) >= nc.unsqueeze(1)
                lo[pad] = ref
                up[pad] = ref
                cells = (lo, up)
            acq_cpu = cpu_reference_setup(p, baseline_idx=idx, cell_bounds=cells)
            n = args.cpu_sample or ((2048 if cells is None else 16) if p["acqf"] == "qnehvi" else 4096)  # ~10 s of CPU work
            Xc = X_host[:n]
            cpu_time_forward(acq_cpu, Xc[:8], 8)
            v8, dt8 = cpu_time_forward(acq_cpu, Xc, 8)
            vall, dtall = cpu_time_forward(acq_cpu, Xc, n)
            ref_vals = acq_cpu.forward(Xc)
            err = float((vals[:n].cpu() - ref_vals).abs().max() / ref_vals.abs().max())
            cpu = {"value": max(v8, vall), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"{n} of {b} q-batches; as BoFire calls it (chunks of 8): {v8:.1f} evals/s in {dt8:.1f} s; "
                             f"best case (one call over the sample): {vall:.1f} evals/s in {dtall:.1f} s; value = the faster",
                   "max_rel_err_gpu_vs_cpu_on_sample": err}
        except Exception as exc:  # the baseline must never take the headline number down with it
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {exc!r}"}

    # ---- ask() latency: acqf construction + screen + L-BFGS-B refinement of the restarts (rank 0, N=1) ----
    ask = None
    if rank == 0 and world == 1 and p.get("bounds") is not None:
        from everest_b200 import optim

        torch.cuda.synchronize(device)
        torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))  # the CPU baseline above may have raised it
        t0 = time.perf_counter()
        acq2 = Cf.build_acqf(p, st)
        torch.cuda.synchronize(device)
        t1 = time.perf_counter()
        if os.environ.get("EVEREST_BENCH_PROFILE"):
            import cProfile, pstats
            pr = cProfile.Profile(); pr.enable()
            acq2 = Cf.build_acqf(p, st); torch.cuda.synchronize(device)
            pr.disable()
            pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(16)
        bnds = torch.as_tensor(p["bounds"])
        Xic, Yic, _, _ = optim.gen_batch_initial_conditions(acq2, bnds, p["q"], p["num_restarts"], p["raw_samples"], seed=0)
        torch.cuda.synchronize(device)
        t2 = time.perf_counter()
        maxit = 200
        _, Yref, info = optim.gen_candidates_scipy(Xic, acq2, bnds[0], bnds[1], options={"maxiter": maxit})
        torch.cuda.synchronize(device)
        t3 = time.perf_counter()
        # one forward+backward of the restarts alone (device time of the adjoint path, CUDA events)
        Xr = Xic.to(device)
        acq2.forward_backward(Xr)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(10):
            acq2.forward_backward(Xr)
        ev1.record()
        torch.cuda.synchronize(device)
        ask = {"acqf_build_s": t1 - t0, "screen_s": t2 - t1, "refine_s": t3 - t2, "refine_maxiter": maxit,
               "refine_iterations": info["nit"], "refine_acqf_evals": info["n_acqf_evals"],
               "forward_backward_ms": ev0.elapsed_time(ev1) / 10.0,
               "best_screened": float(Yic.max()), "best_refined": float(torch.maximum(Yref, Yic).max()),
               "total_s": t3 - t0,
               "note": "refinement = scipy L-BFGS-B over all restarts, gradient from the analytic adjoint kernels "
                       "(bo_acqf_forward_backward)"}

    if world > 1 and p.get("bounds") is not None:
        # ask() over the GPUs of the box: raw samples and restarts sharded, one value all-gather + one arg-max exchange
        from everest_b200 import distributed as D

        barrier()
        t0 = time.perf_counter()
        acq2 = Cf.build_acqf(p, st)
        barrier()
        t1 = time.perf_counter()
        cand, val = D.sharded_optimize_acqf(acq2, torch.as_tensor(p["bounds"]), p["q"], p["num_restarts"],
                                            p["raw_samples"] * world, options={"maxiter": 200}, seed=0)
        barrier()
        t2 = time.perf_counter()
        ask = {"acqf_build_s": t1 - t0, "screen_and_refine_s": t2 - t1, "total_s": t2 - t0, "refine_maxiter": 200,
               "raw_samples_total": int(p["raw_samples"] * world), "best_refined": float(val),
               "note": "sharded_optimize_acqf: raw samples and restarts split over the ranks, analytic gradients, "
                       "one all-gather of the screen values + one (value, rank) arg-max exchange + one broadcast"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(p, world, {
                "n_baseline_after_pruning": int(acq.nb), "max_cells_per_sample": int(getattr(acq, "max_cells", 0)),
                "l2": "per-step working set (K(X*,X) 1.05 GB per output + factors) >> 126 MB L2; no flush needed",
                "qbatch_x_mc_samples_per_sec": value * p["S"], "setup_s": {"factorize": t_factor, "acqf_prepare": t_prepare}, "ask_latency": ask}),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(b * q * d * 8), "d2h_bytes_per_step": int(b * 8)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
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
