#!/usr/bin/env python
"""bench.py — headline benchmark of BASELINE.json: ALTRO solves/s (and iLQR iterations/s) on a
batch of quadrotor problems (n=13, m=4, N=101, rk3, u>=0 + terminal box; problems/quadrotor.jl with
the AL phase of benchmark/quadrotor_benchmarks.jl:12-34), one process per GPU.

    python bench.py --gpus N --steps K --warmup W            # the B200 engine
    python bench.py --impl reference --gpus N ...            # the reference algorithm on the host cores

A "step" is ONE batched solve of `--batch` problems per GPU (default 65,536 = BASELINE configs[2]).
Weak scaling: every rank solves its own `--batch` problems (seeded by rank), no collective on the
solve path, one NCCL allgather of the 32-byte result records per step.

value      = problems solved per second, inputs already resident in HBM (CUDA events on the engine's stream)
e2e        = same, through the C ABI from pinned HOST buffers: H2D of x0/U0, solve, D2H of X/U/results
roofline   = algorithmic HBM bytes of the solve kernel (SURVEY §8d formula) / its event-timed duration,
             against MEASURED_PEAKS.json; plus the FP64-pipe figure against a measured DFMA peak
cpu_baseline = the CPU oracle (faithful port of the reference algorithm; Julia is not installable
             here) on all host cores, bounded sample
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

QUAD = dict(n=13, m=4, N=101, p_stage=4, p_term=18, c_f=190.0)


def algorithmic_bp():
    """FLOPs and HBM bytes of ONE backward pass of one QUAD problem = one problem's share of one launch of the
    dominant kernel ls_bp_kernel (SURVEY.md §8d F_bp; bytes: A,B,X,U,lambda,mu read once, K,d written once)."""
    n, m, N, p, pN = QUAD["n"], QUAD["m"], QUAD["N"], QUAD["p_stage"], QUAD["p_term"]
    F_bp = (N - 1) * (4 * n**3 + 6 * n * n * m + 2 * n * m * m + 2 * n * n + 2 * n * m + m**3 / 3 + 4 * m**3 / 3 + 2 * m * m * n
                      + 2 * m * m + 6 * n * n * m + 4 * n * m * m + 6 * n * m + 5 * n * n + 3 * n + 2 * m * m + 4 * m)
    sum_p = (N - 1) * p + pN
    nbytes = 8 * ((N - 1) * n * (n + m) + N * n + (N - 1) * m + (N - 1) * (m * n + m) + 2 * sum_p)
    return F_bp, nbytes


def algorithmic_per_iter(L):
    """FLOPs and HBM bytes of one iLQR iteration of one QUAD problem (FMA = 2)."""
    n, m, N, p, pN, cf = QUAD["n"], QUAD["m"], QUAD["N"], QUAD["p_stage"], QUAD["p_term"], QUAD["c_f"]
    F_bp = (N - 1) * (4 * n**3 + 6 * n * n * m + 2 * n * m * m + 2 * n * n + 2 * n * m + m**3 / 3 + 4 * m**3 / 3 + 2 * m * m * n
                      + 2 * m * m + 6 * n * n * m + 4 * n * m * m + 6 * n * m + 5 * n * n + 3 * n + 2 * m * m + 4 * m)
    F_jac = (N - 1) * 3 * cf * (n + m + 2)
    F_ce = (N - 1) * (2 * n * n + 4 * n * m + 2 * m * m + (n * n + m * m + n * m + n + m)) \
        + (N - 1) * p * (2 * n * n + 2 * m * m + 2 * n * m + 2 * (n + m) + 3) + pN * (2 * n * n + 2 * n)
    F_fp = (N - 1) * (2 * m * n + 2 * m + n + 3 * cf + 7 * n + 2 * n * n + 2 * m * m + 4 * (n + m) + 5 * p)
    flops = F_jac + F_ce + F_bp + L * F_fp
    sum_p = (N - 1) * p + pN
    nbytes = 8 * (2 * (N - 1) * n * (n + m) + (1 + L) * (N - 1) * (m * n + m) + (2 + 2 * L) * (N * n + (N - 1) * m) + 3 * sum_p)
    return flops, nbytes


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


T0 = time.time()

# stdout carries exactly ONE line, the JSON record.  Native libraries do not know that (NCCL prints "NCCL version ..." on
# stdout when the process group initialises), so file descriptor 1 is pointed at stderr for the whole run and the record is
# written to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(record):
    os.write(_REAL_STDOUT, (json.dumps(record) + "\n").encode())


def log(msg):
    """progress on stderr (stdout carries only the JSON line)"""
    sys.stderr.write("[bench %7.1fs] %s\n" % (time.time() - T0, msg))
    sys.stderr.flush()


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def make_problem():
    import trajopt_b200 as to  # noqa: F401
    from trajopt_b200 import problems
    return problems.quadrotor(), problems.quadrotor_bench_options()


def run_reference(args, rank, world):
    """The reference arm: the reference's algorithm (CPU oracle port; Julia cannot be installed in
    this image) on all host cores, bounded sample per step."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    from trajopt_b200 import problems
    prob, opts = make_problem()
    cores = host_cores()
    sample = args.cpu_sample or 24 * cores  # ~8 s of host work per step on the 16-core GPU box
    x0 = problems.batch_x0("quadrotor", sample, offset=0)
    for _ in range(args.warmup):
        oracle_py.solve(prob, opts, x0=x0[:cores], B=cores, inner_cap=0, outer_cap=0, threads=cores)
    t0 = time.perf_counter()
    steps_total = 0
    for _ in range(args.steps):
        r = oracle_py.solve(prob, opts, x0=x0, B=sample, inner_cap=0, outer_cap=0, threads=cores)
        steps_total += int(r["results"]["steps"].sum())
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    out = {"impl": "reference", "metric": "ALTRO solves/s (batched quadrotor N=101)", "value": value, "unit": "solves/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "ilqr_iters_per_s": steps_total / dt,
           "config": {"workload": "quadrotor ALTRO (AL phase), n=13 m=4 N=101 rk3, u>=0 + terminal box; bounded CPU sample",
                      "batch_per_step": sample},
           "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": "port",
                            "sample": "%d problems/step x %d steps, %d std::threads, g++ -O2 -ffp-contract=off -mfma" % (sample, args.steps, cores)},
           "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=65536, help="problems per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import trajopt_b200 as to
    from trajopt_b200 import abi, api, problems
    lib = abi.load_library()  # raises if the CUDA extension is missing: no fallback
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (the engine has no CPU path)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))

    prob, opts = make_problem()
    B = args.batch
    n, m, N = QUAD["n"], QUAD["m"], QUAD["N"]
    # synthetic inputs (seeded per rank), in PINNED host memory for the e2e leg
    x0_np = problems.batch_x0("quadrotor", B, offset=rank * B)
    x0_h = torch.empty((B, n), dtype=torch.float64).pin_memory()
    U0_h = torch.empty((B, N - 1, m), dtype=torch.float64).pin_memory()
    x0_h.numpy()[:] = x0_np
    U0_h.numpy()[:] = prob.U[None]
    X_h = torch.empty((B, N, n), dtype=torch.float64).pin_memory()
    U_h = torch.empty((B, N - 1, m), dtype=torch.float64).pin_memory()
    dts_h = torch.empty((B, N - 1), dtype=torch.float64).pin_memory()
    res_h = torch.empty((B, 32), dtype=torch.uint8).pin_memory()

    bs = api.BatchSolver(prob, B, local_rank, 0, 0)
    mode, copts = api.as_altro_options(opts)
    stream_ptr = C.c_void_p()
    lib.to_stream(bs.h, C.byref(stream_ptr))
    ext = torch.cuda.ExternalStream(stream_ptr.value, device=torch.device("cuda", local_rank))
    gather_src = torch.empty((B, 32), dtype=torch.uint8, device="cuda")
    gather_dst = torch.empty((world * B, 32), dtype=torch.uint8, device="cuda") if world > 1 else None

    def check(rc):
        if rc != 0:
            raise RuntimeError((lib.to_last_error(bs.h) or b"").decode())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step():
        """inputs already in HBM; returns device ms (kernel events + allgather events)"""
        check(lib.to_solve_altro(bs.h, C.byref(copts)))
        ms = bs.kernel_ms()
        if dist is not None:
            check(lib.to_copy_results_device(bs.h, gather_src.data_ptr()))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dist.all_gather_into_tensor(gather_dst, gather_src)
            e1.record()
            e1.synchronize()
            ms += e0.elapsed_time(e1)
        return ms

    def e2e_step():
        """host buffers in, host buffers out, everything on the engine's stream between two events"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        check(lib.to_set_batch(bs.h, x0_h.data_ptr(), U0_h.data_ptr(), None))
        check(lib.to_solve_altro_async(bs.h, C.byref(copts)))
        check(lib.to_get_solution(bs.h, X_h.data_ptr(), U_h.data_ptr(), dts_h.data_ptr()))
        check(lib.to_get_results(bs.h, res_h.data_ptr()))
        e1.record(ext)
        e1.synchronize()
        return e0.elapsed_time(e1)

    log("inputs ready (B=%d per GPU), engine=%s" % (B, os.environ.get("TRAJOPT_B200_ENGINE", "lockstep")))
    check(lib.to_set_batch(bs.h, x0_h.data_ptr(), U0_h.data_ptr(), None))
    for i in range(args.warmup):
        ms = resident_step()
        log("warmup %d: %.1f ms, %d launches" % (i, ms, bs.launches()))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lockstep = os.environ.get("TRAJOPT_B200_ENGINE", "lockstep")[0] not in "p0"
    if lockstep:
        lib.to_debug_phase_timing(bs.h, 1)  # one CUDA event per phase kernel on the engine's stream
    phase_ms = np.zeros(5)
    ticks_total = 0
    barrier()
    t_wall0 = time.perf_counter()
    step_ms, kernel_ms = [], []
    for _ in range(args.steps):
        ms = resident_step()
        step_ms.append(ms)
        kernel_ms.append(bs.kernel_ms())
        if lockstep:
            pm = (C.c_double * 5)()
            lib.to_debug_phase_ms(bs.h, pm)
            phase_ms += np.array(list(pm))
            tk = C.c_int32()
            lib.to_debug_ticks(bs.h, C.byref(tk))
            ticks_total += tk.value
        log("timed step: %.1f ms" % ms)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = bs.launches() * args.steps
    res = bs.results()
    trials = C.c_int64()
    check(lib.to_last_linesearch_trials(bs.h, C.byref(trials)))
    e2e_ms = [e2e_step() for _ in range(max(1, min(2, args.steps)))]
    log("e2e steps: %s ms" % e2e_ms)
    clocks = sampler.stop() if rank == 0 else None
    barrier()

    total_ms = float(sum(step_ms))
    e2e_mean = float(np.mean(e2e_ms))
    if dist is not None:
        t = torch.tensor([total_ms, e2e_mean], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_mean = float(t[0]), float(t[1])
        cnt = torch.tensor([float(res["steps"].sum()), float(trials.value)], dtype=torch.float64, device="cuda")
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        steps_all, trials_all = float(cnt[0]), float(cnt[1])
    else:
        steps_all, trials_all = float(res["steps"].sum()), float(trials.value)

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = world * B / (ms_per_step * 1e-3)
        iters_per_s = steps_all / (ms_per_step * 1e-3)
        # ---- roofline ------------------------------------------------------------------------------------
        iters_rank = float(res["steps"].sum())          # iLQR iterations (= backward passes) of this rank, last step
        L = float(trials.value) / max(1.0, iters_rank)  # sequential-equivalent line-search trials per iteration
        f_it, b_it = algorithmic_per_iter(L)
        kms = float(np.mean(kernel_ms))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_src = "measured copy bandwidth (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        fp64 = C.c_double()
        lib.to_measure_fp64_peak(local_rank, C.byref(fp64))
        prof = {}
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "r01_dram_traffic.json")))
        except Exception:
            pass
        h2d = x0_h.numel() * 8 + U0_h.numel() * 8
        d2h = X_h.numel() * 8 + U_h.numel() * 8 + dts_h.numel() * 8 + res_h.numel()
        whole = {"achieved_tflops": iters_rank * f_it / (kms * 1e-3) / 1e12, "achieved_gbs": iters_rank * b_it / (kms * 1e-3) / 1e9,
                 "algorithmic_flops_per_iter": f_it, "algorithmic_bytes_per_iter": b_it, "solve_ms": kms}
        if lockstep and ticks_total > 0 and phase_ms[1] > 0:
            # dominant kernel: the backward pass (one launch per tick, one pass per live problem per launch)
            F_bp, B_bp = algorithmic_bp()
            launches_bp = ticks_total                       # over the timed steps
            passes = iters_rank * args.steps                # every step solves the same batch: same iteration counts
            bp_ms = float(phase_ms[1])
            ach_tf = passes * F_bp / (bp_ms * 1e-3) / 1e12
            ach_gbs = passes * B_bp / (bp_ms * 1e-3) / 1e9
            traffic_pp = prof.get("ls_bp_kernel_dram_bytes_per_problem_pass")
            roofline = {
                "bound": "fp64", "achieved": ach_tf, "peak": fp64.value, "unit": "TFLOP/s", "frac": ach_tf / fp64.value if fp64.value > 0 else None,
                "traffic": (traffic_pp * passes / launches_bp) if traffic_pp else None,
                "kernel": "backward-pass phase: tob::ls_bp_kernel<Cfg<4,0,false,false,2>,4,3> (16 lanes per problem; ticks with > 4,096 live "
                          "problems) and tob::ls_expand_kernel + tob::ls_bp_cta_kernel<..,256,2> (CTA per problem; the other ticks); "
                          "%.0f%% of the device time of a step" % (100.0 * bp_ms / max(1e-9, float(phase_ms.sum()))),
                "peak_source": "measured register-resident DFMA probe (to_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry; "
                               "tensor cores do not apply (13x13 FP64 contractions)",
                "launches": int(launches_bp), "avg_launch_ms": bp_ms / launches_bp,
                "algorithmic_flops_per_launch": passes * F_bp / launches_bp, "algorithmic_bytes_per_launch": passes * B_bp / launches_bp,
                "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak, "peak_source": hbm_src},
                "phase_ms_per_step": {k: float(v) / args.steps for k, v in zip(("jac", "bp", "trial", "accept", "outer"), phase_ms)},
                "whole_solve": whole,
            }
        else:
            roofline = {"bound": "fp64", "achieved": whole["achieved_tflops"], "peak": fp64.value, "unit": "TFLOP/s",
                        "frac": whole["achieved_tflops"] / fp64.value if fp64.value > 0 else None, "traffic": None,
                        "kernel": "tob::solve_kernel<Cfg<4,0,false,false,2>> (persistent engine: the whole solve is one kernel)",
                        "peak_source": "measured DFMA probe (to_measure_fp64_peak)",
                        "hbm": {"achieved": whole["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": whole["achieved_gbs"] / hbm_peak,
                                "peak_source": hbm_src}, "whole_solve": whole}
        out = {
            "metric": "ALTRO solves/s (batched quadrotor N=101)", "value": value, "unit": "solves/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "ilqr_iters_per_s": iters_per_s,
            "config": {"workload": "quadrotor ALTRO (AL phase of benchmark/quadrotor_benchmarks.jl, PN off), n=13 m=4 N=101 rk3, "
                                   "u>=0 + terminal box, per-problem random x0 (SURVEY 8d item 3)",
                       "batch_per_gpu": B, "global_batch": world * B, "parallelism": "dp%d (batch sharded, no solve-path collective)" % world,
                       "engine": "lockstep" if lockstep else "persistent",
                       "mean_iters_per_solve": steps_all / (world * B), "mean_linesearch_trials": L,
                       "ticks_per_step": ticks_total / max(1, args.steps)},
            "e2e": {"value": world * B / (e2e_mean * 1e-3), "unit": "solves/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_mean, "timing": "CUDA events on the engine stream around H2D(pinned)+solve+D2H"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "clocks": clocks,
            "wall_s_timed_region": t_wall,
            "status_histogram": {str(int(k)): int(v) for k, v in zip(*np.unique(res["status"], return_counts=True))},
        }
        ws = C.c_uint64()
        g, sm = C.c_int32(), C.c_int32()
        if lib.to_debug_grid(bs.h, 0, C.byref(g), C.byref(sm), C.byref(ws)) == 0:
            nws = B if lockstep else g.value
            out["config"]["l2_policy"] = ("inputs (x0+U0 = %.0f MB per GPU) and the solver workspaces (%d x %.2f MB = %.0f MB) exceed the 126 MB L2; "
                                          "no flush needed between steps" % ((x0_h.numel() + U0_h.numel()) * 8 / 1e6, nws, ws.value * 8 / 1e6,
                                                                           nws * ws.value * 8 / 1e6))
        if not args.no_cpu_baseline and world == 1:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import oracle_py
            cores = host_cores()
            sample = args.cpu_sample or 48 * cores  # ~10 s of host work on the 16-core GPU box
            log("cpu baseline: %d problems on %d threads" % (sample, cores))
            t0 = time.perf_counter()
            r = oracle_py.solve(prob, opts, x0=x0_np[:sample], B=sample, inner_cap=0, outer_cap=0, threads=cores)
            dt = time.perf_counter() - t0
            same = bool(np.array_equal(r["results"]["iterations_total"], res["iterations_total"][:sample]) and
                        np.array_equal(r["results"]["status"], res["status"][:sample]))
            out["cpu_baseline"] = {"value": sample / dt, "unit": "solves/s", "cores": cores, "kind": "port",
                                   "ilqr_iters_per_s": float(r["results"]["steps"].sum()) / dt,
                                   "sample": "first %d problems of the same batch, %d std::threads (one problem per thread), %.1f s; "
                                             "oracle = C++ port of the reference algorithm (Julia not installable here)" % (sample, cores, dt),
                                   "iteration_counts_match_gpu": same}
        emit(out)
    bs.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
