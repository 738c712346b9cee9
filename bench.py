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


def kernel_source_hash():
    """hash of the kernel sources: the ncu traffic figure in profiles/ is only used when it was captured from these sources"""
    import hashlib
    h = hashlib.sha1()
    pkg = os.path.join(ROOT, "trajectoryoptimization.jl-c79d492b-0548-5874-b488-5a62c1d9d0ca_b200", "csrc")
    for f in ("lockstep.cuh", "engine.cuh", "models.cuh"):
        h.update(open(os.path.join(pkg, f), "rb").read())
    return h.hexdigest()[:16]


class Arm:
    """one resident batch on this rank: pinned host buffers, the solver handle, device-resident and end-to-end steps"""

    def __init__(self, lib, torch, dist, api, prob, opts, x0_np, local_rank, world, B_global):
        self.lib, self.torch, self.dist = lib, torch, dist
        n, m, N = QUAD["n"], QUAD["m"], QUAD["N"]
        B = len(x0_np)
        self.B, self.world = B, world
        self.x0_h = torch.empty((B, n), dtype=torch.float64).pin_memory()
        self.U0_h = torch.empty((B, N - 1, m), dtype=torch.float64).pin_memory()
        self.x0_h.numpy()[:] = x0_np
        self.U0_h.numpy()[:] = prob.U[None]
        self.X_h = torch.empty((B, N, n), dtype=torch.float64).pin_memory()
        self.U_h = torch.empty((B, N - 1, m), dtype=torch.float64).pin_memory()
        self.dts_h = torch.empty((B, N - 1), dtype=torch.float64).pin_memory()
        self.res_h = torch.empty((B, 32), dtype=torch.uint8).pin_memory()
        self.bs = api.BatchSolver(prob, B, local_rank, 0, 0)
        _, self.copts = api.as_altro_options(opts)
        sp = C.c_void_p()
        lib.to_stream(self.bs.h, C.byref(sp))
        self.ext = torch.cuda.ExternalStream(sp.value, device=torch.device("cuda", local_rank))
        # allgather of the 32-byte result records: every rank contributes `per` records (ragged shards padded)
        self.per = (B_global + world - 1) // world if B_global else B
        self.gather_src = torch.zeros((self.per, 32), dtype=torch.uint8, device="cuda")
        self.gather_dst = torch.empty((world * self.per, 32), dtype=torch.uint8, device="cuda") if world > 1 else None
        self.check(lib.to_set_batch(self.bs.h, self.x0_h.data_ptr(), self.U0_h.data_ptr(), None))

    def check(self, rc):
        if rc != 0:
            raise RuntimeError((self.lib.to_last_error(self.bs.h) or b"").decode())

    def resident_step(self):
        """inputs already in HBM; returns device ms (kernel events + allgather events)"""
        self.check(self.lib.to_solve_altro(self.bs.h, C.byref(self.copts)))
        ms = self.bs.kernel_ms()
        if self.dist is not None:
            torch = self.torch
            self.check(self.lib.to_copy_results_device(self.bs.h, self.gather_src.data_ptr()))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self.dist.all_gather_into_tensor(self.gather_dst, self.gather_src)
            e1.record()
            e1.synchronize()
            ms += e0.elapsed_time(e1)
        return ms

    def verify_gather(self, rank):
        """content check of the NCCL allgather: this rank's slice of the gathered buffer is its own result records, and
        every rank holds the same gathered buffer (checksum all-reduced with MIN and MAX)"""
        if self.dist is None:
            return True
        torch = self.torch
        mine = self.gather_dst[rank * self.per:(rank + 1) * self.per]
        ok = bool(torch.equal(mine, self.gather_src))
        chk = self.gather_dst.to(torch.int64).sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        self.dist.all_reduce(lo, op=self.dist.ReduceOp.MIN)
        self.dist.all_reduce(hi, op=self.dist.ReduceOp.MAX)
        ok = ok and int(lo) == int(hi)
        flag = torch.tensor([1 if ok else 0], device="cuda")
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN)
        return bool(int(flag))

    def e2e_step(self):
        """host buffers in, host buffers out, everything on the engine's stream between two events"""
        torch, lib, bs = self.torch, self.lib, self.bs
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.ext)
        self.check(lib.to_set_batch(bs.h, self.x0_h.data_ptr(), self.U0_h.data_ptr(), None))
        self.check(lib.to_solve_altro_async(bs.h, C.byref(self.copts)))
        self.check(lib.to_get_solution(bs.h, self.X_h.data_ptr(), self.U_h.data_ptr(), self.dts_h.data_ptr()))
        self.check(lib.to_get_results(bs.h, self.res_h.data_ptr()))
        e1.record(self.ext)
        e1.synchronize()
        return e0.elapsed_time(e1)

    def bytes_moved(self):
        h2d = self.x0_h.numel() * 8 + self.U0_h.numel() * 8
        d2h = self.X_h.numel() * 8 + self.U_h.numel() * 8 + self.dts_h.numel() * 8 + self.res_h.numel()
        return h2d, d2h

    def close(self):
        self.bs.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=65536, help="GLOBAL batch (strong scaling: split over the GPUs); with --scaling weak: per GPU")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--interleave", action="store_true", help="strong scaling: rank r takes problems r, r+N, ... instead of a contiguous slice")
    ap.add_argument("--no-weak-extra", action="store_true", help="N>1, strong: skip the extra weak-scaling measurement (full batch per GPU)")
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
    from trajopt_b200 import abi, api, problems, sharding
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
    strong = args.scaling == "strong"
    B_global = args.batch if strong else args.batch * world
    # synthetic inputs: problem b of the GLOBAL batch is seeded by b, whatever the number of GPUs
    if strong:
        idx = sharding.shard_indices(B_global, rank, world, args.interleave)
    else:
        idx = np.arange(rank * args.batch, (rank + 1) * args.batch)
    if len(idx) and idx[-1] - idx[0] + 1 == len(idx):
        x0_np = problems.batch_x0("quadrotor", len(idx), offset=int(idx[0]))
    else:
        x0_np = problems.batch_x0("quadrotor", B_global)[idx]
    B = len(idx)
    arm = Arm(lib, torch, dist, api, prob, opts, x0_np, local_rank, world, B_global if strong else 0)
    bs = arm.bs

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    lockstep = os.environ.get("TRAJOPT_B200_ENGINE", "lockstep")[0] not in "p0"
    log("inputs ready (B=%d on this GPU, %d global, %s scaling), engine=%s" % (B, B_global, args.scaling, "lockstep" if lockstep else "persistent"))
    for i in range(args.warmup):
        ms = arm.resident_step()
        log("warmup %d: %.1f ms, %d launches" % (i, ms, bs.launches()))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t_wall0 = time.perf_counter()
    step_ms, kernel_ms = [], []
    launches = 0
    for _ in range(args.steps):       # the timed region: no per-phase events on the stream
        ms = arm.resident_step()
        step_ms.append(ms)
        kernel_ms.append(bs.kernel_ms())
        launches += bs.launches()
        log("timed step: %.1f ms" % ms)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    gather_ok = arm.verify_gather(rank)
    res = bs.results()
    trials = C.c_int64()
    arm.check(lib.to_last_linesearch_trials(bs.h, C.byref(trials)))
    e2e_ms = [arm.e2e_step() for _ in range(max(1, min(2, args.steps)))]
    log("e2e steps: %s ms" % e2e_ms)
    clocks = sampler.stop() if rank == 0 else None
    # one EXTRA step with a CUDA event per phase kernel (roofline of the dominant kernel); not part of the timed region
    phase_ms = np.zeros(5)
    ticks_total, resident_ms, resident_problems, lockstep_passes, phase_step_ms = 0, 0.0, 0, 0, 0.0
    if lockstep and rank == 0:
        lib.to_debug_phase_timing(bs.h, 1)
        arm.check(lib.to_solve_altro(bs.h, C.byref(arm.copts)))
        phase_step_ms = bs.kernel_ms()
        pm = (C.c_double * 5)()
        lib.to_debug_phase_ms(bs.h, pm)
        phase_ms = np.array(list(pm))
        tk = C.c_int32()
        lib.to_debug_ticks(bs.h, C.byref(tk))
        ticks_total = tk.value
        rms, rpb, lsp = C.c_double(), C.c_int32(), C.c_int64()
        lib.to_debug_resident(bs.h, C.byref(rms), C.byref(rpb), C.byref(lsp))
        resident_ms, resident_problems, lockstep_passes = rms.value, rpb.value, lsp.value
        lib.to_debug_phase_timing(bs.h, 0)
        log("phase step: %.1f ms; phases %s; resident kernel %.1f ms (<= %d problems); lockstep passes %d" %
            (phase_step_ms, np.round(phase_ms, 1), resident_ms, resident_problems, lockstep_passes))
    barrier()

    total_ms = float(sum(step_ms))
    e2e_mean = float(np.mean(e2e_ms))
    rank_ms = [total_ms / args.steps]
    if dist is not None:
        mine = torch.tensor([total_ms / args.steps], dtype=torch.float64, device="cuda")
        allms = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allms, mine)
        rank_ms = [float(x) for x in allms]
        t = torch.tensor([total_ms, e2e_mean], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_mean = float(t[0]), float(t[1])
        cnt = torch.tensor([float(res["steps"].sum()), float(trials.value), float(launches), float(B)], dtype=torch.float64, device="cuda")
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        steps_all, trials_all, launches_all, B_all = float(cnt[0]), float(cnt[1]), float(cnt[2]), int(cnt[3])
    else:
        steps_all, trials_all, launches_all, B_all = float(res["steps"].sum()), float(trials.value), float(launches), B

    # ---- optional extra: the weak-scaling figure (full batch on every GPU), N > 1 strong runs only ----
    weak = None
    if strong and world > 1 and not args.no_weak_extra:
        h2d_s, d2h_s = arm.bytes_moved()
        arm.close()
        x0_w = problems.batch_x0("quadrotor", args.batch, offset=rank * args.batch)
        arm_w = Arm(lib, torch, dist, api, prob, opts, x0_w, local_rank, world, 0)
        arm_w.resident_step()
        barrier()
        wms = arm_w.resident_step()
        barrier()
        tw = torch.tensor([wms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        weak = {"value": world * args.batch / (float(tw[0]) * 1e-3), "unit": "solves/s", "ms_per_step": float(tw[0]),
                "batch_per_gpu": args.batch, "steps": 1, "warmup": 1, "note": "every rank solves its own full batch (round-1 headline definition)"}
        arm_w.close()
        arm = None
    else:
        h2d_s, d2h_s = arm.bytes_moved()

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = B_all / (ms_per_step * 1e-3)
        iters_per_s = steps_all / (ms_per_step * 1e-3)
        # ---- roofline ------------------------------------------------------------------------------------
        iters_rank = float(res["steps"].sum())          # iLQR iterations (= backward passes) of this rank, one step
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
        # DRAM traffic of the dominant kernel: only from an ncu capture of THESE kernel sources
        traffic_pp, traffic_note = None, "no ncu capture in profiles/"
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
            if prof.get("kernel_source_hash") == kernel_source_hash():
                traffic_pp, traffic_note = prof.get("ls_bp_kernel_dram_bytes_per_problem_pass"), prof.get("source", "")
            else:
                traffic_note = "profiles/dram_traffic.json was captured from other kernel sources (hash %s != %s): not used" % (
                    prof.get("kernel_source_hash"), kernel_source_hash())
        except Exception:
            pass
        whole = {"achieved_tflops": iters_rank * f_it / (kms * 1e-3) / 1e12, "achieved_gbs": iters_rank * b_it / (kms * 1e-3) / 1e9,
                 "algorithmic_flops_per_iter": f_it, "algorithmic_bytes_per_iter": b_it, "solve_ms": kms}
        if lockstep and ticks_total > 0 and phase_ms[1] > 0:
            # dominant kernel: the backward pass of the lockstep ticks (one launch per tick, one pass per live problem per launch);
            # the iterations the resident kernel served are not in this phase time and not in `passes`
            F_bp, B_bp = algorithmic_bp()
            launches_bp = ticks_total
            passes = float(lockstep_passes) if lockstep_passes > 0 else iters_rank
            bp_ms = float(phase_ms[1])
            ach_tf = passes * F_bp / (bp_ms * 1e-3) / 1e12
            ach_gbs = passes * B_bp / (bp_ms * 1e-3) / 1e9
            roofline = {
                "bound": "fp64", "achieved": ach_tf, "peak": fp64.value, "unit": "TFLOP/s", "frac": ach_tf / fp64.value if fp64.value > 0 else None,
                "traffic": (traffic_pp * passes / launches_bp) if traffic_pp else None, "traffic_note": traffic_note,
                "kernel": "backward-pass phase of the lockstep ticks: tob::ls_bp_kernel<Cfg<4,0,false,false,2>,4,3> (16 lanes per problem; ticks with "
                          "> 4,096 live problems) and tob::ls_expand_kernel + tob::ls_bp_cta_kernel<..,256,2> (CTA per problem; the other ticks); "
                          "%.0f%% of the device time of a step" % (100.0 * bp_ms / max(1e-9, phase_step_ms)),
                "peak_source": "measured register-resident DFMA probe (to_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry; "
                               "tensor cores do not apply (13x13 FP64 contractions)",
                "launches": int(launches_bp), "avg_launch_ms": bp_ms / launches_bp, "passes": passes,
                "algorithmic_flops_per_launch": passes * F_bp / launches_bp, "algorithmic_bytes_per_launch": passes * B_bp / launches_bp,
                "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak, "peak_source": hbm_src},
                "phase_ms_per_step": {k: float(v) for k, v in zip(("jac", "bp", "trial", "accept", "outer"), phase_ms)},
                "resident_kernel": {"ms": resident_ms, "problems_taken_over": resident_problems, "iterations": iters_rank - passes,
                                    "kernel": "tob::ls_resident_kernel (one CTA per problem, whole iLQR iterations in-kernel)"},
                "measured_on": "one extra step with a CUDA event per phase kernel (%.1f ms), outside the timed region" % phase_step_ms,
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
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "ilqr_iters_per_s": iters_per_s,
            "config": {"workload": "quadrotor ALTRO (AL phase of benchmark/quadrotor_benchmarks.jl, PN off), n=13 m=4 N=101 rk3, "
                                   "u>=0 + terminal box, per-problem random x0 (SURVEY 8d item 3)",
                       "global_batch": B_all, "batch_per_gpu": B,
                       "parallelism": "dp%d: ONE global batch of %d split over the GPUs (%s shards), no solve-path collective, "
                                      "one NCCL allgather of the 32-byte result records per step" % (
                                          world, B_all, "interleaved" if args.interleave else "contiguous") if strong else
                                      "dp%d: every rank solves its own batch (weak)" % world,
                       "engine": "lockstep + resident tail" if lockstep else "persistent",
                       "mean_iters_per_solve": steps_all / max(1, B_all), "mean_linesearch_trials": L,
                       "ticks_per_step": ticks_total},
            "e2e": {"value": B_all / (e2e_mean * 1e-3), "unit": "solves/s", "h2d_bytes_per_step": int(h2d_s), "d2h_bytes_per_step": int(d2h_s),
                    "ms_per_step": e2e_mean, "timing": "CUDA events on the engine stream around H2D(pinned)+solve+D2H, max over ranks; bytes are per rank"},
            "gpu_launches": int(launches_all),
            "roofline": roofline,
            "clocks": clocks,
            "wall_s_timed_region": t_wall,
            "rank_ms_per_step": rank_ms,
            "allgather_verified": gather_ok,
            "status_histogram": {str(int(k)): int(v) for k, v in zip(*np.unique(res["status"], return_counts=True))},
        }
        if weak is not None:
            out["weak_scaling"] = weak
        ws = C.c_uint64()
        g, sm = C.c_int32(), C.c_int32()
        if arm is not None and lib.to_debug_grid(bs.h, 0, C.byref(g), C.byref(sm), C.byref(ws)) == 0:
            nws = B if lockstep else g.value
            out["config"]["l2_policy"] = ("inputs (x0+U0 = %.0f MB per GPU) and the solver workspaces (%d x %.2f MB = %.0f MB) exceed the 126 MB L2; "
                                          "no flush needed between steps" % (h2d_s / 1e6, nws, ws.value * 8 / 1e6, nws * ws.value * 8 / 1e6))
        if not args.no_cpu_baseline and world == 1:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import oracle_py
            cores = host_cores()
            sample = args.cpu_sample or 48 * cores  # ~10 s of host work on the 16-core GPU box
            sample = min(sample, B)
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
    if arm is not None:
        arm.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
