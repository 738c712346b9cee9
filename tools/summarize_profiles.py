"""Turn the ncu artefacts of a GPU round (gpurun_out/<tag>_launches.csv, <tag>_prof.ncu-rep) into the tracked summaries under
profiles/: per-kernel launch-time shares, the key counters of the full capture, and the DRAM traffic the bench quotes."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

tag = sys.argv[1]
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

# ---- launch list --------------------------------------------------------------------------------
rows = [l for l in open(os.path.join(G, tag + "_launches.csv")) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in csv.DictReader(rows):
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except Exception:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r["Metric Unit"], 1.0)
    k = re.sub(r"<.*", "", r["Kernel Name"]).replace("void ", "")
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
out = ["# ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --batch %d --steps 1 --warmup 0 (first %d launches)\n" % (
    batch, sum(v[0] for v in agg.values())),
    "# per-launch times are cold-cache and serialised: compare SHARES with the live per-phase CUDA-event shares of bench.py\n",
    "| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|\n"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("| %s | %d | %.2f | %.1f | %.1f %% |\n" % (k, v[0], v[1] / 1e3, v[1] / v[0], 100 * v[1] / tot))
open(os.path.join(P, tag + "_ncu_launches_b%d_summary.md" % batch), "w").write("".join(out))
print("".join(out))

# ---- full capture -------------------------------------------------------------------------------
rep = os.path.join(G, tag + "_prof.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(P, tag + "_ncu_full_b%d_raw_metrics.csv" % batch), "w").write(raw)
rr = list(csv.reader(raw.splitlines()))
hdr, units, data = rr[0], rr[1], rr[2:]
ix = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
lines = ["# ncu --set full --clock-control none, one tick with %d live quadrotor problems\n" % batch,
         "| kernel | ms | regs | warps/SMSP | FP64 pipe % | issue/cycle/SMSP | DRAM rd GB | DRAM wr GB | DRAM % | top stalls (cycles per issue) |\n|---|---|---|---|---|---|---|---|---|---|\n"]
traffic = {}
for d in data:
    t = float(d[ix["gpu__time_duration.sum"]])
    if t < 0.05:
        continue
    name = d[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    rd, wr = float(d[ix["dram__bytes_read.sum"]]), float(d[ix["dram__bytes_write.sum"]])
    st = sorted([(float(d[ix[s]]), s.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for s in stall], reverse=True)
    lines.append("| %s | %.3f | %s | %.2f | %.1f | %s | %.3f | %.3f | %.1f | %s |\n" % (
        name, t, d[ix["launch__registers_per_thread"]], float(d[ix["smsp__warps_active.avg.per_cycle_active"]]),
        float(d[ix["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]]), d[ix["smsp__issue_active.avg.per_cycle_active"]], rd, wr,
        float(d[ix["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]), ", ".join("%s %.2f" % (n, v) for v, n in st[:4])))
    if "ls_bp_kernel" in name:
        traffic = {"ls_bp_kernel_dram_bytes_per_problem_pass": (rd + wr) * 1e9 / batch, "ls_bp_kernel_ms_at_%d" % batch: t,
                   "source": "profiles/%s_ncu_full_b%d_raw_metrics.csv (dram__bytes_read.sum + dram__bytes_write.sum of one ls_bp_kernel launch "
                             "with %d live problems) / %d" % (tag, batch, batch, batch)}
open(os.path.join(P, tag + "_ncu_full_b%d_summary.md" % batch), "w").write("".join(lines))
print("".join(lines))
if traffic:
    # bench.py quotes this figure only while the kernel sources are the ones it was captured from (same hash as bench.kernel_source_hash)
    import hashlib
    h = hashlib.sha1()
    for f in ("lockstep.cuh", "engine.cuh", "models.cuh"):
        h.update(open(os.path.join(ROOT, "trajectoryoptimization.jl-c79d492b-0548-5874-b488-5a62c1d9d0ca_b200", "csrc", f), "rb").read())
    traffic["kernel_source_hash"] = h.hexdigest()[:16]
    json.dump(traffic, open(os.path.join(P, "dram_traffic.json"), "w"), indent=1)
    print(traffic)
