"""Where does a kernel wait?  Reads the SASS source page of an `ncu --set full --import-source on` report and prints, for the
first launch whose name matches, the instruction mix, the stall reasons and the instructions that collect the most warp-stall
samples (with executed count and average active threads -- a hot instruction running with 2 of 32 threads is a divergence bug,
a hot one at the top of a loop after a load is a serialised-latency bug).  Found this round with it: the 17 dependent DRAM
loads of the Jacobian kernel's finite check, the NaN slow paths of diverged line-search rollouts.

    python tools/ncu_hotspots.py gpurun_out/<tag>_prof.ncu-rep ls_trial_kernel [top=25] [launch=0]
"""
import collections
import csv
import io
import re
import subprocess
import sys


def load(rep, kernel, launch=0):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    if not starts:
        raise SystemExit("no kernel matching %r in %s" % (kernel, rep))
    s0 = starts[min(launch, len(starts) - 1)]
    s1 = next((s for s in starts if s > s0), len(rows))
    hdr = rows[s0 + 1]
    data = [dict(zip(hdr, r)) for r in rows[s0 + 2:s1] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
    return rows[s0][1], hdr, data


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    name, hdr, data = load(rep, kernel, launch)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(d["# Samples"]) for d in data) or 1
    inst = sum(int(d["Instructions Executed"]) for d in data)
    print("%s\n%d SASS instructions, %d warp-instructions executed, %d stall samples" % (name[:140], len(data), inst, tot))
    ops, opn = collections.Counter(), collections.Counter()
    for d in data:
        op = re.sub(r"^@!?U?P\d+\s+", "", d["Source"].strip()).split()[0]
        ops[op] += int(d["# Samples"])
        opn[op] += int(d["Instructions Executed"])
    print("-- by opcode (samples, share, executed)")
    for op, c in ops.most_common(12):
        print("   %-26s %8d %5.1f%% %12d" % (op, c, 100.0 * c / tot, opn[op]))
    agg = sorted(((sum(int(d[h]) for d in data), h[6:]) for h in stalls), reverse=True)
    print("-- stall reasons: " + ", ".join("%s %.1f%%" % (h, 100.0 * v / tot) for v, h in agg[:7]))
    addr = {int(d["Address"], 16): i for i, d in enumerate(data)}
    print("-- hottest instructions (index, samples, share, executed, avg active threads, SASS, top stalls)")
    for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i]["# Samples"]))[:top]):
        d = data[i]
        s = d["Source"].strip()
        m = re.search(r"0x7f[0-9a-f]+", s)
        if m and int(m.group(0), 16) in addr:
            s = s.replace(m.group(0), "->%d" % addr[int(m.group(0), 16)])
        st = sorted(((int(d[h]), h[6:]) for h in stalls if int(d[h]) > 0), reverse=True)[:2]
        print("   %5d %7s %5.1f%% %10s %3s  %-64s %s" % (i, d["# Samples"], 100.0 * int(d["# Samples"]) / tot, d["Instructions Executed"],
                                                     d["Avg. Threads Executed"], s[:64], ", ".join("%s:%d" % (h, v) for v, h in st)))


if __name__ == "__main__":
    main()
