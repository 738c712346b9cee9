"""Stall samples and executed warp-instructions of one kernel aggregated by CUDA SOURCE LINE.

The SASS page of an `ncu --set full --import-source on` report has no line column in CSV form, so the line of every SASS
instruction is taken from `nvdisasm --print-line-info` of the cubin inside the object file the kernel was built into (the
n-th instruction of the function there is the n-th row of ncu's SASS page).

    python tools/ncu_lines.py gpurun_out/<tag>_prof.ncu-rep <kernel regex> <object .o> <mangled-name substring> [top=40] [launch=0]
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_hotspots import load  # noqa: E402


def line_table(obj, mangled):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    out = subprocess.run(["nvdisasm", "--print-line-info", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    lines, cur, inside = [], ("?", 0), False
    for ln in out:
        if ln.startswith("//---") and ".text." in ln:
            inside = mangled in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:  # an annotation holds until the next one
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            lines.append(cur)
    return lines


def main():
    rep, kernel, obj, mangled = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    launch = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    name, hdr, data = load(rep, kernel, launch)
    lt = line_table(obj, mangled)
    if len(lt) != len(data):
        print("warning: %d SASS rows in the report, %d instructions in the cubin (line mapping may be off)" % (len(data), len(lt)))
    samples, execd, last = collections.Counter(), collections.Counter(), ("?", 0)
    for i, d in enumerate(data):
        key = lt[i] if i < len(lt) and lt[i] else last
        last = key
        samples[key] += int(d["# Samples"])
        execd[key] += int(d["Instructions Executed"])
    tot = sum(samples.values()) or 1
    tote = sum(execd.values()) or 1
    print("%s\n%d stall samples, %d warp-instructions" % (name[:120], tot, tote))
    print("-- by source line (file:line, samples, share, executed, share)")
    for key, c in samples.most_common(top):
        print("   %-22s %8d %5.1f%% %12d %5.1f%%" % ("%s:%d" % key, c, 100.0 * c / tot, execd[key], 100.0 * execd[key] / tote))


if __name__ == "__main__":
    main()
