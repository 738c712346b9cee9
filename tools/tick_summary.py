"""Summarise a TRAJOPT_B200_TICK_LOG file (with TRAJOPT_B200_TICK_DETAIL=1): per-phase ms for the first ticks
(bulk) and for the tail."""
import sys
import numpy as np

segs, cur = [], []
for l in open(sys.argv[1]):
    if l.startswith("#"):
        if cur:
            segs.append(cur)
        cur = []
    else:
        cur.append([float(x) for x in l.split()])
if cur:
    segs.append(cur)
s = np.array(segs[-1])
names = ["jac", "bp", "trial", "accept", "outer"] if s.shape[1] == 7 else ["tick"]
ph = s[:, 1:1 + len(names)]
act = s[:, -1]
def line(tag, sl):
    if len(ph[sl]) == 0:
        return
    print("%-14s n=%4d  " % (tag, len(ph[sl])) + "  ".join("%s %.3f" % (n, v) for n, v in zip(names, ph[sl].mean(axis=0))) +
          "  | tick %.3f ms  total %.1f ms" % (ph[sl].sum(axis=1).mean(), ph[sl].sum()))
line("ticks 2-40", slice(2, 40))
line("ticks 40-100", slice(40, 100))
line("ticks 100-200", slice(100, 200))
line("ticks 200-400", slice(200, 400))
line("ticks 400+", slice(400, None))
print("ticks %d total %.1f ms" % (len(s), ph.sum()))
