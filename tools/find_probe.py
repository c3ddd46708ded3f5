"""Quick GPU probe of the find* path (not a test): phase times for C1-like and C3/C4-like inputs.
   python tools/find_probe.py [n_boxes] [n_queries]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from axom_b200 import BVH, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
q = int(sys.argv[2]) if len(sys.argv) > 2 else n
dev = torch.device("cuda", 0)
boxes = torch.from_numpy(synth.triangle_aabbs(n, seed=12345)).to(dev)
pts = torch.from_numpy(synth.random_points(q, seed=12346)).to(dev)
qb = torch.from_numpy(synth.triangle_aabbs(q, seed=54321, shift=(0.5 * n ** (-1 / 3),) * 3)).to(dev)
o, d = synth.random_rays(min(q, 2_000_000), seed=777, lo=0.0, hi=1.0)
rays = torch.from_numpy(np.concatenate([o, d], axis=1)).to(dev)
b = BVH(3)
b.initialize(boxes)
b.setProfiling(True)
for _ in range(5):
    b.initialize(boxes)
print("build n=%d: total %.3f ms " % (n, b.phase_ms("build.total")),
      {k: round(b.phase_ms("build." + k), 3) for k in ("bounds", "morton", "sort", "tree", "refit")}, flush=True)
only = sys.argv[3] if len(sys.argv) > 3 else ""
for name, fn, arg in (("points", b.findPoints, pts), ("boxes", b.findBoundingBoxes, qb), ("rays", b.findRays, rays)):
    if only and name != only:
        continue
    fn(arg)
    b.setProfiling(True)
    for _ in range(3):
        off, cnt, cand = fn(arg)
    nq = cnt.numel()
    tot = b.phase_ms("find.total")
    print("find %-6s q=%d cand=%d (%.2f/q): total %.3f ms -> %.1f Mq/s" % (name, nq, cand.numel(), cand.numel() / nq, tot, nq / tot / 1e3),
          {k: round(b.phase_ms("find." + k), 3) for k in ("sortq", "count", "scan", "fill")}, flush=True)
