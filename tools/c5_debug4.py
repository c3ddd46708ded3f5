"""Ground truth from the reference-order kernel (mode 0: AABB pruning, no overlay, no hints) for one C5 part; the
queries mode 1 gets wrong are then re-run alone (no first bound) and in small groups."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from axom_b200 import SignedDistance, synth
from axom_b200 import dist as D
freq = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000_000
part = int(sys.argv[3]) if len(sys.argv) > 3 else 7
dev = torch.device("cuda", 0)
x, y, z, conn = synth.icosphere(freq)
P = np.stack([x, y, z], 1)
cen = (P[conn[:, 0]] + P[conn[:, 1]] + P[conn[:, 2]]) / 3.0
parts = D.morton_partition(cen, 8)
qd = bench._points_device(nq, 999, -1.0, 1.0, dev)
c = conn[parts[part]]
sd = SignedDistance(x, y, z, c, 3, False, False, device=0)
sd.setMode(0)
truth = sd.computeDistances(qd)[0]
sd.setMode(1)
got = sd.computeDistances(qd)[0]
bad = torch.nonzero(got != truth).reshape(-1)
print("mode 1 != mode 0:", int(bad.numel()), "of", nq, "finite-but-wrong:", int((got[bad] < 1e100).sum()), flush=True)
got2 = sd.computeDistances(qd)[0]
print("second run differs from first in", int((got2 != got).sum()), "queries; wrong in second run:", int((got2 != truth).sum()), flush=True)
alone_bad = 0
for i in bad[:40].tolist():
    one = sd.computeDistances(qd[i:i + 1].contiguous())[0]
    alone_bad += int(one[0] != truth[i])
print("wrong when evaluated alone (no first bound):", alone_bad, "of", min(40, int(bad.numel())), flush=True)
grp = qd[bad].contiguous()
g = sd.computeDistances(grp)[0]
print("wrong when the failing queries are evaluated together (%d queries):" % grp.shape[0], int((g != truth[bad]).sum()), flush=True)
os.environ["AXB_SD_KERNEL"] = "fast"
sd2 = SignedDistance(x, y, z, c, 3, False, False, device=0)
f = sd2.computeDistances(qd)[0]
print("sd_fast_kernel (reference visiting order, double bounds) != mode 0:", int((f != truth).sum()), flush=True)
