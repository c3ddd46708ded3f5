import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from axom_b200 import SignedDistance, synth
from axom_b200 import dist as D
dev = torch.device("cuda", 0)
x, y, z, conn = synth.icosphere(1000)
P = np.stack([x, y, z], 1)
cen = (P[conn[:, 0]] + P[conn[:, 1]] + P[conn[:, 2]]) / 3.0
parts = D.morton_partition(cen, 8)
qd = bench._points_device(50_000_000, 999, -1.0, 1.0, dev)
sd = SignedDistance(x, y, z, conn[parts[7]], 3, False, False, device=0)
sd.setProfiling(2)
for rep in range(2):
    got = sd.computeDistances(qd)[0]
    print("misses", int((got > 1e100).sum()), flush=True)
