"""Do exact first bounds (every query repeated 64 times in a row, so a lane's previous closest point IS the answer) lose
the closest leaf when computeSign is off (zero tie window)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from axom_b200 import SignedDistance, synth
from axom_b200 import dist as D
freq = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
dev = torch.device("cuda", 0)
x, y, z, conn = synth.icosphere(freq)
P = np.stack([x, y, z], 1)
cen = (P[conn[:, 0]] + P[conn[:, 1]] + P[conn[:, 2]]) / 3.0
parts = D.morton_partition(cen, 8)
base = bench._points_device(300_000, 5, -1.0, 1.0, dev)
rep = base.repeat_interleave(64, dim=0).contiguous()
for k, v in {"AXB_SD_HINT_SHIFT": "0", "AXB_SD_HEAVY": "1000000000", "AXB_SD_NO_SOLO": "1"}.items():
    os.environ[k] = v  # the previous path: own-previous hint only
for p in (5, 7):
    c = conn[parts[p]]
    for cs in (False, True):
        sd = SignedDistance(x, y, z, c, 3, False, cs, device=0)
        one = torch.cat([sd.computeDistances(base[i:i + 1000].contiguous())[0] for i in range(0, 20000, 1000)])  # < 4096 per call: unsorted, loose hints
        many = sd.computeDistances(rep)[0].reshape(-1, 64)
        spread = (many != many[:, :1]).any(dim=1)
        miss = (many > 1e100).any(dim=1)
        dif = (many[:20000, 0] != one)
        print("part", p, "computeSign", cs, "queries whose 64 copies disagree:", int(spread.sum()), "with a miss:", int(miss.sum()),
              "first copy != small-batch answer:", int(dif.sum()), flush=True)
        bad = torch.nonzero(miss).reshape(-1)[:5]
        for i in bad.tolist():
            print("   q", base[i].tolist(), "copies", sorted(set(many[i].tolist()))[:4])
