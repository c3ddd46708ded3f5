"""C5 parity bisect: one Morton part of an icosphere, unsigned distance, random queries in [-1,1]^3, GPU against the oracle;
prints the mismatching queries.  Env toggles: AXB_SD_HINT_SHIFT=0, AXB_SD_HEAVY=1000000, AXB_SD_NO_SOLO=1."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from axom_b200 import SignedDistance, synth
from axom_b200 import dist as D
from oracle import oracle as O
freq = int(sys.argv[1]) if len(sys.argv) > 1 else 224
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
x, y, z, conn = synth.icosphere(freq)
P = np.stack([x, y, z], 1)
cen = (P[conn[:, 0]] + P[conn[:, 1]] + P[conn[:, 2]]) / 3.0
parts = D.morton_partition(cen, 8)
q = synth.random_points(nq, seed=999, lo=-1.0, hi=1.0)
kind = "reference" if O.have_reference() else "port"
tot_bad = 0
for p in range(8):
    c = conn[parts[p]]
    sd = SignedDistance(x, y, z, c, 3, False, False, device=0)
    got = sd.computeDistances(torch.from_numpy(q).cuda())[0].cpu().numpy()
    ref = O.SignedDistance(x, y, z, c, 3, False, False, kind=kind)
    want, _, _ = ref.compute(q, nthreads=64)
    bad = np.nonzero(want != got)[0]
    tot_bad += len(bad)
    print("part", p, "triangles", len(c), "mismatches", len(bad), flush=True)
    for i in bad[:5]:
        print("   q", q[i].tolist(), "want %.17g got %.17g rel %.3g" % (want[i], got[i], (got[i] - want[i]) / want[i]))
print("total mismatches", tot_bad)
