"""C5 at full size: new path vs the previous path (no hint table, no heavy cutoff, no cooperative kernel) on every part,
then the oracle on the queries where they differ."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from axom_b200 import SignedDistance, synth
from axom_b200 import dist as D
from oracle import oracle as O
freq = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000_000
dev = torch.device("cuda", 0)
x, y, z, conn = synth.icosphere(freq)
P = np.stack([x, y, z], 1)
cen = (P[conn[:, 0]] + P[conn[:, 1]] + P[conn[:, 2]]) / 3.0
parts = D.morton_partition(cen, 8)
qd = bench._points_device(nq, 999, -1.0, 1.0, dev)
kind = "reference" if O.have_reference() else "port"
TOGGLES = {"AXB_SD_HINT_SHIFT": "0", "AXB_SD_HEAVY": "1000000000", "AXB_SD_NO_SOLO": "1"}
for p in range(8):
    c = conn[parts[p]]
    sd = SignedDistance(x, y, z, c, 3, False, False, device=0)
    new = sd.computeDistances(qd)[0]
    res = {}
    for name in list(TOGGLES) + ["all"]:
        for k, v in TOGGLES.items():
            if name in (k, "all"):
                os.environ[k] = v
        res[name] = sd.computeDistances(qd)[0]
        for k in TOGGLES:
            os.environ.pop(k, None)
    old = res["all"]
    bad = torch.nonzero(new != old).reshape(-1)
    print("part", p, "triangles", len(c), "new != old:", int(bad.numel()), {k: int((v != old).sum()) for k, v in res.items()}, flush=True)
    if bad.numel():
        idx = bad[:20].cpu().numpy()
        qs = qd[bad[:20]].cpu().numpy()
        ref = O.SignedDistance(x, y, z, c, 3, False, False, kind=kind)
        want, _, _ = ref.compute(qs, nthreads=1)
        for j, i in enumerate(idx):
            print("   query", int(i), qs[j].tolist(), "oracle %.17g new %.17g old %.17g" % (want[j], float(new[i]), float(old[i])), flush=True)
    del sd, new, old, res
