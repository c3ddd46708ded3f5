"""SignedDistance kernel cost by distance-from-centre shell on the C2 workload: ms, leaf tests and inner visits per query"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from axom_b200 import SignedDistance, synth
x, y, z, conn = synth.icosphere(316)
dev = torch.device("cuda", 0)
ax = torch.linspace(-1.0, 1.0, 256, dtype=torch.float64, device=dev)
zz, yy, xx = torch.meshgrid(ax, ax, ax, indexing="ij")
q = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=1).contiguous()
r = q.norm(dim=1)
sd = SignedDistance(x, y, z, conn, 3, True, True, device=0)
edges = [0.0, 0.02, 0.05, 0.1, 0.2, 0.3, 0.4, 0.47, 0.53, 0.6, 0.8, 1.0, 2.0]
for lo, hi in zip(edges[:-1], edges[1:]):
    sub = q[(r >= lo) & (r < hi)].contiguous()
    if sub.shape[0] == 0:
        continue
    phi = torch.empty(sub.shape[0], dtype=torch.float64, device=dev)
    sd.computeDistances(sub, out=phi)
    sd.setProfiling(1)
    for _ in range(2):
        sd.computeDistances(sub, out=phi)
    ms = sd.phase_ms("query.kernel")
    sd.setProfiling(2)
    sd.computeDistances(sub, out=phi)
    lt, iv = sd.work_counters()
    sd.setProfiling(0)
    n = sub.shape[0]
    print(json.dumps({"r": [lo, hi], "queries": n, "kernel_ms": round(ms, 3), "ns_per_query": round(ms * 1e6 / n, 2),
                      "leaf_tests": round(lt / n, 1), "inner_visits": round(iv / n, 1)}), flush=True)
