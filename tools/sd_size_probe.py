"""SignedDistance kernel time vs number of queries (uniform random points in the shell 0.6 < r < 1, no pathological centre
queries), and the effect of the guided chunk size, on one GPU"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from axom_b200 import SignedDistance, synth
x, y, z, conn = synth.icosphere(316)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
d = torch.randn((9_000_000, 3), dtype=torch.float64, device=dev, generator=g)
d = d / d.norm(dim=1, keepdim=True)
r = 0.6 + 0.4 * torch.rand(9_000_000, dtype=torch.float64, device=dev, generator=g)
allq = (d * r[:, None]).contiguous()
for chunk in (os.environ.get("AXB_SD_CHUNK", "128"),):
    sd = SignedDistance(x, y, z, conn, 3, True, True, device=0)
    for n in (65536, 262144, 524288, 1048576, 2097152, 4194304, 8388608):
        q = allq[:n].contiguous()
        phi = torch.empty(n, dtype=torch.float64, device=dev)
        sd.computeDistances(q, out=phi)
        sd.setProfiling(1)
        for _ in range(3):
            sd.computeDistances(q, out=phi)
        ms = sd.phase_ms("query.kernel")
        sd.setProfiling(2)
        sd.computeDistances(q, out=phi)
        lt, iv = sd.work_counters()
        sd.setProfiling(0)
        print(json.dumps({"queries": n, "kernel_ms": round(ms, 3), "ns_per_query": round(ms * 1e6 / n, 2), "leaf_tests": round(lt / n, 1),
                          "inner_visits": round(iv / n, 1)}), flush=True)
