"""SignedDistance kernel time for different 1/8 shards of the C2 grid on ONE GPU: z-planes k = r mod 8 (what bench.py deals to
rank r of 8) vs contiguous 32-plane slabs, to separate the small-problem effect from the sharding pattern"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from axom_b200 import SignedDistance, synth
x, y, z, conn = synth.icosphere(316)
dev = torch.device("cuda", 0)
ax = torch.linspace(-1.0, 1.0, 256, dtype=torch.float64, device=dev)
sd = SignedDistance(x, y, z, conn, 3, True, True, device=0)


def run(name, planes):
    zz, yy, xx = torch.meshgrid(ax[planes], ax, ax, indexing="ij")
    q = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=1).contiguous()
    phi = torch.empty(q.shape[0], dtype=torch.float64, device=dev)
    sd.computeDistances(q, out=phi)
    sd.setProfiling(1)
    for _ in range(3):
        sd.computeDistances(q, out=phi)
    ms = sd.phase_ms("query.kernel")
    sd.setProfiling(2)
    sd.computeDistances(q, out=phi)
    lt, iv = sd.work_counters()
    sd.setProfiling(0)
    n = q.shape[0]
    print(json.dumps({"shard": name, "queries": n, "kernel_ms": round(ms, 3), "ns_per_query": round(ms * 1e6 / n, 2),
                      "leaf_tests": round(lt / n, 1), "inner_visits": round(iv / n, 1)}), flush=True)


run("all 256 planes", torch.arange(0, 256, device=dev))
run("planes 0 mod 8 (rank 0 of 8)", torch.arange(0, 256, 8, device=dev))
run("planes 3 mod 8 (rank 3 of 8)", torch.arange(3, 256, 8, device=dev))
run("slab 0..31 (outer)", torch.arange(0, 32, device=dev))
run("slab 96..127 (near centre)", torch.arange(96, 128, device=dev))
run("blocks of 4 planes, stride 32 (rank 0)", torch.cat([torch.arange(b, b + 4, device=dev) for b in range(0, 256, 32)]))
run("planes 0 mod 2 (rank 0 of 2)", torch.arange(0, 256, 2, device=dev))
