"""Per-slab time of the two SignedDistance kernels on ONE GPU for the z-slabs bench.py deals to the ranks of an N-GPU run
(N = 8, 4, 2, 1): which slab sets the step, and in which kernel."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from axom_b200 import SignedDistance, synth
x, y, z, conn = synth.icosphere(316)
dev = torch.device("cuda", 0)
ax = torch.linspace(-1.0, 1.0, 256, dtype=torch.float64, device=dev)
sd = SignedDistance(x, y, z, conn, 3, True, True, device=0)
mode = "planes" if "planes" in sys.argv[1:] else ("blocks4" if "blocks4" in sys.argv[1:] else "slabs")
worlds = [int(a) for a in sys.argv[1:] if a.isdigit()] or [8, 4, 2, 1]
for world in worlds:
    rows = []
    for rank in range(world):
        if mode == "planes":
            planes = torch.arange(rank, 256, world, device=dev)
        elif mode == "blocks4":  # blocks of 4 planes dealt round-robin
            planes = torch.cat([torch.arange(b, b + 4, device=dev) for b in range(4 * rank, 256, 4 * world)])
        else:
            planes = torch.arange((256 * rank) // world, (256 * (rank + 1)) // world, device=dev)
        zz, yy, xx = torch.meshgrid(ax[planes], ax, ax, indexing="ij")
        q = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=1).contiguous()
        phi = torch.empty(q.shape[0], dtype=torch.float64, device=dev)
        sd.computeDistances(q, out=phi)
        sd.setProfiling(1)
        for _ in range(3):
            sd.computeDistances(q, out=phi)
        row = {"world": world, "rank": rank, "queries": q.shape[0]}
        for ph in ("query.sortq", "query.kernel", "query.min", "query.resolve", "query.total"):
            try:
                row[ph.split(".")[1] + "_ms"] = round(sd.phase_ms(ph), 3)
            except Exception:
                pass
        sd.setProfiling(2)
        sd.computeDistances(q, out=phi)
        lt, iv = sd.work_counters()
        row["leaf_tests"], row["inner_visits"] = round(lt / q.shape[0], 2), round(iv / q.shape[0], 2)
        sd.setProfiling(0)
        rows.append(row)
        print(json.dumps(row), flush=True)
    tot = [r["total_ms"] for r in rows]
    print(json.dumps({"world": world, "max_total_ms": max(tot), "mean_total_ms": sum(tot) / len(tot),
                      "efficiency_vs_mean": sum(tot) / len(tot) / max(tot)}), flush=True)
