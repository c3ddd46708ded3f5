"""How much would a good first bound help the DistributedClosestPoint search?  Unbounded search vs a search bounded by the
exact answer (the best bound a hint could give) and by the answer of a query 16 Morton ranks away."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from axom_b200 import DistributedClosestPoint, synth
x, y, z, _ = synth.icosphere(1000)
P = np.stack([x, y, z], 1)
d = DistributedClosestPoint(3, device=0)
d.setObjectMesh([P])
d.generateBVHTree()
g = torch.Generator(device="cuda"); g.manual_seed(5)
q = (torch.rand((5_000_000, 3), generator=g, dtype=torch.float64, device="cuda") * 2 - 1).contiguous()
b = d._b
def timed(f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); return r, (time.perf_counter() - t0) * 1e3
st, _ = timed(lambda: b.compute_local(0, q))
for env in ("1", None, "1", None):
    if env:
        os.environ["AXB_DCP_NO_HINT"] = env
    else:
        os.environ.pop("AXB_DCP_NO_HINT", None)
    st, t = timed(lambda: b.compute_local(0, q))
    print("first-visit search, sample-pass bounds", "off" if env else "on", round(t, 1), "ms", flush=True)
st, t_un = timed(lambda: b.compute_local(0, q))
sq = ((st["cp_coords"] - q) ** 2)
sq = sq[:, 0] + sq[:, 1] + sq[:, 2]
st2, t_exact = timed(lambda: b.compute_bounded(0, q, sq.contiguous()))
loose = (sq.sqrt() + 0.03) ** 2
st3, t_loose = timed(lambda: b.compute_bounded(0, q, loose.contiguous()))
print({"unbounded_ms": round(t_un, 1), "bounded_by_exact_ms": round(t_exact, 1), "bounded_by_exact_plus_0.03_ms": round(t_loose, 1),
       "same": bool(torch.equal(st["cp_index"], st2["cp_index"]) and torch.equal(st["cp_index"], st3["cp_index"]))})
