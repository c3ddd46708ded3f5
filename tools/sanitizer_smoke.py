import sys, numpy as np
sys.path.insert(0, ".")
import torch
from axom_b200 import SignedDistance, synth, DistributedClosestPoint, BVH
from axom_b200.comm import Comm, unique_id
rng = np.random.default_rng(3)
x, y, z, conn = synth.icosphere(20)
q = np.concatenate([np.zeros((1, 3)), rng.normal(0, 1e-6, (16, 3)), rng.normal(0, 0.02, (200, 3)), rng.uniform(-1.2, 1.2, (int(__import__("os").environ.get("SAN_NQ", "300000")), 3))])
for cs in (True, False):
    sd = SignedDistance(x, y, z, conn, 3, True, cs)
    phi, cp, nr = sd.computeDistances(q, True, True)
    phid = sd.computeDistances(torch.from_numpy(q).cuda())[0]
    print("sd", cs, float(np.abs(phi).sum()), float(phid.abs().sum()))
comm = Comm(1, 0, unique_id(), 0)
sd = SignedDistance(x, y, z, conn, 3, False, False)
print("minreduce", float(sd.computeDistancesMinReduce(comm, q[:50000]).sum()))
pts = rng.random((20000, 3))
d = DistributedClosestPoint(3, device=0)
d.setComm(comm)
d.setObjectMesh([pts])
d.generateBVHTree()
out = d.computeClosestPoints(torch.from_numpy(rng.random((30000, 3))).cuda())
print("dcp", int(out["cp_index"].sum()))
b = BVH(3, device=0)
b.initialize(synth.triangle_aabbs(50000, seed=1))
rays = np.concatenate([rng.uniform(-1, 1, (40000, 3)), rng.normal(0, 1, (40000, 3))], axis=1)
o, c, cand = b.findRays(rays)
print("rays", int(np.sum(c)))
