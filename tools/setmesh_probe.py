"""Device time of the phases of SignedDistance::setMesh (always recorded at creation) on the C2 and C4 surfaces"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from axom_b200 import SignedDistance, synth
for freq in (316, 1000):
    x, y, z, conn = synth.icosphere(freq)
    xd, yd, zd = (torch.from_numpy(a).cuda() for a in (x, y, z))
    cd = torch.from_numpy(conn).cuda()
    for rep in range(3):
        sd = SignedDistance(xd, yd, zd, cd, 3, True, True, device=0)
        row = {"triangles": len(conn), "rep": rep}
        for ph in ("setmesh.total", "setmesh.upload", "setmesh.cell_boxes", "setmesh_build.total", "setmesh.gather_soup", "setmesh.obb_build"):
            row[ph] = round(sd.phase_ms(ph), 3)
        del sd
        torch.cuda.synchronize()
    print(json.dumps(row), flush=True)
