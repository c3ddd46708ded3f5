"""host->host SignedDistance query: wall time and per-chunk phase times with / without the two-stream chunk pipeline"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from axom_b200 import SignedDistance, synth
x, y, z, conn = synth.icosphere(316)
q = torch.from_numpy(synth.uniform_grid_points(-1, 1, 256)).pin_memory()
phi = torch.empty(q.shape[0], dtype=torch.float64).pin_memory()
qn, pn = q.numpy(), phi.numpy()
for chunk in (0, 2097152, 4194304, 8388608):
    os.environ["AXB_SD_PIPE_CHUNK"] = str(chunk)
    sd = SignedDistance(x, y, z, conn)
    sd.computeDistances(qn, out=pn)
    for prof in (0, 1):
        sd.setProfiling(prof)
        t0 = time.perf_counter()
        for _ in range(3):
            sd.computeDistances(qn, out=pn)
        dt = (time.perf_counter() - t0) / 3 * 1e3
        if prof:
            print(chunk, "wall %.1f ms" % dt, "total %.2f kernel(mean per chunk) %.2f sortq %.2f" % (sd.phase_ms("query.total"), sd.phase_ms("query.kernel"), sd.phase_ms("query.sortq")), flush=True)
        else:
            print(chunk, "wall %.1f ms (no profiling)" % dt, flush=True)
