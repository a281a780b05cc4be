"""GPU box: kernel split of the Love sweep (same workload as bench.py)."""
import sys
import torch
sys.path.insert(0, ".")
from pysurfinv_b200 import api, synth
M = 1 << 19
lay, nl = synth.crustal_models(M, seed=5)
per = synth.log_periods(40)
s = api.DispersionSolver("cuda:0")
dl, dn = torch.from_numpy(lay).cuda(), torch.from_numpy(nl).cuda()
for kind in (2, 1):
    out = s.forward(dl, dn, per, kind=kind)
    for _ in range(2):
        s.forward(dl, dn, per, kind=kind, out=out)
    ms = [0, 0, 0]
    s.forward(dl, dn, per, kind=kind, out=out, kernel_ms=ms)
    print("kind", kind, "prep/phase1/phase2 ms", [round(x, 1) for x in ms], "counters", s.counters())
