"""GPU box: cost per executed layer-step of phase 1 when all periods are long (every sweep spans the full stack)
versus the 8-80 s range of config 2 (depths 30..77 mixed inside a warp)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from pysurfinv_b200 import api, synth
M = 1 << 19
lay, nl = synth.crustal_models(M, seed=5)
s = api.DispersionSolver("cuda:0")
dl, dn = torch.from_numpy(lay).cuda(), torch.from_numpy(nl).cuda()
for name, per in (("8-80 s", synth.log_periods(40, 8.0, 80.0)), ("20-80 s", synth.log_periods(40, 20.0, 80.0)),
                  ("30-80 s", synth.log_periods(40, 30.0, 80.0))):
    out = s.forward(dl, dn, per, kind=2)
    for _ in range(2):
        s.forward(dl, dn, per, kind=2, out=out)
    ms = [0, 0, 0]
    s.forward(dl, dn, per, kind=2, out=out, kernel_ms=ms)
    steps, sweeps, subu, models = s.counters()
    print("%8s: phase1 %.1f ms, %.0f layer-steps/eval, %.1f sweeps/eval, depth %.1f, %.3f ps per layer-step"
          % (name, ms[1], steps / (M * 40), sweeps / (M * 40), steps / sweeps, ms[1] * 1e9 / steps))
