"""GPU box: end-to-end (pinned host -> device -> host) throughput of forward_pinned for several chunk counts."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from pysurfinv_b200 import api, synth
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
per = synth.log_periods(40)
lay, nl = synth.crustal_models(M, seed=1)
s = api.DispersionSolver("cuda:0")
hl = torch.from_numpy(lay).pin_memory(); hn = torch.from_numpy(nl).pin_memory()
for chunks in (1, 2, 4, 8, 16):
    for _ in range(2):
        s.forward_pinned(hl, hn, per, 2, chunks=chunks)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        s.forward_pinned(hl, hn, per, 2, chunks=chunks)
    dt = (time.perf_counter() - t0) / 3
    print("chunks %2d: %.1f ms per step, %.1f M evals/s" % (chunks, dt * 1e3, M * 40 / dt / 1e6))
# raw copy bandwidths
d = torch.empty_like(hl, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(hl, non_blocking=True); torch.cuda.synchronize()
print("H2D %.1f GB/s" % (hl.numel() * 4 / (time.perf_counter() - t0) / 1e9))
