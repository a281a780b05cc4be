"""Debug helper (GPU box): per-period comparison of the CUDA path with the oracle on a few models."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import oracle as O
from pysurfinv_b200 import api, synth

M = int(sys.argv[1]) if len(sys.argv) > 1 else 64
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lay, nl = synth.crustal_models(M, seed=11)
per = synth.log_periods()
s = api.DispersionSolver("cuda:0")
out = s.forward(torch.from_numpy(lay).cuda(), torch.from_numpy(nl).cuda(), per, kind=kind)
torch.cuda.synchronize()
g = {k: v.cpu().numpy() for k, v in out.items()}
c0, u0, nf0, st0 = O.forward_batch(kind, lay, nl, per, opts=O.make_opts(precision=0), nthreads=8)
print("keys", list(g.keys()))
bad = np.nonzero(g["nfound"] != nf0)[0]
print("nfound mismatches:", len(bad), "of", M)
for i in bad[:8]:
    print("model", i, "gpu nfound", g["nfound"][i], "oracle", nf0[i], "flags", g.get("flags", np.zeros(M, int))[i])
    kk = min(g["nfound"][i], nf0[i])
    print("  gpu c   ", np.round(g["c"][i][max(0, kk - 3):kk + 2], 5))
    print("  oracle c", np.round(c0[i][max(0, kk - 3):kk + 2], 5))
dc = np.abs(g["c"] - c0)
ok = g["nfound"] == nf0
print("dc max over matching models %.3e" % dc[ok].max(), "worst", np.unravel_index(dc[ok].argmax(), dc[ok].shape))
for i in bad[:3]:
    print("model", i)
    print(" gpu   ", np.round(g["c"][i][:12], 4))
    print(" oracle", np.round(c0[i][:12], 4))
