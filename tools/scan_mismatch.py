"""GPU: list models whose root count differs between the default (coarse-to-fine) and exact_scan modes."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pysurfinv_b200 import api, synth
M = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 51
kind = int(sys.argv[3]) if len(sys.argv) > 3 else 2
lay, nl = synth.crustal_models(M, seed=seed)
per = synth.log_periods()
dl, dn = torch.from_numpy(lay).cuda(), torch.from_numpy(nl).cuda()
a = api.DispersionSolver("cuda:0").forward(dl, dn, per, kind=kind)
b = api.DispersionSolver("cuda:0", opts=api.default_opts(exact_scan=1)).forward(dl, dn, per, kind=kind)
na, nb = a["nfound"].cpu().numpy(), b["nfound"].cpu().numpy()
idx = np.nonzero(na != nb)[0]
dc = (a["c"] - b["c"]).abs().cpu().numpy()
same = na == nb
print("mismatch", len(idx), "of", M, " max dc on same-count models", dc[same].max())
out = {"seed": seed, "M": M, "kind": kind, "idx": idx.tolist(), "n_default": na[idx].tolist(), "n_exact": nb[idx].tolist(),
       "c_exact": b["c"].cpu().numpy()[idx].tolist(), "c_default": a["c"].cpu().numpy()[idx].tolist()}
big = np.argwhere((dc > 1e-4) & same[:, None])
out["big_dc"] = [[int(i), int(k), float(a["c"][i, k]), float(b["c"][i, k])] for i, k in big[:50]]
print("big dc", out["big_dc"][:5])
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "scan_mismatch.json"), "w"))
