import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from pysurfinv_b200 import api, synth
from oracle import oracle as O
fast = api.DispersionSolver("cuda:0")
exact = api.DispersionSolver("cuda:0", opts=api.default_opts(exact_scan=1))
lay, nl = synth.crustal_models(16384, seed=303)
per = synth.log_periods(100, 5.0, 120.0)
dl, dn = torch.from_numpy(lay).cuda(), torch.from_numpy(nl).cuda()
a = fast.forward(dl, dn, per, kind=1); b = exact.forward(dl, dn, per, kind=1)
na, nb = a["nfound"].cpu().numpy(), b["nfound"].cpu().numpy()
idx = np.nonzero(na != nb)[0]
print("mismatches", len(idx), "fast>exact", int((na[idx] > nb[idx]).sum()), "fast<exact", int((na[idx] < nb[idx]).sum()))
sub = idx[:40]
c0, u0, nf0, st0 = O.forward_batch(1, lay[:, sub], nl[sub], per, opts=O.make_opts(precision=0), nthreads=8)
print("oracle agrees with fast:", int((nf0 == na[sub]).sum()), "with exact:", int((nf0 == nb[sub]).sum()), "of", len(sub))
for j, i in enumerate(sub[:6]):
    k = min(na[i], nb[i])
    print(i, "fast nf", na[i], "exact nf", nb[i], "oracle nf", nf0[j], "flags", int(a["flags"][i]), int(b["flags"][i]),
          "c fast", a["c"][i, k - 2:k + 2].cpu().numpy(), "c exact", b["c"][i, k - 2:k + 2].cpu().numpy(), "oracle", c0[j, k - 2:k + 2], "b_hs", lay[1, i, nl[i] - 1])
