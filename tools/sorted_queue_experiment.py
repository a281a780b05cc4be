"""GPU box: does the order of the models in the queue matter?  Same models, (a) as generated, (b) sorted by the
top-layer Vs (a proxy of the length of the first-period scan), (c) sorted by the first-period phase velocity."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from pysurfinv_b200 import api, synth
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
per = synth.log_periods(40)
lay, nl = synth.crustal_models(M, seed=1)
s = api.DispersionSolver("cuda:0")
def run(order, name):
    dl = torch.from_numpy(np.ascontiguousarray(lay[:, order])).cuda(); dn = torch.from_numpy(np.ascontiguousarray(nl[order])).cuda()
    for _ in range(2): out = s.forward(dl, dn, per, kind=2)
    acc = []
    for _ in range(3):
        kms = [0.0, 0.0, 0.0]
        s.forward(dl, dn, per, kind=2, kernel_ms=kms); acc.append(kms)
    ms = np.mean(acc, axis=0)
    print("%-28s prep %.2f phase1 %.2f phase2 %.2f" % (name, ms[0], ms[1], ms[2]))
    return out
ident = np.arange(M)
out = run(ident, "as generated")
run(np.argsort(lay[1, :, 0], kind="stable"), "sorted by top-layer Vs")
c0 = out["c"][:, 0].cpu().numpy()
run(np.argsort(c0 - 0.9 * lay[1, :, 0], kind="stable"), "sorted by scan distance")
run(np.argsort(out["c"][:, 20].cpu().numpy(), kind="stable"), "sorted by c(T20)")
