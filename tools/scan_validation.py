"""GPU box: default path (cluster rounds + coarse scan with kink guard) against exact_scan = 1 (the reference's
point-by-point scan and uniform-section polish everywhere) over several model families and period lists."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from pysurfinv_b200 import api, synth

fast = api.DispersionSolver("cuda:0")
exact = api.DispersionSolver("cuda:0", opts=api.default_opts(exact_scan=1))
P19 = np.array([8, 10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 50, 60, 70, 80], np.float32)
P7 = np.array([10, 14, 20, 28, 40, 60, 80], np.float32)
cases = [("crustal/19 periods", synth.crustal_models(65536, seed=301), P19),
         ("crustal/7 periods", synth.crustal_models(65536, seed=302), P7),
         ("crustal/100 periods 5-120 s", synth.crustal_models(16384, seed=303), synth.log_periods(100, 5.0, 120.0)),
         ("ragged+water/19 periods", synth.ragged_models(65536, seed=304), P19),
         ("hand 4-layer/24 periods", synth.hand_models(65536, seed=305), synth.log_periods(24, 6.0, 60.0))]
bad_total = 0
for name, (lay, nl), per in cases:
    dl, dn = torch.from_numpy(lay).cuda(), torch.from_numpy(nl).cuda()
    for kind in (2, 1):
        a = fast.forward(dl, dn, per, kind=kind)
        b = exact.forward(dl, dn, per, kind=kind)
        na, nb = a["nfound"], b["nfound"]
        bad = int((na != nb).sum())
        same = (na == nb)
        dc = float((a["c"] - b["c"]).abs()[same].max())
        du = float((a["u"] - b["u"]).abs()[same].median())
        bad_total += bad
        print("%-30s kind %d: root-count mismatches %d of %d, max|dc| %.2e, median|dU| %.2e, full %.3f"
              % (name, kind, bad, lay.shape[1], dc, du, float((nb == len(per)).float().mean())))
print("TOTAL mismatches", bad_total)
