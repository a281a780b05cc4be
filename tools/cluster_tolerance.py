"""CPU (host mirror): first-round acceptance rate and accuracy against the oracle as a function of the agreement
tolerance of the two interpolation orders (HM_CTOL) and the widest accepted bracket (HM_WMAX).
usage: HM_CTOL=5e-6 HM_WMAX=3e-3 python tools/cluster_tolerance.py [n_models]"""
import os, sys
import numpy as np
sys.path.insert(0, ".")
from oracle import oracle as O
from pysurfinv_b200 import synth
from tests.hostmirror import mirror as HM
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
fam = sys.argv[2] if len(sys.argv) > 2 else "crustal"
kind = int(sys.argv[3]) if len(sys.argv) > 3 else 2
if fam == "crustal": lay, nl = synth.crustal_models(n, seed=11); per = synth.log_periods()
elif fam == "ragged": lay, nl = synth.ragged_models(n, seed=12); per = np.array([10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 50, 60, 70, 80], np.float32)
elif fam == "hand": lay, nl = synth.hand_models(n, seed=13); per = synth.log_periods(24, 6.0, 60.0)
else: lay, nl = synth.crustal_models(n, seed=14); per = synth.log_periods(100, 5.0, 120.0)
c0, u0, nf0, st0 = O.forward_batch(kind, lay, nl, per, opts=O.make_opts(precision=0), nthreads=8)
tot = dict(sweeps=0, rounds=0, direct=0, windows=0); nev = 0; dc = []
for i in range(n):
    m = int(nl[i])
    r = HM.forward(kind, lay[0, i, :m], lay[1, i, :m], lay[2, i, :m], lay[3, i, :m], lay[4, i, :m], per, G=4)
    for k in tot: tot[k] += r[k]
    nev += r["nfound"]
    if st0[i] == 3: continue
    assert r["nfound"] == nf0[i], (i, r["nfound"], nf0[i])
    dc.append(np.abs(r["c"] - c0[i]))
dc = np.array(dc)
print(fam, kind, "CTOL %s WMAX %s: sweeps/eval %.3f direct %.3f | dc max %.2e p99.9 %.2e median %.2e" % (
    os.environ.get("HM_CTOL", "2e-6"), os.environ.get("HM_WMAX", "3e-3"), tot["sweeps"] / nev, tot["direct"] / nev,
    dc.max(), np.quantile(dc, 0.999), np.median(dc)))
