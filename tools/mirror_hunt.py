"""CPU: hunts for rare disagreements between the kernels' logic (host mirror) and the oracle over many random
models: root counts and |dc|.  usage: python tools/mirror_hunt.py <seed0> <n_seeds> <models_per_seed>"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from oracle import oracle as O
from pysurfinv_b200 import synth
from tests.hostmirror import mirror as HM
seed0, nseeds, M = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
P18 = np.array([10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 50, 60, 70, 80], np.float32)
fams = [("crustal40", lambda s: synth.crustal_models(M, seed=s), synth.log_periods()),
        ("crustal100", lambda s: synth.crustal_models(M, seed=s), synth.log_periods(100, 5.0, 120.0)),
        ("ragged18", lambda s: synth.ragged_models(M, seed=s), P18),
        ("hand24", lambda s: synth.hand_models(M, seed=s), synth.log_periods(24, 6.0, 60.0))]
t0 = time.time(); total = 0; nbad = 0; worst = 0.0
for s in range(seed0, seed0 + nseeds):
    for name, gen, per in fams:
        lay, nl = gen(s)
        for kind in (2, 1):
            c0, u0, nf0, st0 = O.forward_batch(kind, lay, nl, per, opts=O.make_opts(precision=0), nthreads=8)
            for i in range(M):
                if st0[i] == 3: continue
                n = int(nl[i])
                r = HM.forward(kind, lay[0, i, :n], lay[1, i, :n], lay[2, i, :n], lay[3, i, :n], lay[4, i, :n], per, G=4)
                total += 1
                if r["nfound"] != nf0[i]:
                    nbad += 1
                    print("MISMATCH seed %d %s kind %d model %d: mirror %d oracle %d" % (s, name, kind, i, r["nfound"], nf0[i]), flush=True)
                else:
                    d = float(np.abs(r["c"] - c0[i]).max())
                    if d > worst:
                        worst = d
                        if d > 2e-5: print("dc %.2e seed %d %s kind %d model %d" % (d, s, name, kind, i), flush=True)
    print("seed %d done: %d curves, %d mismatches, worst dc %.2e, %.0f s" % (s, total, nbad, worst, time.time() - t0), flush=True)
