"""CPU: like mirror_hunt.py, on coarse period lists (5, 7, 15 periods) and config-4 deep stacks: root counts and |dc| > 1e-4."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from oracle import oracle as O
from pysurfinv_b200 import synth
from tests.hostmirror import mirror as HM
P7 = np.array([10, 14, 20, 28, 40, 60, 80], np.float32)
P15 = np.arange(10.0, 151.0, 10.0, dtype=np.float32)
P5 = np.array([8, 16, 32, 64, 128], np.float32)
fams = [("deep/P15", lambda s: synth.crustal_models(100, seed=s, n_crust=15, n_mantle=130, zmax=400.0), P15),
        ("crustal/P7", lambda s: synth.crustal_models(200, seed=s), P7),
        ("ragged/P7", lambda s: synth.ragged_models(200, seed=s), P7),
        ("crustal/P5", lambda s: synth.crustal_models(200, seed=s), P5),
        ("hand/P7", lambda s: synth.hand_models(200, seed=s), P7),
        ("deep/P5", lambda s: synth.crustal_models(100, seed=s, n_crust=15, n_mantle=130, zmax=400.0), P5)]
t0=time.time(); total=0; bad=0; big=0; worst=0
for s in range(4000, 4025):
    for name, gen, per in fams:
        lay, nl = gen(s)
        for kind in (2, 1):
            c0, u0, nf0, st0 = O.forward_batch(kind, lay, nl, per, opts=O.make_opts(precision=0), nthreads=8)
            for i in range(lay.shape[1]):
                if st0[i] == 3: continue
                n=int(nl[i])
                r = HM.forward(kind, lay[0,i,:n], lay[1,i,:n], lay[2,i,:n], lay[3,i,:n], lay[4,i,:n], per, G=4)
                total+=1
                if r["nfound"] != nf0[i]:
                    bad+=1; print("MISMATCH seed %d %s kind %d model %d: mirror %d oracle %d"%(s,name,kind,i,r["nfound"],nf0[i]), flush=True)
                else:
                    d=float(np.abs(r["c"]-c0[i]).max()); worst=max(worst,d)
                    if d > 1e-4: big+=1; print("BIG dc %.3g seed %d %s kind %d model %d"%(d,s,name,kind,i), flush=True)
    if s % 5 == 4: print("seed", s, "curves", total, "count-mismatch", bad, "big-dc", big, "worst dc %.2e"%worst, "%.0f s"%(time.time()-t0), flush=True)
