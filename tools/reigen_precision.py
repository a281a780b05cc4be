"""CPU: REIGEN's ODE state in float32 (packed pairs, every sub-layer re-orthogonalised: the product's default) against
float64 (opts.group_f64 = 1), both through the host build of the kernels' per-lane code (tests/hostmirror), and both
against the float32 oracle and its float64-solver spread.

usage: python tools/reigen_precision.py <family> <models>      families: crustal crustal100 lvz hand ragged deep147 deep500 thermal
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "hostmirror")); sys.path.insert(0, os.path.join(ROOT, "tests"))
fam = sys.argv[1]; M = int(sys.argv[2]); mode = sys.argv[3] if len(sys.argv)>3 else None
from pysurfinv_b200 import synth
def gen():
    if fam=='crustal': return synth.crustal_models(M, seed=5242), synth.log_periods()
    if fam=='crustal100': return synth.crustal_models(M, seed=5243), synth.log_periods(100, 5.0, 120.0)
    if fam=='lvz': return synth.crustal_models(M, seed=5244, lvz=True), synth.log_periods(100, 5.0, 120.0)
    if fam=='hand': return synth.hand_models(M, seed=5245), synth.log_periods(24, 6.0, 60.0)
    if fam=='ragged': return synth.ragged_models(M, seed=5246), np.array([10,12,14,16,18,20,22,24,26,28,30,32,36,40,50,60,70,80],np.float32)
    if fam=='deep147': return synth.crustal_models(M, seed=5247, n_crust=15, n_mantle=130, zmax=400.0), np.arange(10.0, 151.0, 10.0, dtype=np.float32)
    if fam=='deep500': return synth.crustal_models(M, seed=5248, n_crust=40, n_mantle=455, zmax=600.0), synth.log_periods(60, 5.0, 200.0)
    if fam=='thermal':
        import bench
        from oracle import model_builder as MB
        from pysurfinv_b200 import stack as S
        t = S.StackTemplate(bench.THERMAL_SETTING, prior_mask=S.PRIOR_OCEAN); lo, hi, _ = t.bounds(); rng = np.random.default_rng(5249); ps=[]
        while len(ps) < M:
            p = (lo + (hi - lo) * rng.random(t.nparams)).astype(np.float32)
            if MB.priors_ocean(t, p.astype(np.float64)) & S.PRIOR_OCEAN == 0: ps.append(p)
        return MB.build_stacks(t, np.array(ps), t.max_layers()), bench.THERMAL_PERIODS
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
(lay, nl), per = gen()
if mode is not None:
    import mirror
    U = np.zeros((M, len(per)), np.float32); C = np.zeros_like(U)
    for m in range(M):
        n = nl[m]
        r = mirror.forward(2, lay[0,m,:n], lay[1,m,:n], lay[2,m,:n], lay[3,m,:n], lay[4,m,:n], per)
        U[m,:len(r['u'])] = r['u']; C[m,:len(r['c'])] = r['c']
    np.save(os.path.join(ROOT, 'gpurun_out', 'U_%s_%s.npy' % (fam, mode)), U)
    sys.exit(0)
from oracle import oracle as O
for mode, env in (('f64', {'HM_REIGEN_F64': '1'}), ('f32', {})):
    e = dict(os.environ); e.update(env)
    subprocess.check_call([sys.executable, __file__, fam, str(M), mode], env=e)
U64 = np.load(os.path.join(ROOT, 'gpurun_out', 'U_%s_f64.npy' % fam)); U32 = np.load(os.path.join(ROOT, 'gpurun_out', 'U_%s_f32.npy' % fam))
c0,u0,nf0,st0 = O.forward_batch(2, lay, nl, per, opts=O.make_opts(precision=0), nthreads=4)
c1,u1,nf1,st1 = O.forward_batch(2, lay, nl, per, opts=O.make_opts(precision=1), nthreads=4)
ok = (np.arange(len(per))[None,:] < nf0[:,None])
def st(x): return 'frac>1e-4 %.2e p99.9 %.2e max %.2e med %.2e' % ((x>1e-4).mean(), np.quantile(x,0.999), x.max(), np.median(x))
print(fam, int(ok.sum()), 'evals, layers', int(nl.max()))
print('   f64 vs oracle :', st(np.abs(U64-u0)[ok]))
print('   f32 vs oracle :', st(np.abs(U32-u0)[ok]))
print('   ref own noise :', st(np.abs(u0-u1)[ok]))
print('   f32 vs f64    :', st(np.abs(U32-U64)[ok]))
