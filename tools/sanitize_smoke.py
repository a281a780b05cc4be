"""GPU box: a small pass over every kernel for compute-sanitizer (memcheck / racecheck / initcheck are slow: keep it
small).  usage: compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from pysurfinv_b200 import api, mc, stack as S, synth  # noqa: E402


def main():
    solver = api.DispersionSolver("cuda:0")
    per = synth.log_periods(16)
    for split in (0, 1):                       # single general launch / fast-path launch + hand-over
        solver.lib.surfdisp_set_split_min_models(split)
        for kind in (2, 1):
            lay, nl = synth.crustal_models(192, seed=900 + kind, lvz=True)
            out = solver.forward(torch.from_numpy(lay).cuda(), torch.from_numpy(nl).cuda(), per, kind=kind)
            print("forward split=%d kind=%d nfound %d..%d" % (split, kind, int(out["nfound"].min()), int(out["nfound"].max())), flush=True)
        lay, nl = synth.ragged_models(96, seed=903)
        h = solver.forward_host(lay, nl, per, 2, chunks=3)
        print("host path split=%d nfound %d..%d" % (split, h["nfound"].min(), h["nfound"].max()), flush=True)
    solver.lib.surfdisp_set_split_min_models(0)
    lay, nl = synth.crustal_models(32, seed=904)
    p = solver.partials(torch.from_numpy(lay).cuda(), torch.from_numpy(nl).cuda(), per[:6])
    print("partials", float(p["dcdb"].abs().max()), flush=True)
    for setting, prior, periods in ((bench.MC_SETTING, S.PRIOR_PRISM, bench.MC_PERIODS), (bench.THERMAL_SETTING, S.PRIOR_OCEAN, bench.THERMAL_PERIODS)):
        t = S.StackTemplate(setting, prior_mask=prior)
        start = torch.from_numpy(np.tile(t.start_values(), (1, 1))).cuda().contiguous()
        truth = solver.mc_propose(t, start, seed=11, step_index=0, reset_mask=torch.ones(1, dtype=torch.uint8, device="cuda"))
        l2, n2 = solver.build_stacks(t, truth)
        obs = solver.forward(l2, n2, periods, group=False)["c"].cpu().numpy()
        ens = mc.ChainEnsemble(solver, t, periods, obs, np.full_like(obs, 0.01), n_chains=8, seed=5, chain_length=3, track_steps=8, use_graph=False)
        ens.run(6)
        print("mc", t.nparams, "params, best misfit", float(ens.best_misfit()), flush=True)
    params = S.config2_params(64, seed=905)
    t, _ = S.config2_template()
    r = solver.forward_params_pinned(t, torch.from_numpy(params).pin_memory(), per)
    print("params path nfound", int(r["nfound"].min()), flush=True)
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
