"""GPU box: device time of the model-assembly / prior / proposal kernels for the two Monte-Carlo settings of bench.py
(B-spline crust + mantle, and the ocean model with the thermal mantle).  usage: python tools/builder_cost.py [M ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from pysurfinv_b200 import api, stack as S  # noqa: E402


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [256, 32768]
    solver = api.DispersionSolver("cuda:0")
    for name, setting, prior, per in (("bspline", bench.MC_SETTING, S.PRIOR_PRISM, bench.MC_PERIODS),
                                      ("thermal", bench.THERMAL_SETTING, S.PRIOR_OCEAN, bench.THERMAL_PERIODS)):
        t = S.StackTemplate(setting, prior_mask=prior)
        for M in sizes:
            start = torch.from_numpy(np.tile(t.start_values(), (M, 1))).cuda().contiguous()
            ones = torch.ones(M, dtype=torch.uint8, device="cuda")
            cur = solver.mc_propose(t, start, seed=3, step_index=0, reset_mask=ones)      # admissible uniform redraws
            bad = int((solver.check_priors(t, cur) & prior != 0).sum())
            step = [1]

            def prop():
                step[0] += 1
                solver.mc_propose(t, cur, seed=3, step_index=step[0])
            ms_b = timed(lambda: solver.build_stacks(t, cur))
            ms_c = timed(lambda: solver.check_priors(t, cur))
            ms_p = timed(prop)
            ms_r = timed(lambda: solver.mc_propose(t, cur, seed=5, step_index=7, reset_mask=ones), reps=2)
            print("%-8s M=%6d  build %.3f ms  check_priors %.3f ms  propose %.3f ms  uniform restart %.3f ms  (inadmissible after restart: %d)"
                  % (name, M, ms_b, ms_c, ms_p, ms_r, bad), flush=True)
            # the solver on these stacks: all periods, the first period alone, all periods with the own curve as hint
            lay, nl = solver.build_stacks(t, cur)
            res = solver.forward(lay, nl, per, group=False)
            hint = res["c"].clone()
            ms_all = timed(lambda: solver.forward(lay, nl, per, group=False))
            ms_first = timed(lambda: solver.forward(lay, nl, per[:1], group=False))
            ms_hint = timed(lambda: solver.forward(lay, nl, per, group=False, hint=hint))
            print("%-8s M=%6d  solve c (%d periods, %d layers) %.3f ms  first period alone %.3f ms  with hints %.3f ms  c(T0) %.2f-%.2f km/s, top Vs %.2f"
                  % (name, M, len(per), int(nl.max()), ms_all, ms_first, ms_hint, float(res["c"][:, 0].min()), float(res["c"][:, 0].max()),
                     float(lay[1, :, 0].min())), flush=True)


if __name__ == "__main__":
    main()
