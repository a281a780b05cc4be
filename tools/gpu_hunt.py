"""GPU hunt: the CUDA path (through the C ABI) against the float32 oracle on many random curves per model family
and wave type -- root counts, |dc|, |dU| -- with the oracle's own float32-vs-float64-solver spread beside it.

usage: python tools/gpu_hunt.py [curves_per_family] [out.json]        (run on the GPU box)

`hunt()` is also what tests/test_gpu_hunt.py runs (smaller).  The oracle is the checker here, never the thing
measured.  Families: config-2 stacks on the bench period list, the same on 100 periods 5-120 s (long periods:
roots just below the half-space velocity), stacks with crustal low-velocity zones and slow half-spaces
(velocity inversions: scan rounds on own truncations, calcul.f:155-159), ragged stacks with water layers on the
18-period list of point.py:400, hand models with the un-clamped ndiv = 5.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from pysurfinv_b200 import api, synth  # noqa: E402

P18 = np.array([10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 50, 60, 70, 80], np.float32)   # point.py:400

FAMILIES = {
    "crustal77_40": (lambda M, s: synth.crustal_models(M, seed=s), synth.log_periods()),
    "crustal77_100": (lambda M, s: synth.crustal_models(M, seed=s), synth.log_periods(100, 5.0, 120.0)),
    "lvz77_40": (lambda M, s: synth.crustal_models(M, seed=s, lvz=True), synth.log_periods()),
    "lvz77_100": (lambda M, s: synth.crustal_models(M, seed=s, lvz=True), synth.log_periods(100, 5.0, 120.0)),
    "ragged_water_18": (lambda M, s: synth.ragged_models(M, seed=s), P18),
    "hand4_24": (lambda M, s: synth.hand_models(M, seed=s), synth.log_periods(24, 6.0, 60.0)),
}


def _q(x, q):
    return float(np.quantile(x, q)) if x.size else 0.0


def hunt(solver, family, kind, curves, seed=7001, chunk=16384, noise=True, log=None):
    """Returns the statistics of `curves` models of `family` for wave type `kind` (1 Love, 2 Rayleigh)."""
    import torch
    gen, per = FAMILIES[family]
    K = len(per)
    nth = os.cpu_count() or 1
    r = dict(family=family, kind=kind, periods=K, curves=0, evaluations=0, lstop_excluded=0, nfound_mismatch=0, unexplained_mismatch=0,
             mismatches=[], dc_max=0.0, dc_gt_1e4=0, du_gt_1e4=0, du_max=0.0, noise_du_gt_1e4=0, noise_du_max=0.0,
             noise_dc_max=0.0, n_eval_cmp=0, full_curves=0, worst_du=[])
    dcs, dus, nus = [], [], []
    done = 0
    while done < curves:
        m = min(chunk, curves - done)
        s = seed + done
        lay, nl = gen(m, s)
        out = solver.forward(torch.from_numpy(lay).cuda(), torch.from_numpy(nl).cuda(), per, kind=kind)
        g = {k: v.cpu().numpy() for k, v in out.items()}
        c0, u0, nf0, st0 = O.forward_batch(kind, lay, nl, per, opts=O.make_opts(precision=0), nthreads=nth)
        ok = st0 != 3          # the reference's LSTOP aborts are excluded and counted
        bad = ok & (g["nfound"] != nf0)
        for i in np.nonzero(bad)[0]:
            # Which side does the float64 solver on the same float32 model take?  A root that only one of the two
            # float32 evaluations sees is float32 noise of the reference itself: (a) the float64 solver agrees with the
            # CUDA path, or (b) both curves end early, one period apart (the last root sits on the square-root cusp
            # at the half-space velocity, where the secular function touches zero within its rounding noise).
            c1, u1, nf1, st1 = O.forward_batch(kind, lay[:, i:i + 1], nl[i:i + 1], per, opts=O.make_opts(precision=1))
            gi, oi = int(g["nfound"][i]), int(nf0[i])
            cls = "f64_solver_agrees_with_gpu" if int(nf1[0]) == gi else ("cusp_cutoff_one_period_apart" if (abs(gi - oi) == 1 and max(gi, oi) < K) else "unexplained")
            r["unexplained_mismatch"] += int(cls == "unexplained")
            if len(r["mismatches"]) < 40:
                r["mismatches"].append(dict(seed=int(s), model=int(i), gpu=gi, oracle=oi, oracle_f64_solver=int(nf1[0]), kind_of=cls))
        same = ok & ~bad
        dc = np.abs(g["c"] - c0)[same]
        du = np.abs(g["u"] - u0)[same]
        r["curves"] += int(ok.sum()); r["lstop_excluded"] += int((~ok).sum()); r["nfound_mismatch"] += int(bad.sum())
        r["evaluations"] += int(nf0[ok].sum()); r["full_curves"] += int((nf0 == K).sum())
        r["dc_max"] = max(r["dc_max"], float(dc.max()) if dc.size else 0.0)
        r["du_max"] = max(r["du_max"], float(du.max()) if du.size else 0.0)
        duf = np.abs(g["u"] - u0); duf[~same] = 0
        for flat in np.argsort(duf.ravel())[-3:]:
            i, k = np.unravel_index(flat, duf.shape)
            if duf[i, k] > 1e-3:
                r["worst_du"].append(dict(seed=int(s), model=int(i), k=int(k), T=float(per[k]), c_gpu=float(g["c"][i, k]), c_oracle=float(c0[i, k]),
                                          u_gpu=float(g["u"][i, k]), u_oracle=float(u0[i, k]), nfound=int(nf0[i])))
        r["worst_du"] = sorted(r["worst_du"], key=lambda e: -abs(e["u_gpu"] - e["u_oracle"]))[:10]
        r["dc_gt_1e4"] += int((dc > 1e-4).sum()); r["du_gt_1e4"] += int((du > 1e-4).sum()); r["n_eval_cmp"] += int(dc.size)
        dcs.append(dc.ravel()[:: max(1, dc.size // 200000)]); dus.append(du.ravel()[:: max(1, du.size // 200000)])
        if noise and done == 0:
            # the reference's own float32 noise (float32 solver vs float64 solver on the same float32 model), on the first chunk
            mm = min(m, 4096)
            c1, u1, nf1, st1 = O.forward_batch(kind, lay[:, :mm], nl[:mm], per, opts=O.make_opts(precision=1), nthreads=nth)
            sm = (st0[:mm] != 3) & (nf0[:mm] == nf1)
            nu = np.abs(u0[:mm] - u1)[sm]; nc = np.abs(c0[:mm] - c1)[sm]
            r["noise_du_gt_1e4"] = int((nu > 1e-4).sum()); r["noise_du_max"] = float(nu.max()) if nu.size else 0.0
            r["noise_dc_max"] = float(nc.max()) if nc.size else 0.0; r["noise_n"] = int(nu.size)
            r["noise_nfound_mismatch"] = int((nf0[:mm] != nf1).sum())
            nus.append(nu.ravel())
        done += m
        if log:
            log("%s kind %d: %d / %d curves, %d root-count mismatches, max dc %.2e" % (family, kind, done, curves, r["nfound_mismatch"], r["dc_max"]))
    dc = np.concatenate(dcs) if dcs else np.zeros(0); du = np.concatenate(dus) if dus else np.zeros(0)
    r["dc_median"] = _q(dc, 0.5); r["dc_p999"] = _q(dc, 0.999)
    r["du_median"] = _q(du, 0.5); r["du_p999"] = _q(du, 0.999)
    r["du_frac_gt_1e4"] = r["du_gt_1e4"] / max(1, r["n_eval_cmp"])
    if nus:
        nu = np.concatenate(nus)
        r["noise_du_frac_gt_1e4"] = r["noise_du_gt_1e4"] / max(1, r["noise_n"]); r["noise_du_p999"] = _q(nu, 0.999)
    return r


def main():
    curves = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "parity_report.json")
    only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
    solver = api.DispersionSolver("cuda:0")
    # the chunks of 16384 curves take the three-launch root search of large batches (fast-path launch + hand-over)
    solver.lib.surfdisp_set_split_min_models(8192)
    rep = {"curves_per_family_and_wave_type": curves, "host_threads": os.cpu_count(), "split_min_models": 8192, "results": {}}
    t0 = time.time()
    for fam in FAMILIES:
        for kind in (2, 1):
            if only and "%s_kind%d" % (fam, kind) not in only:
                continue
            r = hunt(solver, fam, kind, curves, log=lambda s: print(s, "(%.0f s)" % (time.time() - t0), flush=True))
            rep["results"]["%s_kind%d" % (fam, kind)] = r
            with open(out, "w") as f:
                json.dump(rep, f, indent=1)
    rep["seconds"] = time.time() - t0
    tot = sum(r["curves"] for r in rep["results"].values()); bad = sum(r["nfound_mismatch"] for r in rep["results"].values())
    rep["total_curves"] = tot; rep["total_nfound_mismatch"] = bad
    rep["total_unexplained_mismatch"] = sum(r["unexplained_mismatch"] for r in rep["results"].values())
    rep["reference_own_nfound_noise"] = "oracle float32 solver vs float64 solver on the same float32 models: %d of %d curves" % (
        sum(r.get("noise_nfound_mismatch", 0) for r in rep["results"].values()), sum(min(4096, r["curves"]) for r in rep["results"].values()))
    rep["total_evaluations"] = sum(r["evaluations"] for r in rep["results"].values())
    rep["dc_max"] = max(r["dc_max"] for r in rep["results"].values())
    with open(out, "w") as f:
        json.dump(rep, f, indent=1)
    print("TOTAL %d curves, %d root-count mismatches, max dc %.2e, %.0f s" % (tot, bad, rep["dc_max"], rep["seconds"]))


if __name__ == "__main__":
    main()
