"""Parity statistics of the CUDA path against the CPU oracle (run on the GPU box).
Writes gpurun_out/parity_report.json; copy the summary into profiles/ when it is the one to be judged."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from pysurfinv_b200 import api, synth  # noqa: E402


def stats(x):
    x = np.asarray(x).ravel()
    return {"max": float(x.max()), "p999": float(np.quantile(x, 0.999)), "p99": float(np.quantile(x, 0.99)),
            "median": float(np.median(x)), "mean": float(x.mean()), "frac_gt_1e-4": float((x > 1e-4).mean()),
            "frac_gt_1e-5": float((x > 1e-5).mean())}


def main():
    import torch
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    solver = api.DispersionSolver("cuda:0")
    rep = {}
    sets = {"crustal77": (synth.crustal_models(M, seed=101), synth.log_periods()),
            "hand4": (synth.hand_models(M, seed=102), synth.log_periods(24, 6.0, 60.0)),
            "ragged_water": (synth.ragged_models(M // 2, seed=103), synth.log_periods(18, 10.0, 80.0))}
    for name, ((lay, nl), per) in sets.items():
        for kind in (2, 1):
            out = solver.forward(torch.from_numpy(lay).cuda(), torch.from_numpy(nl).cuda(), per, kind=kind)
            g = {k: v.cpu().numpy() for k, v in out.items()}
            nth = os.cpu_count() or 1
            c0, u0, nf0, st0 = O.forward_batch(kind, lay, nl, per, opts=O.make_opts(precision=0), nthreads=nth)
            c1, u1, nf1, st1 = O.forward_batch(kind, lay, nl, per, opts=O.make_opts(precision=1), nthreads=nth)
            ok = (st0 != 3)
            r = {"models": int(lay.shape[1]), "periods": int(len(per)), "lstop_excluded": int((~ok).sum()),
                 "nfound_mismatch_vs_f32_oracle": int((g["nfound"][ok] != nf0[ok]).sum()),
                 "nfound_mismatch_f32_vs_f64_oracle": int((nf0 != nf1).sum()),
                 "full_curves": int((nf0 == len(per)).sum())}
            same = ok & (g["nfound"] == nf0) & (nf0 == nf1)
            r["dc_gpu_vs_oracle_f32"] = stats(np.abs(g["c"] - c0)[same])
            r["du_gpu_vs_oracle_f32"] = stats(np.abs(g["u"] - u0)[same])
            r["dc_oracle_f32_vs_f64solver"] = stats(np.abs(c0 - c1)[same])
            r["du_oracle_f32_vs_f64solver"] = stats(np.abs(u0 - u1)[same])
            r["dc_gpu_vs_oracle_f64solver"] = stats(np.abs(g["c"] - c1)[same])
            r["du_gpu_vs_oracle_f64solver"] = stats(np.abs(g["u"] - u1)[same])
            du = np.abs(g["u"] - u0); du[~same] = 0
            worst = np.dstack(np.unravel_index(np.argsort(du.ravel())[-5:], du.shape))[0]
            r["worst_du"] = [{"model": int(i), "T": float(per[k]), "c_gpu": float(g["c"][i, k]), "c_f32": float(c0[i, k]),
                              "u_gpu": float(g["u"][i, k]), "u_f32": float(u0[i, k]), "u_f64solver": float(u1[i, k])}
                             for i, k in worst]
            rep["%s_kind%d" % (name, kind)] = r
            print(name, kind, json.dumps({k: v for k, v in r.items() if k != "worst_du"}), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w") as f:
        json.dump(rep, f, indent=1)


if __name__ == "__main__":
    main()
