"""Regenerates tests/golden/thermal_reference.json by RUNNING the reference's thermal mantle parameterisation
(layers.OceanMantleHybrid, layers.py:297-363, with ThermSeis.HSCM / OceanSeisRitz / OceanSeisRuan) inside ocean models
built by the reference's buildModel1D (config 1, point.py:374-391): full stacks (seisPropLayers with the reference
mantle) and the CascadiaOcean.isgood verdicts.  Triforce.mathPlus.logQuad (used only by OceanSeisJack) is stubbed.
Build container only."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
OUT = os.path.join(HERE, "thermal_reference.json")


def main():
    import make_golden_point as GP
    brown, layers, utils, models, point = GP.import_point()
    sys.modules["Triforce.mathPlus"].logQuad = lambda *a, **k: 0.0
    rng = np.random.default_rng(31)
    out = {"stacks": []}
    for trial in range(24):
        age = float(rng.uniform(0.3, 14.0)); Hw = float(rng.uniform(0.5, 4.0)); Hs = float(rng.uniform(0.02, 1.5))
        Hc = float(rng.uniform(5.0, 9.0))
        scale = 0.05 if trial < 12 else 0.3          # small perturbations mostly pass the priors, large ones do not
        coefs = [float(x) for x in rng.uniform(-scale, scale, 4)]
        info = {"modelType": "CascadiaOcean", "period": float(rng.choice([1, 10, 50])), "refLayer": True, "topo": -Hw}
        if trial % 3 == 0:
            info["lithoAgeQ"] = True; info["lithoAge"] = float(rng.uniform(0.5, 10.0))
        s = {"OceanWater": {"H": Hw}, "OceanSedimentCascadia": {"H": Hs}, "OceanCrust": {"H": Hc, "Vs": [[3.25, "fixed"], [3.94, "fixed"]]},
             "OceanMantleHybrid": {"BottomDepth": 200, "Conversion": "Ritzwoller", "ThermAge": age, "Vs": [[c, "fixed"] for c in coefs]},
             "Info": info}
        if trial % 4 == 1:
            s["OceanMantleHybrid"]["Tp"] = float(rng.uniform(1280, 1380))
        mod = models.buildModel1D(json.loads(json.dumps(s)))
        h, vs, vp, rho, qs, qp, grp = mod.seisPropLayers(refLayer=True)
        out["stacks"].append({"setting": s, "h": h.tolist(), "vs": vs.tolist(), "vp": vp.tolist(), "rho": rho.tolist(), "qs": qs.tolist(),
                              "isgood": bool(mod.isgood())})
    with open(OUT, "w") as f:
        json.dump(out, f)
    print("wrote", OUT, len(out["stacks"]), "stacks,", sum(c["isgood"] for c in out["stacks"]), "admissible")


if __name__ == "__main__":
    main()
