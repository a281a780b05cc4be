"""Regenerates tests/golden/layers_reference.json by RUNNING the reference's own layer classes
(/root/reference/layers.py imports cleanly here) and the stack assembly of reference models.py:72-102
(Model1D.seisPropGrids / seisPropLayers).  models.py itself cannot be imported (Triforce, matplotlib ...
are absent), so Model1D is imported with those third-party modules stubbed -- only plotting helpers live
there; the arithmetic that produces the fixtures is the reference's.  Build container only.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "layers_reference.json")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def import_reference():
    pkg = types.ModuleType("pySurfInv"); pkg.__path__ = [REF]; sys.modules["pySurfInv"] = pkg
    for stub in ("Triforce", "Triforce.pltHead", "Triforce.utils", "Triforce.obspyPlus", "Triforce.mathPlus",
                 "netCDF4", "geographiclib", "geographiclib.geodesic", "matplotlib", "matplotlib.pyplot"):
        m = types.ModuleType(stub); m.__all__ = []; sys.modules.setdefault(stub, m)
    sys.modules["netCDF4"].Dataset = object
    sys.modules["geographiclib.geodesic"].Geodesic = object
    sys.modules["Triforce.utils"].GeoGrid = object; sys.modules["Triforce.utils"].GeoMap = object
    brown = _load("pySurfInv.brownian", os.path.join(REF, "brownian.py"))
    layers = _load("pySurfInv.layers", os.path.join(REF, "layers.py"))
    utils = _load("pySurfInv.utils", os.path.join(REF, "utils.py"))
    models = _load("pySurfInv.models", os.path.join(REF, "models.py"))
    return brown, layers, utils, models


def main():
    brown, layers, utils, models = import_reference()
    out = {"bspline": [], "stacks": []}
    # B-spline bases for every (N, nBasis) the layer classes can ask for (layers.py:161-173, 226, 243-255)
    for N in (2, 3, 4, 5, 7, 10, 15, 30, 60):
        for nb in (1, 2, 3, 4, 5, 6, 8):
            if nb > N + 1:
                continue
            z = np.linspace(0, 37.5, N + 1)
            out["bspline"].append({"N": N, "nbasis": nb, "basis": layers.BsplBasis(z, nb).basis.tolist()})
    settings = [
        # continental: sediment + crust + mantle (config 2 family)
        {"Sediment": {"H": 2.5, "Vs": 1.8}, "Crust": {"H": 33.0, "Vs": [3.3, 3.5, 3.7, 3.9]},
         "Mantle": {"BottomDepth": 200.0, "Vs": [4.4, 4.3, 4.35, 4.5, 4.6]}, "Info": {}},
        {"Sediment": {"H": 0.8, "Vs": [1.2, 2.0]}, "Crust": {"H": 18.0, "Vs": [3.2, 3.6, 3.8, 3.95]},
         "Mantle": {"BottomDepth": 160.0, "Vs": [4.2, 4.5, 4.4]}, "Info": {"refLayer": True}},
        {"Sediment": {"H": 4.0, "Vs": 2.4}, "Crust": {"H": 62.0, "Vs": [3.4, 3.5, 3.9]},
         "Mantle": {"BottomDepth": 230.0, "Vs": [4.5, 4.1, 4.7, 4.3, 4.6, 4.4]}, "Info": {}},
        # oceanic (point.py:374-391 family, thermal mantle replaced by the B-spline OceanMantle)
        {"OceanWater": {"H": 2.7}, "OceanSedimentCascadia": {"H": 0.35}, "OceanCrust": {"H": 7.0, "Vs": [3.25, 3.94]},
         "OceanMantle": {"BottomDepth": 200.0, "Vs": [4.4, 4.2, 4.1, 4.3, 4.5]}, "Info": {"refLayer": True, "topo": -2.7}},
        {"OceanWater": {"H": 1.1}, "OceanSediment": {"H": 1.5, "Vs": 0.9}, "OceanCrust": {"H": 5.2, "Vs": 3.6},
         "OceanMantle": {"BottomDepth": 120.0, "Vs": [4.3, 4.15, 4.45, 4.55]}, "Info": {"topo": -1.1}},
        # a vanishing sediment group (H below the 0.01 km limit of models.py:82)
        {"Sediment": {"H": 0.005, "Vs": 1.5}, "Crust": {"H": 9.0, "Vs": [3.3, 3.7]},
         "Mantle": {"BottomDepth": 80.0, "Vs": [4.4, 4.5]}, "Info": {}},
    ]
    def fix(v):  # a bare 4-number list would be read as a BrownianVar spec [v, vmin, vmax, step] (layers.py:592)
        return [[x, "fixed"] for x in v] if isinstance(v, list) else v
    for s in settings:
        s2 = json.loads(json.dumps(s))
        for k, parm in s2.items():
            if k != "Info" and "Vs" in parm:
                parm["Vs"] = fix(parm["Vs"])
        mod = models.buildModel1D(s2)
        ref = bool(s["Info"].get("refLayer", False))
        h, vs, vp, rho, qs, qp, grp = mod.seisPropLayers(refLayer=ref)
        z, gvs, gvp, grho, gqs, gqp, ggrp = mod.seisPropGrids(refLayer=ref)
        out["stacks"].append({"setting": s2, "refLayer": ref, "h": h.tolist(), "vs": vs.tolist(), "vp": vp.tolist(),
                              "rho": rho.tolist(), "qs": qs.tolist(), "groups": list(grp),
                              "grid_z": z.tolist(), "grid_vs": gvs.tolist()})
    # prior checks: the reference's own CascadiaPrism.isgood (models.py:294-360) on perturbed models
    out["priors"] = []
    base = {"Sediment": {"H": 2.0, "Vs": [1.2, 2.0]}, "Crust": {"H": 30.0, "Vs": [3.3, 3.5, 3.7, 3.9]},
            "Mantle": {"BottomDepth": 200.0, "Vs": [4.4, 4.3, 4.5, 4.4, 4.6]}, "Info": {"modelType": "MCInv"}}
    rng = np.random.default_rng(7)
    for trial in range(60):
        s2 = json.loads(json.dumps(base))
        s2["Sediment"]["H"] = float(rng.uniform(0.5, 4.0))
        s2["Sediment"]["Vs"] = [float(x) for x in rng.uniform(0.9, 2.6, 2)]
        s2["Crust"]["H"] = float(rng.uniform(18.0, 45.0))
        s2["Crust"]["Vs"] = [float(x) for x in np.sort(rng.uniform(3.1, 4.1, 4)) + rng.normal(0, 0.08, 4)]
        s2["Mantle"]["Vs"] = [float(x) for x in rng.uniform(4.0, 5.0, 5)]
        s3 = json.loads(json.dumps(s2))
        for k, parm in s3.items():
            if k != "Info" and "Vs" in parm:
                parm["Vs"] = fix(parm["Vs"])
        # CascadiaPrism cannot be instantiated through buildModel1D (its _loadLocalInfo refers to an undefined
        # name, models.py:279); its isgood only needs seisPropGrids, so it is applied to the MCinv model
        mod = models.buildModel1D(s3)
        out["priors"].append({"setting": s3, "isgood": bool(models.CascadiaPrism.isgood(mod))})
    with open(OUT, "w") as f:
        json.dump(out, f)
    print("wrote", OUT, len(out["bspline"]), "bases,", len(out["stacks"]), "stacks,", len(out["priors"]), "prior cases,",
          sum(p["isgood"] for p in out["priors"]), "good")


if __name__ == "__main__":
    main()
