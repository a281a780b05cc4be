"""Regenerates tests/golden/test1_senskernel.json from the reference tree (run in the build container only).

Source: /root/reference/senskernel-1.0/TEST1/{eus_model,test.R.phv,test.R.grv,test.L.phv,test.L.grv}
(pre-computed outputs of the real*8 sibling program SURF_PERTURB, run as KERNELS.csh:24 says:
 modes 0-1, T = 10..100 s step 10, flags -a -f).  These are the only golden vectors in the reference.
"""
import json
import os
import numpy as np

SRC = "/root/reference/senskernel-1.0/TEST1"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "test1_senskernel.json")


def main():
    m = np.loadtxt(os.path.join(SRC, "eus_model"))
    g = {"source": "senskernel-1.0/TEST1 (SURF_PERTURB, real*8, modes 0-1, -a -f)",
         "model_columns": ["h_km", "vp", "vs", "rho", "Qs"],
         "model": m.tolist(), "periods": list(np.arange(10.0, 101.0, 10.0))}
    for w in ("R", "L"):
        phv = np.loadtxt(os.path.join(SRC, "test.%s.phv" % w))
        grv = np.loadtxt(os.path.join(SRC, "test.%s.grv" % w))
        assert phv.shape == (20, 3) and grv.shape == (20, 2)
        g[w] = {"c": [phv[:10, 1].tolist(), phv[10:, 1].tolist()],
                "cvar": [phv[:10, 2].tolist(), phv[10:, 2].tolist()],
                "u": [grv[:10, 1].tolist(), grv[10:, 1].tolist()]}
    # scalars quoted in TEST1/test.R lines 3-4 (T = 10 s, fundamental Rayleigh)
    g["R_T10_extra"] = {"ellipticity": 1.038500, "I0": 45.21353, "I1": 546.7788, "I2": -34.64232, "I3": 9.003018}
    with open(OUT, "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
