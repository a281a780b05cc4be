"""Stack templates: the reference's layered-model description (YAML/dict settings of reference
models.py:41-51 + layers.py:573-604) compiled into a fixed-size record the device-side model builder
(``surfdisp_build_stacks``) consumes, plus the free-parameter table of the Monte-Carlo walk.

A setting is an ordered dict ``{layerType: parm, ..., 'Info': {...}}`` exactly like the reference's.  Every
numeric entry written as a Brownian spec -- ``[v, 'abs'|'abs_pos'|'rel'|'rel_pos', width, step]`` or
``[v, vmin, vmax, step]`` (layers.py:583-598) -- becomes one free parameter; the parameter order is the order
of ``MCinv._brownians`` (models.py:227-240), i.e. the column order of ``mcTrack`` rows (point.py:57,73).
Supported layer types: Sediment, Crust, Mantle/OceanMantle, OceanWater, OceanSediment, OceanCrust,
OceanSedimentCascadia, OceanMantleHybrid (the thermal mantle, 'Ritzwoller' conversion), ReferenceMantle (also
through ``Info.refLayer``).  Parameters of a layer the builder does not implement (Crust 'Gauss', OceanMantle
'deg', other thermal conversions) raise instead of being ignored.
"""
import ctypes as C

import numpy as np

MAX_GROUPS = 8
MAX_COEF = 8

# group kinds (Vs profile)                       reference class
G_WATER, G_CONST, G_LINEAR, G_BSPLINE, G_CASCADIA, G_REFMANTLE, G_HYBRID = 0, 1, 2, 3, 4, 5, 6
# fine-layer rules
N_FIXED, N_CRUST, N_OCRUST = 0, 1, 2
# density rules
R_QUARTIC, R_OCEAN, R_MANTLE, R_CONST = 0, 1, 2, 3
# group classes (layer.prop['Group']) and prior rules (CascadiaPrism.isgood, reference models.py:294-360)
C_WATER, C_SEDIMENT, C_CRUST, C_MANTLE, C_OTHER = 0, 1, 2, 3, 4
P_JUMP, P_VSMAX, P_MONO, P_BOTTOM = 1, 2, 4, 8
P_ALL = 15
# CascadiaOcean.isgood (models.py:571-677): sediment Vs >= 0.2, first grid pair not decreasing (what is left of the
# jump rule there), oscillation limit, no local maximum in the mantle, no extreme decrease below the moho, wavelet rule
P_SEDMIN, P_FIRSTPAIR, P_OSCI, P_LOCALMAX, P_SLOPE, P_CWT = 16, 32, 64, 128, 256, 512
# the rule sets of the reference's model classes
PRIOR_PRISM = P_JUMP | P_VSMAX | P_MONO | P_BOTTOM                    # CascadiaPrism.isgood      models.py:294-360
PRIOR_CONTINENT = P_JUMP | P_VSMAX | P_MONO                           # CascadiaContinent.isgood  models.py:385-523
PRIOR_OCEAN = P_SEDMIN | P_FIRSTPAIR | P_BOTTOM | P_OSCI | P_LOCALMAX | P_SLOPE | P_CWT   # CascadiaOcean.isgood :571-677
PRIOR_OF_MODELTYPE = {"CascadiaPrism": PRIOR_PRISM, "CascadiaContinent": PRIOR_CONTINENT, "CascadiaOcean": PRIOR_OCEAN,
                      "MCInv": 0, "General": 0}
_CLASS = {'Sediment': C_SEDIMENT, 'Crust': C_CRUST, 'Mantle': C_MANTLE, 'OceanMantle': C_MANTLE, 'OceanMantleHybrid': C_MANTLE, 'OceanWater': C_WATER,
          'OceanSediment': C_SEDIMENT, 'OceanSedimentCascadia': C_SEDIMENT, 'OceanCrust': C_CRUST}


class Group(C.Structure):
    _fields_ = [("kind", C.c_int), ("nfine_rule", C.c_int), ("nfine", C.c_int), ("h_mode", C.c_int),
                ("h_param", C.c_int), ("ncoef", C.c_int), ("rho_rule", C.c_int), ("gclass", C.c_int),
                ("v_param", C.c_int * MAX_COEF), ("v_fixed", C.c_double * MAX_COEF),
                ("h_fixed", C.c_double), ("vp_a", C.c_double), ("vp_b", C.c_double), ("rho_const", C.c_double),
                ("qs", C.c_double), ("slope", C.c_double),
                ("age_param", C.c_int), ("pad_", C.c_int), ("age_fixed", C.c_double), ("tp", C.c_double),
                ("period", C.c_double), ("q_age", C.c_double)]


class StackTemplateC(C.Structure):
    _fields_ = [("ngroups", C.c_int), ("nparams", C.c_int), ("prior_mask", C.c_int), ("_pad", C.c_int),
                ("topo", C.c_double), ("groups", Group * MAX_GROUPS)]


_TYPES = {
    # name: (kind, nfine_rule, nfine, vp_a, vp_b, rho_rule, rho_const, qs)
    "Sediment": (None, N_FIXED, 1, 2.0, 0.0, R_QUARTIC, 0.0, 80.0),                    # layers.py:139-156
    "Crust": (G_BSPLINE, N_CRUST, 0, 1.80, 0.0, R_QUARTIC, 0.0, 600.0),                # layers.py:158-189
    "Mantle": (G_BSPLINE, N_CRUST, 0, 1.76, 0.0, R_MANTLE, 0.0, 150.0),                # layers.py:239-265
    "OceanMantle": (G_BSPLINE, N_CRUST, 0, 1.76, 0.0, R_MANTLE, 0.0, 150.0),
    "OceanMantleHybrid": (G_HYBRID, N_CRUST, 0, 1.76, 0.0, R_MANTLE, 0.0, 150.0),    # layers.py:297-363 (Qs from the thermal model)
    "OceanWater": (G_WATER, N_FIXED, 1, 0.0, 1.475, R_CONST, 1.027, 10000.0),          # layers.py:191-204
    "OceanSediment": (G_CONST, N_FIXED, 1, 1.23, 1.28, R_OCEAN, 0.0, 80.0),            # layers.py:206-219
    "OceanSedimentCascadia": (G_CASCADIA, N_FIXED, 1, 1.23, 1.28, R_OCEAN, 0.0, 80.0), # layers.py:289-295
    "OceanCrust": (None, N_OCRUST, 0, 1.8, 0.0, R_OCEAN, 0.0, 350.0),                  # layers.py:221-237
}


def _is_spec(v):
    if isinstance(v, (list, tuple)) and len(v) >= 2:
        if v[1] in ("fixed", "total", "abs", "abs_pos", "rel", "rel_pos"):
            return True
        if len(v) == 4 and all(isinstance(x, (int, float)) for x in v):
            return True
    return False


class Param:
    """One free parameter with the bounds/step of BrownianVar / BrownianVarMC (brownian.py:3-68)."""

    def __init__(self, spec, where):
        v = float(spec[0])
        if spec[1] in ("abs", "abs_pos", "rel", "rel_pos"):
            w = float(spec[2])
            lo, hi = (v - w, v + w) if spec[1].startswith("abs") else (v * (1 - w / 100), v * (1 + w / 100))
            if spec[1].endswith("_pos"):
                lo, hi = max(lo, 0.0), max(hi, 0.0)
            step = float(spec[3])
        else:
            lo, hi, step = float(spec[1]), float(spec[2]), float(spec[3])
        self.v0, self.vmin, self.vmax = v, lo, hi
        self.step = abs(hi - lo) / 2 if step > abs(hi - lo) / 2 else step
        self.where = where

    def __repr__(self):
        return "Param(%s v=%g [%g,%g] step=%g)" % (self.where, self.v0, self.vmin, self.vmax, self.step)


class StackTemplate:
    def __init__(self, setting, prior_mask=0):
        self.prior_mask = int(prior_mask)   # which prior rules proposals must satisfy (0 = MCinv.isgood: none)
        self.params = []
        self.groups = []
        info = dict(setting.get("Info", {}))
        self.topo = float(info.get("topo", 0.0))
        self._info = info
        for name, parm in setting.items():
            if name == "Info":
                continue
            if name == "ReferenceMantle" or name == "HalfSpace":
                # ReferenceMantle (layers.py:267-285): linear continuation of the deepest values, 20 fine layers.
                # 'HalfSpace' (an extension, not a reference class): the same with ONE layer -- the explicit closing
                # layer of the config-2 stacks (synth.crustal_models appends the same).
                g = self._blank()
                g.kind, g.nfine_rule, g.nfine, g.h_mode, g.h_param = G_REFMANTLE, N_FIXED, (20 if name == "ReferenceMantle" else 1), 0, -1
                g.h_fixed = float(parm.get("H", 300.0 if name == "ReferenceMantle" else 10.0))
                g.vp_a, g.vp_b, g.rho_rule, g.qs, g.slope = 1.76, 0.0, R_MANTLE, 150.0, float(parm.get("Slope", 0.0))
                g.gclass = C_MANTLE
                unknown = set(parm) - {"H", "Slope"}
                if unknown:
                    raise ValueError("%s: unsupported parameters %s" % (name, sorted(unknown)))
                self.groups.append(g)
                continue
            if name not in _TYPES:
                raise ValueError("layer type %r is not supported by the device-side builder" % name)
            self.groups.append(self._group(name, dict(parm)))
        if info.get("refLayer", False):     # models.py:111-113: H = 300 km, slope 0.35/200
            g = self._blank()
            g.kind, g.nfine_rule, g.nfine, g.h_mode, g.h_param, g.h_fixed = G_REFMANTLE, N_FIXED, 20, 0, -1, 300.0
            g.vp_a, g.vp_b, g.rho_rule, g.qs, g.slope = 1.76, 0.0, R_MANTLE, 150.0, 0.35 / 200
            g.gclass = C_MANTLE
            self.groups.append(g)
        if len(self.groups) > MAX_GROUPS:
            raise ValueError("at most %d layer groups" % MAX_GROUPS)

    @staticmethod
    def _blank():
        g = Group()
        for i in range(MAX_COEF):
            g.v_param[i] = -1
        g.h_param = -1
        g.age_param, g.tp, g.period, g.q_age = -1, 1325.0, 1.0, -1.0
        return g

    def _value(self, v, where):
        """Returns (param_index, fixed_value)."""
        if _is_spec(v):
            if v[1] in ("fixed", "total"):
                return -1, float(v[0])
            self.params.append(Param(v, where))
            return len(self.params) - 1, float(v[0])
        return -1, float(v)

    def _group(self, name, parm):
        kind, nrule, nfine, vp_a, vp_b, rrule, rconst, qs = _TYPES[name]
        g = self._blank()
        g.nfine_rule, g.nfine, g.vp_a, g.vp_b, g.rho_rule, g.rho_const, g.qs = nrule, nfine, vp_a, vp_b, rrule, rconst, qs
        g.gclass = _CLASS[name]
        # parameter order follows the parm dict order, like MCinv._brownians
        for key, val in parm.items():
            if key in ("H", "BottomDepth"):
                g.h_mode = 0 if key == "H" else 1
                g.h_param, g.h_fixed = self._value(val, "%s.%s" % (name, key))
            elif key == "Vs":
                vals = val if (isinstance(val, (list, tuple)) and not _is_spec(val)) else [val]
                if len(vals) > MAX_COEF:
                    raise ValueError("at most %d Vs coefficients per group" % MAX_COEF)
                g.ncoef = len(vals)
                for i, x in enumerate(vals):
                    g.v_param[i], g.v_fixed[i] = self._value(x, "%s.Vs[%d]" % (name, i))
            elif name == "OceanMantleHybrid" and key in ("ThermAge", "Tp", "Conversion"):
                if key == "ThermAge":
                    g.age_param, g.age_fixed = self._value(val, "%s.ThermAge" % name)
                elif key == "Tp":
                    if _is_spec(val) and val[1] not in ("fixed", "total"):
                        raise ValueError("OceanMantleHybrid: Tp must be fixed")
                    g.tp = float(val[0] if _is_spec(val) else val)
                elif val != "Ritzwoller":
                    raise ValueError("OceanMantleHybrid: only the 'Ritzwoller' conversion is built on the device")
            else:
                # (Crust 'Gauss', OceanMantle 'deg', ... change the Vs profile in the reference, layers.py:176-183, 256:
                # dropping them silently would build a different model)
                raise ValueError("%s: parameter %r is not supported by the device-side builder" % (name, key))
        if kind == G_HYBRID:
            # (the perturbation has len(Vs) + 1 B-spline coefficients, the first one 0: layers.py:323, 340)
            if g.ncoef + 1 > MAX_COEF or g.ncoef < 2:
                raise ValueError("OceanMantleHybrid: 2 .. %d Vs coefficients" % (MAX_COEF - 1))
            info = self._info
            g.period = float(info.get("period", 1))
            g.q_age = float(info["lithoAge"]) if (info.get("lithoAgeQ", False) and info.get("lithoAge") is not None) else -1.0
            if "ThermAge" not in parm:
                raise ValueError("OceanMantleHybrid needs ThermAge")
        if kind is None:   # Sediment / OceanCrust: constant or linear (layers.py:146-149, 228-231)
            kind = G_LINEAR if g.ncoef == 2 else G_CONST
            if g.ncoef not in (1, 2):
                raise ValueError("%s takes one Vs or [top, bottom]" % name)
        g.kind = kind
        return g

    @property
    def nparams(self):
        return len(self.params)

    def to_c(self):
        t = StackTemplateC()
        t.ngroups, t.nparams, t.topo, t.prior_mask = len(self.groups), self.nparams, self.topo, self.prior_mask
        for i, g in enumerate(self.groups):
            t.groups[i] = g
        return t

    def start_values(self):
        return np.array([p.v0 for p in self.params], dtype=np.float32)

    def bounds(self):
        return (np.array([p.vmin for p in self.params], np.float32), np.array([p.vmax for p in self.params], np.float32),
                np.array([p.step for p in self.params], np.float32))

    def max_layers(self, lo=None, hi=None):
        """Upper bound of the layer count over the parameter box [lo, hi] (default: the template's own bounds).
        The fine-layer rules grow with the group thickness, so the bound follows from the largest thickness each
        group can take: the array stride of the stacks and the shared-memory record of the root search are sized
        by it (a loose bound halves the resident CTAs of phase 1).  A stack that still overflows it is flagged by
        the builder (status < 0), never truncated."""
        tlo, thi, _ = self.bounds() if self.params else (np.zeros(0), np.zeros(0), None)
        lo = tlo if lo is None else np.asarray(lo, np.float64).reshape(-1, self.nparams).min(axis=0)
        hi = thi if hi is None else np.asarray(hi, np.float64).reshape(-1, self.nparams).max(axis=0)
        n, zlo, zhi, have = 0, -max(self.topo, 0.0), -max(self.topo, 0.0), False   # (models.py:76-77)
        for g in self.groups:
            h_lo, h_hi = (float(lo[g.h_param]), float(hi[g.h_param])) if g.h_param >= 0 else (g.h_fixed, g.h_fixed)
            if g.h_mode == 1 and have:     # BottomDepth: what is left under the groups above
                h_lo, h_hi = max(h_lo - zhi, 0.0), max(h_hi - zlo, 0.0)
            if g.nfine_rule == N_CRUST:
                n += 60 if h_hi >= 150.0 else 30 if h_hi > 60.0 else 15 if h_hi > 20.0 else 10 if h_hi > 10.0 else 5
            elif g.nfine_rule == N_OCRUST:
                n += min(max(int(np.rint(h_hi / 2.0)), 2), 10)
            else:
                n += g.nfine
            zlo, zhi, have = zlo + max(h_lo, 0.0), zhi + max(h_hi, 0.0), True
        return n


def config2_template():
    """BASELINE config 2 as a setting of the reference's own layer classes: sediment (1 layer), crust
    (cubic B-spline, 4 coefficients, 15 fine layers for 20 < H <= 60 km), mantle to 200 km (5 coefficients, 60 fine
    layers) and a closing half-space with the deepest grid values: n = 77.  Free parameters (12, in the order of
    MCinv._brownians): sediment H and Vs, crust H and 4 coefficients, 5 mantle coefficients."""
    setting = {"Sediment": {"H": [2.25, 0.5, 4.0, 0.1], "Vs": [1.75, 1.0, 2.5, 0.05]},
               "Crust": {"H": [32.5, 20.0001, 45.0, 1.0], "Vs": [[3.3, 3.2, 4.0, 0.02], [3.5, 3.2, 4.0, 0.02],
                                                                [3.7, 3.2, 4.0, 0.02], [3.9, 3.2, 4.0, 0.02]]},
               "Mantle": {"BottomDepth": 200.0, "Vs": [[4.4, 4.1, 4.7, 0.02]] * 5},
               "HalfSpace": {"H": 10.0}, "Info": {}}
    return StackTemplate(setting), setting


def config2_params(M, seed=20261018):
    """M random parameter vectors of config 2, float32 [M, 12]: uniform in the boxes of SURVEY 8d, the crustal
    coefficients sorted (monotone crust)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p = np.empty((M, 12), np.float32)
    p[:, 0] = rng.uniform(0.5, 4.0, M); p[:, 1] = rng.uniform(1.0, 2.5, M); p[:, 2] = rng.uniform(20.0001, 45.0, M)
    p[:, 3:7] = np.sort(rng.uniform(3.2, 4.0, (M, 4)), axis=1)
    p[:, 7:12] = rng.uniform(4.1, 4.7, (M, 5))
    return p
