#!/usr/bin/env python
"""bench.py -- dispersion evaluations per second (BASELINE.json metric).

Default workload = BASELINE config 2: batched forward sweep, 1 Mi random sediment + crust + mantle models (n = 77
layers, described through the reference's own layer classes: pysurfinv_b200.stack.config2_template) x 40 periods
(8-80 s), Rayleigh phase + group velocity, per GPU (weak scaling: every rank solves its own 1 Mi models).

    python bench.py --gpus N --steps K --warmup W            (torchrun launches N ranks for N > 1)
    python bench.py --impl reference ...                      (CPU oracle port on the host cores)
    python bench.py --workload mc   [--chains 256] [--love]   (configs 1 / 3: Monte-Carlo chains of one point)
    python bench.py --workload mc --thermal [--chains 4096]   (config 4: the ocean model with the thermal mantle, 10-150 s)
    python bench.py --workload grid [--points 2000 --chains 16]   (config 5: points sharded over the ranks,
                                                               chain rows gathered inside the timed region)

One JSON line on stdout (rank 0).  1 evaluation = one (model, period) producing c (and U where the workload says so).
  value      device-resident: the stacks are in HBM when the timed region starts
  e2e        the call a user makes (Model1D.forward for a batch: parameter vectors on the host in, c / U / nfound /
             flags on the host out), pinned host buffers, copies inside the timed region
  e2e_layers the same through the fast_surf-level call (layer arrays on the host in)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "dispersion evals/sec (models x periods, Rayleigh c+U)"
UNIT = "evals/s"
# FLOP-equivalents per unit of work, SURVEY.md 8(d): Rayleigh layer-step 150 FLOP + 4 transcendentals
# (10 FLOP each), Love layer-step 20 FLOP + 2 transcendentals, REIGEN sub-layer 1800 FLOP (fp64).
F_R, F_L, F_U = 190.0, 40.0, 1800.0
# FLOP the group-velocity kernel executes per sub-layer (float32 state, packed pairs, one sub-layer per layer as in
# config 2): 71 FFMA + 64 FFMA2 + 26 FMUL2 + 8 FMUL + 6 FADD in the sub-layer body, 26 FFMA + 87 FMUL + 13 FADD of set-up
# per layer (SASS of the loop; the ncu counters of profiles/r2_ncu_phase2.txt give the same FFMA : FMUL : FADD mix with a
# packed instruction counted once), against ~1800 of the reference's stage-by-stage form in float64
F_U_EXEC = 616.0
# DRAM bytes per model of the root-search launches and of the group-velocity launch (ncu dram__bytes_read + write at
# 524288 models x 40 periods, profiles/r2_ncu_*.txt); algorithmic bytes per model: phase 1 reads the constants
# 8 x lpad x 4 B at the first period and once more at the start of the later periods and writes c, ratio
DRAM_P1_PER_MODEL, DRAM_P2_PER_MODEL = 5965.0, 3063.0
WORKLOAD_SWEEP = "config2: batched forward sweep, %d random 77-layer sediment+crust+mantle models x %d periods 8-80 s per GPU, Rayleigh phase+group"


def algo_p1_bytes(lpad, K):
    return 2 * 8 * lpad * 4 + 2 * K * 4 + 16


def hbm_roofline(algo_bytes, traffic_bytes, ms):
    """The memory side of the roofline for a kernel: algorithmic bytes per launch over its duration against the copy
    bandwidth the driver measured on this pool (MEASURED_PEAKS.json; the profiling recipe's fallback if absent).  For this
    path it documents that HBM is not the bound."""
    peak, src = 6650.0, "fallback of B200_PROFILING.md"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except (OSError, KeyError, ValueError):
        pass
    ach = algo_bytes / (ms * 1e-3) * 1e-9
    return {"achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": src,
            "traffic": traffic_bytes, "algorithmic_bytes": algo_bytes}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sweep", choices=["sweep", "mc", "grid"],
                    help="sweep: BASELINE config 2 (default, the headline); mc: one ensemble of Monte-Carlo chains on one "
                         "point (configs 1 / 3); grid: points x chains sharded over the ranks with the result gather inside "
                         "the timed region (config 5)")
    ap.add_argument("--models", type=int, default=1 << 20, help="sweep: models per GPU per step")
    ap.add_argument("--periods", type=int, default=40)
    ap.add_argument("--chains", type=int, default=256, help="mc: chains; grid: chains per point")
    ap.add_argument("--points", type=int, default=2000, help="grid: points in total (sharded over the ranks)")
    ap.add_argument("--mc-steps", type=int, default=50, help="mc / grid: Monte-Carlo steps per timed step")
    ap.add_argument("--love", action="store_true", help="mc: joint Rayleigh + Love (config 3: a second ensemble of Love chains)")
    ap.add_argument("--thermal", action="store_true", help="mc / grid: the ocean model with the thermal mantle, periods 10-150 s (config 4)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ workloads
MC_SETTING = {"Sediment": {"H": [2.0, "abs_pos", 1.5, 0.1], "Vs": [[1.2, 0.8, 2.0, 0.05], [2.0, 1.2, 2.8, 0.05]]},
              "Crust": {"H": [30.0, "abs", 12.0, 1.0], "Vs": [[3.3, "rel", 10, 0.02], [3.5, "rel", 10, 0.02],
                                                            [3.7, "rel", 10, 0.02], [3.9, "rel", 10, 0.02]]},
              "Mantle": {"BottomDepth": 200.0, "Vs": [[4.4, "abs", 0.3, 0.02], [4.3, "abs", 0.3, 0.02],
                                                     [4.5, "abs", 0.3, 0.02], [4.4, "abs", 0.3, 0.02], [4.6, "abs", 0.3, 0.02]]},
              "Info": {"refLayer": True, "modelType": "CascadiaPrism"}}
MC_PERIODS = np.array([8, 10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 50, 60, 70, 80], np.float32)   # point.py:400 + 8 s
# config 4: the ocean model of point.py:374-391 with the thermal mantle (OceanMantleHybrid: half-space cooling + OceanSeisRitz +
# OceanSeisRuan, layers.py:297-363), CascadiaOcean prior rules, Rayleigh 10-150 s.  The reference's classes give this setting
# 86 layers (water, sediment, 4 crustal layers, 60 mantle layers, 20 reference-mantle layers).
THERMAL_SETTING = {"OceanWater": {"H": 2.5}, "OceanSedimentCascadia": {"H": [1, "rel_pos", 100, 0.1]},
                   "OceanCrust": {"H": 7, "Vs": [3.25, 3.94]},
                   "OceanMantleHybrid": {"BottomDepth": 200, "Conversion": "Ritzwoller", "ThermAge": [4, "rel_pos", 200, 0.4],
                                         "Vs": [[0, "abs", 0.4, 0.01], [0, "abs", 0.4, 0.01], [0, "abs", 0.4, 0.01], [0, "abs", 0.2, 0.01]]},
                   "Info": {"modelType": "CascadiaOcean", "period": 10, "refLayer": True, "lithoAgeQ": True, "lithoAge": 4.0, "topo": -2.5}}
THERMAL_PERIODS = np.array([10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 50, 60, 70, 80, 100, 125, 150], np.float32)


def p1_launches(n_models):
    """Launches of the root search per batch: first period + later periods, the later periods as a fast-path launch and a
    hand-over launch from 98304 models on (kSplitMinDefault in csrc/surfdisp_kernels.cu)."""
    return 3 if n_models >= 98304 else 2


def mc_case(args):
    """(setting, prior rule set, periods, layer count, label) of the Monte-Carlo workloads."""
    from pysurfinv_b200 import stack as S
    if getattr(args, "thermal", False):
        return THERMAL_SETTING, S.PRIOR_OCEAN, THERMAL_PERIODS, 86, "86-layer ocean model with the thermal mantle (OceanMantleHybrid)", "10-150 s"
    return MC_SETTING, S.PRIOR_PRISM, MC_PERIODS, 96, "96-layer sediment+crust+mantle B-spline model", "8-80 s"


def cpu_sample(per, seconds, nthreads, kind=2, stacks=None, chunk=None, seed=99):
    """Times the CPU oracle (float32-faithful port of fast_surf: the checker, here the reported CPU baseline) on a
    bounded sample of the workload: config-2 stacks assembled by the numpy restatement of the reference's model
    assembly, or the stacks given."""
    from oracle import oracle as O
    if stacks is None:
        from oracle import model_builder as MB
        from pysurfinv_b200 import stack as S
        chunk = chunk or max(64, 16 * nthreads)
        t, _ = S.config2_template()
        lay, nl = MB.build_stacks(t, S.config2_params(chunk, seed=seed), 77)
    else:
        lay, nl = stacks
        chunk = lay.shape[1]
    O.forward_batch(kind, lay[:, :nthreads], nl[:nthreads], per, nthreads=nthreads)  # warm-up (page in, threads)
    done, t0 = 0, time.perf_counter()
    tot = {}
    while True:
        cnt = O.OracleCounters()
        O.forward_batch(kind, lay, nl, per, opts=O.make_opts(precision=0), nthreads=nthreads, counters=cnt)
        for k, v in cnt.as_dict().items():
            tot[k] = tot.get(k, 0) + v
        done += chunk
        dt = time.perf_counter() - t0
        if dt >= seconds:
            break
    return done * len(per) / dt, done, dt, tot


def mc_cpu_stacks(args, n, seed=5):
    """Stacks of the Monte-Carlo workload for the CPU arm: admissible random models of its setting (numpy restatement of
    the reference's model assembly and prior rules)."""
    from oracle import model_builder as MB
    from pysurfinv_b200 import stack as S
    setting, prior, _per, _nl, _label, _rng = mc_case(args)
    t = S.StackTemplate(setting, prior_mask=prior)
    check = MB.priors_ocean if prior == S.PRIOR_OCEAN else MB.priors
    lo, hi, _ = t.bounds()
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        p = (lo + (hi - lo) * rng.random(t.nparams)).astype(np.float32)
        if check(t, p.astype(np.float64)) & prior == 0:
            out.append(p)
    return MB.build_stacks(t, np.array(out), t.max_layers())


def mc_workload_name(args, world):
    _s, _p, per, _nl, label, prange = mc_case(args)
    if args.workload == "grid":
        return ("config5: grid inversion, %d points x %d chains x %d Monte-Carlo steps per timed step, sharded by point over the "
                "GPUs, chain rows all-gathered and best misfit all-reduced inside the timed region; %s, "
                "%d periods %s, Rayleigh phase velocity" % (args.points, args.chains, args.mc_steps, label, len(per), prange))
    cfg = "4" if getattr(args, "thermal", False) else ("3" if args.love else "1")
    return ("config%s: one-point Monte-Carlo inversion, %d chains x %d steps per timed step, %s, %d periods %s, %s phase velocity"
            % (cfg, args.chains, args.mc_steps, label, len(per), prange, "Rayleigh + Love" if args.love else "Rayleigh"))


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port: no Fortran compiler in the image or on the GPU box),
    all host threads, bounded sample per step, on the same workload, metric and unit as the GPU arm."""
    if rank != 0:
        return
    from pysurfinv_b200 import synth
    nthreads = os.cpu_count() or 1
    sec = max(2.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    if args.workload == "sweep":
        per, stacks, kinds = synth.log_periods(args.periods), None, (2,)
        workload = WORKLOAD_SWEEP % (args.models, args.periods)
        metric = METRIC
    else:
        per, stacks = mc_case(args)[2], mc_cpu_stacks(args, max(64, 16 * nthreads))
        kinds = (2, 1) if args.love else (2,)
        workload = mc_workload_name(args, world)
        metric = METRIC.replace("c+U", "c")
    for _ in range(args.warmup):
        cpu_sample(per, 0.5, nthreads, stacks=stacks)
    tot_models, tot_t = 0, 0.0
    for _ in range(args.steps):
        for kind in kinds:
            v, n, dt, _c = cpu_sample(per, sec / len(kinds), nthreads, kind=kind, stacks=stacks)
            tot_models += n; tot_t += dt
    value = tot_models * len(per) / tot_t
    sample = "%d models x %d periods per step (%.1f s) of the same generator" % (tot_models // max(1, args.steps), len(per), sec)
    line = {"impl": "reference", "metric": metric, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "cpu_sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample,
                             "note": "C++ restatement of fast_surf (-O2, no FMA contraction, per-period std::vector set-up): a "
                                     "conservative stand-in for gfortran -O3; context, not a target"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_JSON_FD = None


def _claim_stdout():
    """Everything any library prints to stdout (NCCL's version banner ...) goes to stderr; the saved descriptor is
    kept for the ONE JSON line of the contract."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def pin_to_gpu_numa(local):
    """Binds this rank to the CPUs NVML reports as local to its GPU (its NUMA node): with N ranks on one host the pinned
    staging buffers and the copy threads then sit next to the GPU's PCIe root instead of wherever the scheduler put
    the process.  Returns what was done (for the JSON line)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (words[i // 64] >> (i % 64)) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "cpus %d-%d (%d)" % (cpus[0], cpus[-1], len(cpus))
    except Exception as e:  # noqa: BLE001
        return "not pinned: %s" % type(e).__name__
    return "not pinned"


def run_mc(args, rank, world, local, dev, barrier):
    """--workload mc / grid: ensembles of Metropolis chains, every stage on the device (pysurfinv_b200.mc)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from pysurfinv_b200 import api, mc, stack as S
    from pysurfinv_b200.distributed import shard_range, gather_chain_rows
    solver = api.DispersionSolver(dev)
    setting, prior, per, nl_case, _label, _prange = mc_case(args)
    t = S.StackTemplate(setting, prior_mask=prior)
    K, P = len(per), t.nparams
    grid = args.workload == "grid"
    if grid:
        lo, hi = shard_range(args.points, rank, world)
        npts = hi - lo
        npts_max = (args.points + world - 1) // world
    else:
        lo, npts, npts_max = 0, 1, 1
    cpp = args.chains
    # synthetic observations: the curve of an admissible random model per point (seeded by the global point index)
    start = torch.from_numpy(np.tile(t.start_values(), (npts, 1))).to(dev).contiguous()
    truth = torch.empty_like(start)
    for i in range(npts):   # (one tiny launch per point, outside the timed region)
        truth[i:i + 1] = solver.mc_propose(t, start[i:i + 1], seed=1000 + lo + i, step_index=0,
                                           reset_mask=torch.ones(1, dtype=torch.uint8, device=dev))
    lay, nl = solver.build_stacks(t, truth)
    kinds = (2, 1) if (args.love and not grid) else (2,)
    nsteps = args.mc_steps
    ens = []
    for kind in kinds:
        obs = solver.forward(lay, nl, per, kind=kind, group=False)["c"].cpu().numpy()
        ens.append(mc.ChainEnsemble(solver, t, per, obs, np.full_like(obs, 0.01), n_chains=cpp, seed=1 + rank, n_points=npts, kind=kind,
                                    chain_length=1 << 30, track_steps=nsteps))
    pinned = {}
    pad_rows = None

    def one_step(e2e):
        rows = []
        for e in ens:
            for _ in range(nsteps):               # (the track is a ring of nsteps rows: a block fills it once)
                e.step()
            rows.append(e.track[:nsteps])
        if grid:
            # the only collective of the inversion loop (point.py:112-123 collects the sub-chain files; model3D.py:50-57
            # reads one file per node): the rows of this block from every rank, and the best misfit of the block
            r = rows[0].permute(1, 0, 2).contiguous()                       # [chains_local, steps, 3 + P]
            if world > 1:
                nonlocal pad_rows
                if pad_rows is None:
                    pad_rows = torch.zeros((npts_max * cpp,) + tuple(r.shape[1:]), dtype=r.dtype, device=dev)
                pad_rows[: r.shape[0]] = r                                   # equal block per rank
                full = gather_chain_rows(pad_rows)
                best = r[:, :, 0].amin()
                dist.all_reduce(best, op=dist.ReduceOp.MIN)
                rows = [full]
            else:
                rows = [r]
        if e2e:
            for i, r in enumerate(rows):
                if i not in pinned or pinned[i].shape != r.shape:
                    pinned[i] = torch.empty(r.shape, dtype=r.dtype, pin_memory=True)
                pinned[i].copy_(r, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
        return rows

    for _ in range(max(args.warmup, 3)):
        one_step(False)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_step(False)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    tm = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms_max = float(tm.item())
    total_chains = (args.points if grid else 1) * cpp
    evals = total_chains * nsteps * K * len(kinds) * args.steps
    value = evals / (ms_max * 1e-3)
    ctr = (C.c_ulonglong * 4)()
    solver.lib.surfdisp_read_counters(ens[0].ws.data_ptr(), ctr, torch.cuda.current_stream(dev).cuda_stream)
    coll_ms = None
    if grid and world > 1:      # the collective alone (its share of the step)
        barrier(); t0 = time.perf_counter()
        for _ in range(10):
            gather_chain_rows(pad_rows)
        barrier(); coll_ms = (time.perf_counter() - t0) / 10 * 1e3
    e2e = None
    if not args.no_e2e:         # end to end: the host receives the chain rows of every block (pinned D2H in the timed region)
        one_step(True)
        barrier(); t0 = time.perf_counter()
        for _ in range(args.steps):
            rows = one_step(True)
        barrier(); dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": evals / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": 0,
               "d2h_bytes_per_step": int(sum(r.numel() for r in rows) * 4),
               "note": "the chains live on the device; per timed step the host receives the [chains][steps][3+P] track rows"}
    if rank == 0:
        peaks = solver.measure_peaks()
        ev_last = ens[0].M * K
        cpu = None
        if not args.no_cpu:
            nth = os.cpu_count() or 1
            v, n, dt, cc = cpu_sample(per, args.cpu_seconds, nth, stacks=mc_cpu_stacks(args, max(64, 16 * nth)))
            cpu = {"value": v, "unit": UNIT, "cores": nth, "kind": "port",
                   "sample": "%d admissible %d-layer models x %d periods in %.1f s (forward solve only: the reference's Python "
                             "model assembly and prior checks per sample are not in it)" % (n, nl_case, K, dt)}
        ach = ctr[0] / ev_last * F_R * value / world * 1e-12
        line = {"metric": METRIC.replace("c+U", "c"), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "strong" if grid else "weak", "vs_baseline": None,
                "dtype": "f32 (root search), f64 (model assembly, misfit)", "data": "synthetic",
                "config": {"workload": mc_workload_name(args, world), "chains_total": total_chains, "mc_steps_per_step": nsteps,
                           "periods": K, "params": P, "ms_per_mc_step": ms_max / args.steps / nsteps / len(kinds),
                           "chain_steps_per_s": total_chains * nsteps * args.steps * len(kinds) / (ms_max * 1e-3),
                           "accept_rate": float(ens[0].accepted.float().mean()), "cuda_graph": True,
                           "l2_policy": "compute / latency bound; every step rewrites %.1f MB of stacks and constants"
                                        % (ens[0].M * ens[0].lmax * 52 / 1e6),
                           "collective_ms_per_step": coll_ms},
                "clocks": clocks, "e2e": e2e, "gpu_launches": (7 + p1_launches(ens[0].M)) * nsteps * len(kinds) * args.steps,
                "roofline": {"bound": "fp32", "kernel": "phase1_kernel<4> inside the Monte-Carlo step",
                             "layer_steps_per_eval": ctr[0] / ev_last, "sweeps_per_eval": ctr[1] / ev_last,
                             "achieved": ach, "peak": peaks[0], "unit": "TFLOP/s", "frac": ach / peaks[0] if peaks[0] else None,
                             "note": "layer-steps of the last step x 190 FLOP-eq over the WHOLE step time (proposal, assembly, "
                                     "search, misfit): per-kernel shares in profiles/r2_mc_launches.txt", "traffic": None},
                "cpu_baseline": cpu}
        emit(line)


def main():
    args = parse()
    _claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from pysurfinv_b200 import api, synth, stack as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the solver has no CPU path)")
    pinned_to = pin_to_gpu_numa(local) if world > 1 else "single rank: not pinned"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    if args.workload != "sweep":
        run_mc(args, rank, world, local, dev, barrier)
        if world > 1:
            dist.destroy_process_group()
        return

    M, K = args.models, args.periods
    per = synth.log_periods(K)
    solver = api.DispersionSolver(dev)
    tmpl, _setting = S.config2_template()
    lmax = 77
    h_par = torch.from_numpy(S.config2_params(M, seed=synth.DEFAULT_SEED + 1000 * rank)).pin_memory()   # every rank its own models
    d_par = h_par.to(dev)
    d_lay, d_nl = solver.build_stacks(tmpl, d_par, lmax=lmax)     # the stacks resident in HBM (model assembly: SURVEY 8 f-1)
    out = solver.forward(d_lay, d_nl, per, kind=2)   # allocates outputs + workspace
    torch.cuda.synchronize(dev)
    assert int(d_nl.min()) == 77 and int(d_nl.max()) == 77

    # ---- device-resident timing: K steps between CUDA events, max over ranks
    for _ in range(args.warmup):
        solver.forward(d_lay, d_nl, per, kind=2, out=out)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        solver.forward(d_lay, d_nl, per, kind=2, out=out)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * M * K * args.steps / (ms_max * 1e-3)
    nfound_ok = int((out["nfound"] == K).sum().item())
    steps_ctr, sweeps_ctr, subu_ctr, models_ctr = solver.counters()

    # ---- per-kernel durations (CUDA events inside the library, same workload) for the roofline
    kms = []
    for _ in range(max(1, args.steps)):
        m3 = [0, 0, 0]
        solver.forward(d_lay, d_nl, per, kind=2, out=out, kernel_ms=m3)
        kms.append(m3)
    kms = np.mean(np.array(kms), axis=0)
    peaks = solver.measure_peaks()

    # ---- the same sweep for Love waves (kind = 1; the headline stays the Rayleigh configuration of BASELINE.json)
    out_l = solver.forward(d_lay, d_nl, per, kind=1)
    barrier()
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record()
    for _ in range(args.steps):
        solver.forward(d_lay, d_nl, per, kind=1, out=out_l)
    l1.record()
    barrier()
    tl = torch.tensor([l0.elapsed_time(l1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tl, op=dist.ReduceOp.MAX)
    love_found = int(out_l["nfound"].sum().item())          # Love curves end where the root reaches the half-space velocity
    love_full = int((out_l["nfound"] == K).sum().item())
    l_steps, l_sweeps, l_sub, _ = solver.counters()
    lk = [0, 0, 0]
    solver.forward(d_lay, d_nl, per, kind=1, out=out_l, kernel_ms=lk)
    l_ach = l_steps * F_L / (lk[1] * 1e-3) * 1e-12
    love = {"value": world * love_found * args.steps / (float(tl.item()) * 1e-3), "unit": UNIT,
            "note": "found evaluations only: %.1f %% of the Love curves end before the longest period (reference behaviour, "
                    "identical root counts on both sides)" % (100.0 * (1.0 - love_full / M)),
            "ms_per_step": float(tl.item()) / args.steps, "evals_found_frac": love_found / (M * K),
            "roofline": {"bound": "fp32", "kernel": "phase1_kernel<4> (Love sweep)", "achieved": l_ach, "peak": peaks[0],
                         "unit": "TFLOP/s", "frac": l_ach / peaks[0] if peaks[0] else None,
                         "kernel_ms": {"prep": lk[0], "phase1": lk[1], "phase2": lk[2]},
                         "layer_steps_per_eval": l_steps / max(1, love_found),
                         "note": "a Love layer-step is 40 FLOP-equivalents (SURVEY 8d): the sweep is bound by the layer-record "
                                 "loads, the reciprocal and the two series per step, not by FMA throughput"}}
    del out_l

    # ---- end to end
    e2e = e2e_layers = None
    if not args.no_e2e:
        def timed(fn):
            for _ in range(min(args.warmup, 2)):
                fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                res = fn()
            barrier()
            tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return world * M * K * args.steps / float(tt.item()), res
        # (a) the user's call: parameter vectors on the host -> c, U, nfound, flags on the host
        v, res = timed(lambda: solver.forward_params_pinned(tmpl, h_par, per, kind=2, lmax=lmax))
        assert int((torch.from_numpy(res["nfound"]) == K).sum()) == nfound_ok
        d2h = int(2 * M * K * 4 + 2 * M * 4)
        e2e = {"value": v, "unit": UNIT, "h2d_bytes_per_step": int(h_par.numel() * 4), "d2h_bytes_per_step": d2h,
               "call": "DispersionSolver.forward_params_pinned -> surfdisp_host_params_pipelined: the batch form of "
                       "Model1D.forward (models.py:93-121): parameter vectors in, stacks assembled on the device"}
        # (b) the fast_surf-level call: layer arrays on the host in
        h_lay = d_lay.cpu().pin_memory(); h_nl = d_nl.cpu().pin_memory()
        v2, res2 = timed(lambda: solver.forward_pinned(h_lay, h_nl, per, kind=2))
        assert int((torch.from_numpy(res2["nfound"]) == K).sum()) == nfound_ok
        e2e_layers = {"value": v2, "unit": UNIT, "h2d_bytes_per_step": int(h_lay.numel() * 4 + h_nl.numel() * 4),
                      "d2h_bytes_per_step": d2h, "call": "DispersionSolver.forward_pinned -> surfdisp_host_batch_pipelined "
                                                          "(the batch form of fast_surf.fast_surf, models.py:27)"}
        del h_lay

    # ---- single-model latency of the drop-in entry (what an unmodified models.py:27 loop sees)
    single = None
    if rank == 0:
        from pysurfinv_b200 import fast_surf as FS
        lay1 = d_lay[:, 0].cpu().numpy().astype(np.float64)
        per200 = np.zeros(200); per200[:18] = MC_PERIODS[1:]
        FS.fast_surf(77, 2, lay1[0], lay1[1], lay1[2], lay1[3], lay1[4], per200, 18)
        t0 = time.perf_counter()
        for _ in range(50):
            FS.fast_surf(77, 2, lay1[0], lay1[1], lay1[2], lay1[3], lay1[4], per200, 18)
        single = {"us_per_call": (time.perf_counter() - t0) / 50 * 1e6, "layers": 77, "periods": 18,
                  "call": "fast_surf.fast_surf -> fast_surf_ (per-thread device context kept between calls)"}
        if not args.no_cpu:     # the same single-model call on one host core (oracle port; cpu_baseline leg)
            from oracle import oracle as O
            f32 = lambda x: x.astype(np.float32).astype(np.float64)
            O.forward(2, f32(lay1[0]), f32(lay1[1]), f32(lay1[2]), f32(lay1[3]), f32(lay1[4]), MC_PERIODS[1:], opts=O.make_opts(precision=0))
            t0 = time.perf_counter()
            for _ in range(10):
                O.forward(2, f32(lay1[0]), f32(lay1[1]), f32(lay1[2]), f32(lay1[3]), f32(lay1[4]), MC_PERIODS[1:], opts=O.make_opts(precision=0))
            single["cpu_port_us_per_call"] = (time.perf_counter() - t0) / 10 * 1e6

    if rank == 0:
        flop_p1 = steps_ctr * F_R
        lpad = 80
        ach = flop_p1 / (kms[1] * 1e-3) * 1e-12
        ex2 = subu_ctr * F_U_EXEC / (kms[2] * 1e-3) * 1e-12
        roof = {"bound": "fp32", "kernel": "phase1_kernel<4> (root search: 3 launches -- first period, fast path, handed-over models)",
                "achieved": ach, "peak": peaks[0], "unit": "TFLOP/s", "frac": ach / peaks[0] if peaks[0] else None,
                "peak_source": "surfdisp_measure_peaks(): register-resident FFMA chain, this run "
                               "(MEASURED_PEAKS.json has no FP32 figure; the path is FP-pipe bound, not HBM or tensor)",
                "work_unit": "secular-function layer-step of one trial velocity = %.0f FLOP-equivalents (SURVEY 8d); "
                             "counted on the device" % F_R,
                "traffic": DRAM_P1_PER_MODEL * M,
                "traffic_note": "dram__bytes_read + write of the two root-search launches (ncu, profiles/r2_ncu_phase1.txt), scaled "
                                "to this batch; algorithmic %.2e B (constants read by both launches + c, ratio written): HBM at "
                                "<1 %% of its peak, the path is FP-pipe bound" % (algo_p1_bytes(lpad, K) * M),
                "kernel_ms": {"prep": float(kms[0]), "phase1": float(kms[1]), "phase2": float(kms[2])},
                "layer_steps_per_eval": steps_ctr / (M * K), "sweeps_per_eval": sweeps_ctr / (M * K),
                "u_sublayers_per_eval": subu_ctr / (M * K),
                "phase2": {"bound": "fp32", "executed": ex2, "peak": peaks[0], "unit": "TFLOP/s",
                           "frac_executed": ex2 / peaks[0] if peaks[0] else None,
                           "reference_equivalent_tflops": subu_ctr * F_U / (kms[2] * 1e-3) * 1e-12, "traffic": DRAM_P2_PER_MODEL * M,
                           "note": "executed = float32 FLOP of this kernel's formulation (%.0f per sub-layer, ncu instruction counts; "
                                   "float32 ODE state re-orthogonalised after every sub-layer, the float64 state of the reference is "
                                   "opts.group_f64 = 1); reference_equivalent = the same sub-layers at the reference's %.0f FLOP each "
                                   "-- a work-saving factor, not a utilisation" % (F_U_EXEC, F_U)},
                "hbm": hbm_roofline(algo_p1_bytes(lpad, K) * M, DRAM_P1_PER_MODEL * M, float(kms[1])),
                "mufu_peak_Tops": peaks[2]}
        cpu = None
        if not args.no_cpu:
            nth = os.cpu_count() or 1
            v, n, dt, cc = cpu_sample(per, args.cpu_seconds, nth)
            cpu = {"value": v, "unit": UNIT, "cores": nth, "kind": "port",
                   "sample": "%d models x %d periods of the same generator in %.1f s; float32-faithful C++ oracle "
                             "(no Fortran compiler in the image or on the GPU box)" % (n, K, dt),
                   "note": "-O2, no FMA contraction, per-period std::vector set-up: a conservative stand-in for gfortran -O3",
                   "layer_steps_per_eval": cc["steps_R"] / (n * K), "sweeps_per_eval": cc["sweeps_R"] / (n * K)}
            roof["ref_equiv_tflops"] = value / world * (cc["steps_R"] / (n * K) * F_R + cc["sub_U"] / (n * K) * F_U) * 1e-12
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 (root search and group-velocity ODE state; the reference's float64 state: opts.group_f64)",
                "data": "synthetic",
                "config": {"workload": WORKLOAD_SWEEP % (M, K),
                           "models_per_gpu": M, "periods": K, "layers": lmax, "params_per_model": int(h_par.shape[1]),
                           "l2_policy": "inputs (%.2f GB layers + %.2f GB workspace per step) exceed the 126 MB L2"
                                        % (d_lay.numel() * 4 / 1e9, solver._ws.numel() / 1e9),
                           "roots_found_frac": nfound_ok / M, "rank_cpu_affinity": pinned_to},
                "clocks": clocks, "e2e": e2e, "e2e_layers": e2e_layers, "gpu_launches": (5 + p1_launches(M)) * args.steps, "roofline": roof,
                "cpu_baseline": cpu, "love": love, "single_model_call": single}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
