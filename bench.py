#!/usr/bin/env python
"""bench.py -- dispersion evaluations per second (BASELINE.json metric) on config 2:
batched forward sweep, 1 Mi random sediment+crust+mantle models (n = 77 layers) x 40 periods (8-80 s),
Rayleigh phase + group velocity, per GPU (weak scaling: every rank solves its own 1 Mi models).

    python bench.py --gpus N --steps K --warmup W            (torchrun launches N ranks for N > 1)
    python bench.py --impl reference ...                      (CPU oracle port on the host cores)

One JSON line on stdout (rank 0).  1 evaluation = one (model, period) producing c and U.
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
# (10 FLOP each), REIGEN sub-layer 1800 FLOP (fp64), flattening 12 FLOP + 3 transcendentals per layer.
F_R, F_U, F_FLAT = 190.0, 1800.0, 42.0
# FP64 FLOP the group-velocity kernel actually executes per sub-layer (per-layer step matrix + quadratic-form
# accumulation instead of the reference's stage-by-stage RK4): 2 x DFMA + DMUL + DADD of the ncu capture
# profiles/r1_ncu_prep_phase2_final.txt divided by the sub-layers counted on the device.
F_U_EXEC = 520.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--models", type=int, default=1 << 20, help="models per GPU per step")
    ap.add_argument("--periods", type=int, default=40)
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


def cpu_sample(per, seconds, nthreads, chunk=None, seed=99):
    """Times the CPU oracle (float32-faithful port of fast_surf) on a bounded sample of the workload."""
    from oracle import oracle as O
    from pysurfinv_b200 import synth
    chunk = chunk or max(64, 16 * nthreads)
    lay, nl = synth.crustal_models(chunk, seed=seed)
    O.forward_batch(2, lay[:, :nthreads], nl[:nthreads], per, nthreads=nthreads)  # warm-up (page in, threads)
    done, t0 = 0, time.perf_counter()
    tot = {}
    while True:
        cnt = O.OracleCounters()
        O.forward_batch(2, lay, nl, per, opts=O.make_opts(precision=0), nthreads=nthreads, counters=cnt)
        for k, v in cnt.as_dict().items():
            tot[k] = tot.get(k, 0) + v
        done += chunk
        dt = time.perf_counter() - t0
        if dt >= seconds:
            break
    return done * len(per) / dt, done, dt, tot


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port: no Fortran compiler in the image), all host
    threads, bounded sample per step."""
    if rank != 0:
        return
    from pysurfinv_b200 import synth
    per = synth.log_periods(args.periods)
    nthreads = os.cpu_count() or 1
    sec = max(2.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_sample(per, 0.5, nthreads)
    vals, tot_models, tot_t = [], 0, 0.0
    for _ in range(args.steps):
        v, n, dt, _c = cpu_sample(per, sec, nthreads)
        vals.append(v); tot_models += n; tot_t += dt
    value = tot_models * len(per) / tot_t
    sample = "%d models x %d periods per step (%.1f s), config-2 generator" % (tot_models // max(1, args.steps), len(per), sec)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config2: 1Mi x 77-layer models x 40 periods, Rayleigh c+U (bounded CPU sample)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
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
    from pysurfinv_b200 import api, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the solver has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    M, K = args.models, args.periods
    per = synth.log_periods(K)
    lay, nl = synth.crustal_models(M, seed=synth.DEFAULT_SEED + 1000 * rank)   # every rank its own models
    solver = api.DispersionSolver(dev)
    h_lay = torch.from_numpy(lay).pin_memory()
    h_nl = torch.from_numpy(nl).pin_memory()
    d_lay = h_lay.to(dev); d_nl = h_nl.to(dev)
    out = solver.forward(d_lay, d_nl, per, kind=2)   # allocates outputs + workspace
    torch.cuda.synchronize(dev)

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
    love = {"value": world * M * K * args.steps / (float(tl.item()) * 1e-3), "unit": UNIT,
            "ms_per_step": float(tl.item()) / args.steps, "roots_found_frac": int((out_l["nfound"] == K).sum().item()) / M}
    del out_l

    # ---- per-kernel durations (CUDA events inside the library, same workload) for the roofline
    kms = []
    for _ in range(max(1, args.steps)):
        m3 = [0, 0, 0]
        solver.forward(d_lay, d_nl, per, kind=2, out=out, kernel_ms=m3)
        kms.append(m3)
    kms = np.mean(np.array(kms), axis=0)
    peaks = solver.measure_peaks()

    # ---- end to end: pinned host inputs -> H2D -> solve -> D2H of c, U, nfound, flags, every step
    e2e = None
    if not args.no_e2e:
        for _ in range(min(args.warmup, 2)):
            solver.forward_pinned(h_lay, h_nl, per, kind=2)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = solver.forward_pinned(h_lay, h_nl, per, kind=2)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * M * K * args.steps / float(tt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(h_lay.numel() * 4 + h_nl.numel() * 4),
               "d2h_bytes_per_step": int(2 * M * K * 4 + 2 * M * 4)}
        assert int((torch.from_numpy(res["nfound"]) == K).sum()) == nfound_ok

    if rank == 0:
        flop_p1 = steps_ctr * F_R
        roof = {"bound": "fp32", "kernel": "phase1_kernel<4> (root search: 2 launches, first period / later periods)",
                "achieved": flop_p1 / (kms[1] * 1e-3) * 1e-12, "peak": peaks[0], "unit": "TFLOP/s",
                "frac": flop_p1 / (kms[1] * 1e-3) * 1e-12 / peaks[0] if peaks[0] else None,
                "peak_source": "surfdisp_measure_peaks(): register-resident FFMA chain, this run "
                               "(MEASURED_PEAKS.json has no FP32 figure; the path is FP-pipe bound, not HBM or tensor)",
                "work_unit": "secular-function layer-step of one trial velocity = %.0f FLOP-equivalents (SURVEY 8d); "
                             "counted on the device" % F_R,
                "traffic": None,
                "kernel_ms": {"prep": float(kms[0]), "phase1": float(kms[1]), "phase2": float(kms[2])},
                "layer_steps_per_eval": steps_ctr / (M * K), "sweeps_per_eval": sweeps_ctr / (M * K),
                "u_sublayers_per_eval": subu_ctr / (M * K),
                "phase2": {"bound": "fp64", "achieved": subu_ctr * F_U / (kms[2] * 1e-3) * 1e-12, "peak": peaks[1],
                           "unit": "TFLOP/s", "executed": subu_ctr * F_U_EXEC / (kms[2] * 1e-3) * 1e-12,
                           "frac_executed": (subu_ctr * F_U_EXEC / (kms[2] * 1e-3) * 1e-12 / peaks[1]) if peaks[1] else None,
                           "note": "achieved = reference-equivalent work (SURVEY 8d, %.0f FLOP per sub-layer); executed = "
                                   "FP64 FLOP of this kernel's formulation (%.0f per sub-layer)" % (F_U, F_U_EXEC)},
                "mufu_peak_Tops": peaks[2]}
        cpu = None
        if not args.no_cpu:
            nth = os.cpu_count() or 1
            v, n, dt, cc = cpu_sample(per, args.cpu_seconds, nth)
            cpu = {"value": v, "unit": UNIT, "cores": nth, "kind": "port",
                   "sample": "%d models x %d periods of the same generator in %.1f s; float32-faithful C++ oracle "
                             "(no Fortran compiler in the image)" % (n, K, dt),
                   "layer_steps_per_eval": cc["steps_R"] / (n * K), "sweeps_per_eval": cc["sweeps_R"] / (n * K)}
            roof["ref_equiv_tflops"] = value / world * (cc["steps_R"] / (n * K) * F_R + cc["sub_U"] / (n * K) * F_U) * 1e-12
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 (root search) + f64 (group-velocity ODE)",
                "data": "synthetic",
                "config": {"workload": "config2: batched forward sweep, %d random 77-layer sediment+crust+mantle models "
                                       "x %d periods 8-80 s per GPU, Rayleigh phase+group" % (M, K),
                           "models_per_gpu": M, "periods": K, "layers": int(lay.shape[2]),
                           "l2_policy": "inputs (%.2f GB layers + %.2f GB workspace per step) exceed the 126 MB L2"
                                        % (lay.nbytes / 1e9, solver._ws.numel() / 1e9),
                           "roots_found_frac": nfound_ok / M},
                "clocks": clocks, "e2e": e2e, "gpu_launches": 7 * args.steps, "roofline": roof, "cpu_baseline": cpu,
                "love": love}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
