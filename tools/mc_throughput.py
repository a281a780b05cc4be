"""GPU box: Monte-Carlo ensemble throughput (chain steps per second, every stage on the device, CUDA-graph replay).
usage: python tools/mc_throughput.py [chains] [steps] [graph 0/1]"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from pysurfinv_b200 import api, mc, stack as S
from tests.test_gpu_mc import SETTING

M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
graph = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
s = api.DispersionSolver("cuda:0")
t = S.StackTemplate(SETTING, prior_mask=S.P_ALL)
per = np.array([8, 10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 50, 60, 70, 80], np.float32)
start = torch.from_numpy(t.start_values()[None, :]).cuda().contiguous()
truth = s.mc_propose(t, start, seed=99, step_index=0, reset_mask=torch.ones(1, dtype=torch.uint8, device="cuda"))
lay, nl = s.build_stacks(t, truth)
obs = s.forward(lay, nl, per, kind=2)["c"][0].cpu().numpy()
ens = mc.ChainEnsemble(s, t, per, obs, np.full(len(per), 0.01, np.float32), n_chains=M, seed=1, chain_length=1 << 30, use_graph=graph)
for _ in range(4):
    ens.step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(steps):
    ens.step()
torch.cuda.synchronize(); dt = time.perf_counter() - t0
import ctypes as C
arr = (C.c_ulonglong * 4)()
s.lib.surfdisp_read_counters(ens.ws.data_ptr(), arr, torch.cuda.current_stream().cuda_stream)
ev = M * len(per)
print("last step: layer-steps/eval %.1f, sweeps/eval %.2f, full curves %.4f, mean layers %.1f" % (arr[0] / ev, arr[1] / ev, float((ens.nfound == len(per)).float().mean()), float(ens.stacks[1].float().mean())))
print("%d chains x %d steps x %d periods (graph %d): %.3f ms/step, %.3g chain-steps/s, %.3g evals/s, accept rate %.2f"
      % (M, steps, len(per), graph, dt / steps * 1e3, M * steps / dt, M * steps * len(per) / dt, float(ens.accepted.float().mean())))
