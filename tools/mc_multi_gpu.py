"""GPU box, torchrun: every rank runs its own block of Monte-Carlo chains (seeds differ per rank) and the per-chain
rows of the last step are gathered with NCCL -- the only collective of the inversion loop."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from pysurfinv_b200 import api, mc, stack as S
from pysurfinv_b200.distributed import gather_chain_rows, best_misfit
from tests.test_gpu_mc import SETTING

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
s = api.DispersionSolver("cuda:%d" % local)
t = S.StackTemplate(SETTING, prior_mask=S.P_ALL)
per = np.array([8, 10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 50, 60, 70, 80], np.float32)
start = torch.from_numpy(t.start_values()[None, :]).cuda().contiguous()
truth = s.mc_propose(t, start, seed=99, step_index=0, reset_mask=torch.ones(1, dtype=torch.uint8, device="cuda"))
lay, nl = s.build_stacks(t, truth)
obs = s.forward(lay, nl, per, kind=2)["c"][0].cpu().numpy()
ens = mc.ChainEnsemble(s, t, per, obs, np.full(len(per), 0.01, np.float32), n_chains=1024, seed=1000 + rank)
ens.run(50)
rows = ens.track[-1]
allrows = gather_chain_rows(rows)
best = best_misfit(rows[:, 0].min().clone())
if rank == 0:
    print("gathered rows", tuple(allrows.shape), "best misfit over %d ranks %.4f" % (world, float(best)))
    assert allrows.shape[0] == world * 1024
dist.destroy_process_group()
