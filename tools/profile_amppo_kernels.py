import os, sys
sys.path.insert(0, "/root/repo")
import torch
from torch.profiler import ProfilerActivity, profile
from f16_jsb_b200 import F16BatchedEnv
from f16_jsb_b200.amppo import AMPPO, AMPPOConfig
env = F16BatchedEnv(1024, mode="fp32")
algo = AMPPO(env, AMPPOConfig(n_steps=128, batch_size=131072, n_epochs=1))
algo.collect_rollouts(); algo.train(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    algo.train(); torch.cuda.synchronize()
rows = [(e.key, e.self_device_time_total, e.count) for e in prof.key_averages() if e.self_device_time_total > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print("total CUDA us", tot)
for k, t, c in rows[:45]:
    print("%7.1f us %5.1f%% x%-3d %s" % (t, 100 * t / tot, c, k[:110]))
