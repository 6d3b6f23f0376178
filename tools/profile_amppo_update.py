"""Where one AM-PPO minibatch update goes (torch profiler, top CUDA kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from f16_jsb_b200 import F16BatchedEnv
from f16_jsb_b200.amppo import AMPPO, AMPPOConfig

env = F16BatchedEnv(1024, mode="fp32")
algo = AMPPO(env, AMPPOConfig(n_steps=128, batch_size=131072, n_epochs=1))
algo.collect_rollouts(); algo.train(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    algo.train(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    algo.collect_rollouts(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=70))
