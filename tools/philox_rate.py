"""Step rate with HBM-resident action batches against in-kernel Philox actions (1M envs, FP32): measured equal."""
import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from f16_jsb_b200 import F16BatchedEnv
n=1048576
env=F16BatchedEnv(n, mode="fp32", seed=0); env.reset()
acts=[torch.rand((n,4),device="cuda")*torch.tensor([2,2,2,1],device="cuda")-torch.tensor([1,1,1,0],device="cuda") for _ in range(8)]
for name,get in (("ring8", lambda k: acts[k%8]), ("philox", lambda k: None), ("ring8", lambda k: acts[k%8]), ("philox", lambda k: None)):
    for k in range(300): env.step(get(k))
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(1000): env.step(get(k))
    e1.record(); torch.cuda.synchronize()
    print(name, "ms/step %.4f"%(e0.elapsed_time(e1)/1000))
