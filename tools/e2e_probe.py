"""Per-phase host timing of F16VecEnv.step (host-resident windows) at 1M envs: where an end-to-end step goes."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from f16_jsb_b200 import F16VecEnv  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
steps = 30
for rings, both in ((2, False), (2, True), (1, False)):
    venv = F16VecEnv(n, mode="fp32", host_obs="window", host_rings=rings, host_dma_both=both)
    venv.reset()
    rng = np.random.default_rng(1)
    bufs = [venv.action_buffer(), venv.action_buffer()]
    for b in bufs:
        b[...] = rng.uniform([-1, -1, -1, 0], [1, 1, 1, 1], size=(n, 4)).astype(np.float32)
    for k in range(int(os.environ.get('WARM', '600'))):
        venv.step(bufs[k % 2])
    venv._win.timing(reset=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fin = 0
    for k in range(steps):
        fin += int(np.count_nonzero(venv.step(bufs[k % 2])[2]))
    dt = (time.perf_counter() - t0) / steps
    ph = venv._win.timing()
    print("rings %d dma_both %d: %.3f ms/step (%.3g env-steps/s), in C %.3f ms:" % (rings, both, dt * 1e3, n / dt, sum(v for k, v in ph.items() if k != "carry_over_duration")),
          "finished/step %.0f" % (fin / steps), {k: round(v, 3) for k, v in ph.items() if k != "carry_over_duration" or v})
    venv.close()
