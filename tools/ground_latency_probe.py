#!/usr/bin/env python3
"""How long does a warp take when one of its envs touches the ground? Times f16_step (CUDA events) on small and
large batches in which a chosen fraction of the envs sits in a contact state (steep dive, radome at the surface)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from f16_jsb_b200 import F16BatchedEnv  # noqa: E402
from test_ground_contact import synthetic_state  # noqa: E402
import re  # noqa: E402

txt = open(os.path.join(ROOT, "include", "f16_state_fields.h")).read()
body = txt.split("enum f16_state_field")[1].split("F16_NUM_STATE_FIELDS")[0]
FIELDS = []
for m in re.finditer(r"F16S_([A-Z0-9_]+)", body):
    if m.group(1) not in FIELDS:
        FIELDS.append(m.group(1))


def timed(env, act, reps):
    st = env.pack_states().clone()
    ts = []
    for _ in range(reps):
        env.unpack_states(st)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        env.step(act, auto_reset=False)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    for n, every in ((32, 0), (32, 32), (32, 1), (1 << 20, 0), (1 << 20, 8192), (1 << 20, 1024), (1 << 20, 128)):
        env = F16BatchedEnv(n, mode=mode)
        env.reset()
        base = env.snapshot()[0]
        fly = np.tile(base, (n, 1))
        dive = synthetic_state(base, FIELDS, 20.0, 0.2, -1.3, 5.0, (700.0, 0.0, 30.0), (0.0, -0.2, 0.0))
        if every:
            fly[::every] = dive
        env.unpack_states(torch.from_numpy(fly).cuda())
        act = torch.tensor([[0.2, -0.1, 0.1, 0.7]], device="cuda").repeat(n, 1)
        print("%s n=%d contact envs=%d: %.1f us per step" % (mode, n, 0 if not every else len(range(0, n, every)), timed(env, act, 7)), flush=True)


if __name__ == "__main__":
    main()
