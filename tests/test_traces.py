"""Every tests/golden/*.f16trace.npz (format: f16_jsb_b200/trace.py) is replayed through the kernel source on
the CPU (tests/hostsim), through the oracle's C env layer, and - on the B200 - through the CUDA library.
Traces whose producer is "jsbsim <version>" pin parity against the real JSBSim; the ones committed today
were recorded where JSBSim is not installable (producer "cpu-restatement": the reference's own Python env
layer on top of the oracle FDM), so they pin the format and the env layer, and exercise a steep impact with
ground contact (oracle_dive21). Dropping a JSBSim-recorded file into tests/golden/ is all it takes to pin
the FDM: `python tools/record_trace.py --backend jsbsim --out tests/golden/jsbsim_<name>.f16trace.npz`."""
import glob
import os

import numpy as np
import pytest

from f16_jsb_b200.trace import TRACE_PROPERTIES, load_trace, replay_trace, save_trace

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TRACES = sorted(glob.glob(os.path.join(GOLDEN, "*.f16trace.npz")))


def test_there_are_traces_and_they_say_who_made_them():
    assert TRACES
    for p in TRACES:
        t = load_trace(p)
        assert t["header"]["producer"] and t["header"]["properties"][:12] == TRACE_PROPERTIES[:12]
        # the observation is the float32 cast chain of the first twelve properties (jsbsim_gym.py:172-197)
        k = len(t["frames"]) // 2
        assert np.float32(t["props"][k, 3]) == t["frames"][k, 3] and np.float32(t["props"][k, 2]) == t["frames"][k, 2]


def test_format_round_trip(tmp_path):
    t = load_trace(TRACES[0])
    p = str(tmp_path / "copy.f16trace.npz")
    save_trace(p, producer=t["header"]["producer"], seed=t["header"]["seed"], goal=t["goal"], reset_obs=t["reset_obs"],
               reset_props=t["reset_props"], actions=t["actions"], frames=t["frames"], rewards=t["rewards"],
               terminated=t["terminated"], truncated=t["truncated"], props=t["props"], properties=t["header"]["properties"])
    u = load_trace(p)
    for k in ("goal", "reset_obs", "reset_props", "actions", "frames", "rewards", "terminated", "truncated", "props"):
        assert np.array_equal(t[k], u[k])
    assert u["header"]["version"] == 1 and u["header"]["dt"] == 1.0 / 120.0


@pytest.mark.parametrize("path", TRACES, ids=[os.path.basename(p) for p in TRACES])
def test_replay_through_the_oracle_env_layer(path, oracle):
    t = load_trace(path)
    env = oracle.OracleEnv()

    def step(a):
        obs, r, term, trunc = env.step(a)
        return obs[-1], r, term or trunc, trunc

    rep = replay_trace(t, step, env.reset, mode="fp64")
    if t["header"]["producer"] == "cpu-restatement":
        assert rep["max_err_early"] == 0.0 and rep["max_err_late"] == 0.0      # same FDM: bit for bit


@pytest.mark.parametrize("path", TRACES, ids=[os.path.basename(p) for p in TRACES])
def test_replay_through_the_kernel_source_on_cpu(path, hostsim):
    t = load_trace(path)
    env = hostsim.env(0)

    def step(a):
        obs, r, fl, _ = env.step(a)
        return obs[-1], r, bool(fl & 8), bool(fl & 16)

    replay_trace(t, step, env.reset, mode="fp64")
    if "dive" in path:
        # the last env-step of this trace touches the ground (radome first): the recorded gear force is not zero
        i = t["header"]["properties"].index("forces/fbz-gear-lbs")
        assert abs(t["props"][-1, i]) > 1e3


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fp64", "fp32"])
@pytest.mark.parametrize("path", TRACES, ids=[os.path.basename(p) for p in TRACES])
def test_replay_through_the_cuda_library(path, mode):
    import torch
    from f16_jsb_b200 import F16BatchedEnv
    t = load_trace(path)
    env = F16BatchedEnv(1, mode=mode, ground_reactions=True)
    act = torch.zeros((1, 4), dtype=torch.float32, device="cuda")

    def reset(goal):
        return env.reset(goals=torch.from_numpy(np.asarray(goal, np.float32).reshape(1, 3)).cuda())[0].cpu().numpy()

    def step(a):
        act.copy_(torch.from_numpy(np.asarray(a, np.float32).reshape(1, 4)))
        o, r, d, tr = env.step(act, auto_reset=False)
        return o[0, -1].cpu().numpy(), float(r[0].item()), bool(d[0].item()), bool(tr[0].item())

    rep = replay_trace(t, step, reset, mode=mode)
    assert rep["ended_at"] is not None
