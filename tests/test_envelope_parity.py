"""Teacher-forced parity over the flight envelope: random synthetic states - not only the ones the golden
episodes happen to visit - are lifted into the oracle (f16o_fdm_unpack_state) and into the kernel source
(tests/hostsim; the CUDA library on B200 in test_gpu_parity.py::test_envelope_step_parity), one env-step is
taken on both sides and every state field is compared. Covers what the reference's random-action episodes
reach only rarely: the isothermal layer above 36 089 ft, supersonic calibrated airspeed, the alpha / beta /
Mach table edges and their clamps, the low-speed flap and PID-trigger branches, afterburner on and off,
saturated actuators, inverted and vertical attitudes.

Tolerances: FP64 <= 1e-9 relative per field (floors of conftest; measured 1.1e-12 over 800 states, Mach 0.14 to
2.04, 132 of them above 36 089 ft, afterburner lit in 215), FP32 <= 1e-3 (measured 7e-4, on a body rate)."""
import numpy as np
import pytest

from conftest import state_floors
from test_ground_contact import flying_oracle_env, rel_err, synthetic_state

NF = 53


def envelope_states(base, fields, n, seed, region="wide"):
    """n random states. region "flown": what random-action episodes visit; "wide": the whole table domain and beyond."""
    rng = np.random.default_rng(seed)
    ix = {f: i for i, f in enumerate(fields)}
    out, acts = [], []
    for _ in range(n):
        if region == "flown":
            h = rng.uniform(200.0, 12000.0)
            V = rng.uniform(400.0, 1300.0)
            alpha = np.radians(rng.uniform(-10.0, 35.0))
            beta = np.radians(rng.uniform(-15.0, 15.0))
            rates = rng.uniform(-1.5, 1.5, 3)
        else:
            h = rng.choice([rng.uniform(50.0, 60000.0), rng.uniform(36000.0, 36200.0), rng.uniform(100.0, 3000.0)])
            V = rng.choice([rng.uniform(150.0, 2000.0), rng.uniform(350.0, 450.0), rng.uniform(950.0, 1150.0)])
            alpha = np.radians(rng.uniform(-30.0, 95.0))
            beta = np.radians(rng.uniform(-40.0, 40.0))
            rates = rng.uniform(-4.0, 4.0, 3)
        uvw = V * np.array([np.cos(alpha) * np.cos(beta), np.sin(beta), np.sin(alpha) * np.cos(beta)])
        phi, theta, psi = rng.uniform(-np.pi, np.pi), rng.uniform(-1.5, 1.5), rng.uniform(0, 2 * np.pi)
        s = synthetic_state(base, fields, h, phi, theta, psi, uvw, rates)
        # integrator histories and last accelerations: anything plausible (they are inputs of the step)
        for nm, scale in (("AI0_X", 60.0), ("AI1_X", 60.0), ("WDOT_X", 3.0), ("ABODY_X", 60.0)):
            s[ix[nm]:ix[nm] + 3] = rng.uniform(-scale, scale, 3)
        s[ix["VI1_X"]:ix["VI1_X"] + 3] += rng.uniform(-2.0, 2.0, 3)
        s[ix["VI2_X"]:ix["VI2_X"] + 3] += rng.uniform(-4.0, 4.0, 3)
        # stale Auxiliary outputs: close to, not equal to, the current flight condition
        s[ix["ALPHA"]] = alpha + rng.uniform(-0.02, 0.02)
        s[ix["MACH"]] = max(0.0, s[ix["MACH"]] + rng.uniform(-0.02, 0.02))
        s[ix["VC_KTS"]] = max(0.0, s[ix["VC_KTS"]] * rng.uniform(0.4, 1.1))     # crosses the 250 kt flap threshold
        s[ix["NPY"]], s[ix["NPZ"]] = rng.uniform(-1.0, 1.0), rng.uniform(-9.0, 4.0)
        # FCS memories and engine
        s[ix["TEF"]] = rng.choice([0.0, rng.uniform(0.0, 1.0)])
        s[ix["AIL"]], s[ix["ELEV"]] = rng.uniform(-1.0, 1.0, 2)
        s[ix["SB_DEG"]] = rng.choice([0.0, rng.uniform(0.0, 60.0)])
        for nm in ("ROLL_INPREV", "PITCH_INPREV", "YAW_INPREV"):
            s[ix[nm]] = rng.uniform(-1.0, 1.0)
        for nm in ("ROLL_I", "PITCH_I", "YAW_I"):
            s[ix[nm]] = rng.choice([0.0, rng.uniform(-0.05, 0.05)])
        s[ix["N2"]] = rng.uniform(62.0, 100.0)
        s[ix["AUG"]] = float(rng.random() < 0.3)
        out.append(s)
        a = rng.uniform([-1, -1, -1, 0], [1, 1, 1, 1]).astype(np.float32)
        if rng.random() < 0.2:
            a = np.round(a)                       # saturated sticks, idle or full throttle
        acts.append(a)
    return np.stack(out), np.stack(acts)


def oracle_step_all(oracle, states, acts):
    env, goal = flying_oracle_env(oracle)
    after, frames = [], []
    for s, a in zip(states, acts):
        env.fdm.unpack_state(s)
        obs, r, term, trunc = env.step(a)
        after.append(env.fdm.pack_state())
        frames.append(obs[-1].copy())
    return goal, np.stack(after), np.stack(frames)


@pytest.mark.parametrize("region,mode,tol", [("flown", 0, 1e-9), ("wide", 0, 1e-9), ("flown", 1, 1e-3), ("wide", 1, 1e-3)])
def test_envelope_step_parity_kernel_source(oracle, hostsim, state_fields, region, mode, tol):
    floors = state_floors(state_fields)
    env0, _ = flying_oracle_env(oracle)
    states, acts = envelope_states(env0.fdm.pack_state(), state_fields, 400, seed=11 if region == "flown" else 12, region=region)
    goal, want, frames = oracle_step_all(oracle, states, acts)
    hs = hostsim.env(mode)
    hs.reset(goal)
    worst = (0.0, None, None)
    for k in range(len(states)):
        hs.set_state(states[k], current_step=7)
        obs, r, fl, _ = hs.step(acts[k])
        e = rel_err(hs.get_state(NF), want[k], floors)
        if e.max() > worst[0]:
            worst = (float(e.max()), k, state_fields[int(e.argmax())])
        if mode == 0:
            assert np.allclose(obs[-1][:12], frames[k][:12], rtol=1e-6, atol=1e-6)
    assert worst[0] < tol, worst
