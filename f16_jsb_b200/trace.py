"""`.f16trace.npz` - on-disk record of one episode of the reference env, and its replay on the CUDA path.

Why it exists: the reference's arithmetic lives in the third-party `jsbsim`
package, which is not installable where this repository is built, so parity against real JSBSim is
unpinned. A maintainer who has `jsbsim` runs `python tools/record_trace.py --backend jsbsim` once; the
file it writes is dropped into tests/golden/ and from then on `tests/test_traces.py` replays it through
the CPU restatement and through the CUDA library (B200) with the tolerances below - no code changes.

Format (NumPy .npz, every array little-endian, version 1):

    header       0-d <U array holding a JSON object:
                   format       "f16trace"
                   version      1
                   producer     "jsbsim <version>" | "cpu-restatement" | ...
                   aircraft     "f16"
                   dt           1/120 (s, one FDM frame)        down_sample  4 frames per env-step
                   seed         the reset seed (goal draw, jsbsim_gym.py:312-323) or null
                   properties   list of P JSBSim property names recorded after every env-step
                   notes        free text
    goal         (3,)   float32   env.goal after reset
    reset_obs    (10,15) float32  observation returned by reset
    reset_props  (P,)   float64   the properties right after reset
    actions      (T,4)  float32   what was passed to step()
    frames       (T,15) float32   newest row of the observation after each step (jsbsim_gym.py:172-197)
    rewards      (T,)   float32   PositionReward-shaped reward (jsbsim_gym.py:487-509)
    terminated   (T,)   bool      truncated (T,) bool
    props        (T,P)  float64   the properties after each env-step (= after its fourth FDM frame)

`props` carries everything that must survive between frames and is readable by name from a stock
FGFDMExec (FCS actuator positions, engine spool, body rates and accelerations, airspeeds), so a
mismatch can be localised to a model instead of showing up only in the twelve observed quantities.
"""
import json

import numpy as np

FORMAT_VERSION = 1

# Properties recorded next to the observation: real JSBSim names (the CPU restatement used by the tests serves them too).
TRACE_PROPERTIES = [
    # the twelve STATE_FORMAT entries (jsbsim_gym.py:12-25), as doubles before the float32 cast chain
    "position/lat-gc-rad", "position/long-gc-rad", "position/h-sl-meters", "velocities/mach",
    "aero/alpha-rad", "aero/beta-rad", "velocities/p-rad_sec", "velocities/q-rad_sec",
    "velocities/r-rad_sec", "attitude/phi-rad", "attitude/theta-rad", "attitude/psi-rad",
    # FGPropagate / FGAuxiliary
    "velocities/u-fps", "velocities/v-fps", "velocities/w-fps", "velocities/vt-fps", "velocities/vc-kts",
    "velocities/vg-fps", "aero/qbar-psf", "position/h-agl-ft",
    "accelerations/n-pilot-y-norm", "accelerations/n-pilot-z-norm",
    # FGFCS memories
    "fcs/elevator-pos-norm", "fcs/left-aileron-pos-norm", "fcs/speedbrake-pos-deg", "fcs/tef-control",
    # FGPropulsion
    "propulsion/engine/n2", "propulsion/engine/thrust-lbs",
    # FGAerodynamics / FGGroundReactions totals
    "forces/fbx-aero-lbs", "forces/fby-aero-lbs", "forces/fbz-aero-lbs",
    "moments/l-aero-lbsft", "moments/m-aero-lbsft", "moments/n-aero-lbsft",
    "forces/fbz-gear-lbs",
    # FGMassBalance / FGAtmosphere
    "inertia/weight-lbs", "inertia/cg-x-in", "atmosphere/rho-slugs_ft3",
]

# properties that are also fields of the packed CUDA state (include/f16_state_fields.h): name -> (field, scale)
STATE_FIELD_OF_PROPERTY = {
    "velocities/mach": ("MACH", 1.0), "aero/alpha-rad": ("ALPHA", 1.0), "velocities/vc-kts": ("VC_KTS", 1.0),
    "velocities/vg-fps": ("VG", 1.0), "accelerations/n-pilot-y-norm": ("NPY", 1.0), "accelerations/n-pilot-z-norm": ("NPZ", 1.0),
    "fcs/elevator-pos-norm": ("ELEV", 1.0), "fcs/left-aileron-pos-norm": ("AIL", 1.0), "fcs/speedbrake-pos-deg": ("SB_DEG", 1.0),
    "fcs/tef-control": ("TEF", 1.0), "propulsion/engine/n2": ("N2", 1.0),
}


def save_trace(path, *, producer, seed, goal, reset_obs, reset_props, actions, frames, rewards, terminated, truncated, props,
               properties=None, notes=""):
    properties = list(TRACE_PROPERTIES if properties is None else properties)
    header = dict(format="f16trace", version=FORMAT_VERSION, producer=str(producer), aircraft="f16", dt=1.0 / 120.0, down_sample=4,
                  seed=seed, properties=properties, notes=notes)
    T = len(actions)
    arrays = dict(
        header=np.array(json.dumps(header)),
        goal=np.asarray(goal, np.float32).reshape(3), reset_obs=np.asarray(reset_obs, np.float32).reshape(10, 15),
        reset_props=np.asarray(reset_props, np.float64).reshape(len(properties)),
        actions=np.asarray(actions, np.float32).reshape(T, 4), frames=np.asarray(frames, np.float32).reshape(T, 15),
        rewards=np.asarray(rewards, np.float32).reshape(T), terminated=np.asarray(terminated, bool).reshape(T),
        truncated=np.asarray(truncated, bool).reshape(T), props=np.asarray(props, np.float64).reshape(T, len(properties)))
    np.savez_compressed(path, **arrays)


def load_trace(path):
    z = np.load(path, allow_pickle=False)
    header = json.loads(str(z["header"][()]))
    if header.get("format") != "f16trace" or int(header.get("version", 0)) > FORMAT_VERSION:
        raise ValueError("%s is not an f16trace file this version can read (%r)" % (path, header))
    t = {k: z[k] for k in z.files if k != "header"}
    t["header"] = header
    T, P = len(t["actions"]), len(header["properties"])
    assert t["frames"].shape == (T, 15) and t["props"].shape == (T, P) and t["reset_props"].shape == (P,)
    return t


def record_episode(env, fdm, seed, actions, properties=None):
    """Drive `env` (the reference's wrapped env, Gymnasium API) and read `properties` from `fdm`
    (anything with get_property_value) after reset and after every step. Returns the kwargs of save_trace."""
    properties = list(TRACE_PROPERTIES if properties is None else properties)

    def read():
        return np.array([fdm.get_property_value(p) for p in properties], dtype=np.float64)

    obs, _ = env.reset(seed=seed)
    rec = dict(seed=seed, goal=np.asarray(env.unwrapped.goal, np.float32).copy(), reset_obs=np.asarray(obs, np.float32).copy(),
               reset_props=read(), properties=properties)
    frames, rewards, term, trunc, props = [], [], [], [], []
    for a in actions:
        obs, r, te, tr, _ = env.step(np.asarray(a, np.float32))
        frames.append(np.asarray(obs, np.float32)[-1].copy())
        rewards.append(np.float32(r))
        term.append(bool(te))
        trunc.append(bool(tr))
        props.append(read())
        if te or tr:
            break
    n = len(frames)
    rec.update(actions=np.asarray(actions, np.float32)[:n], frames=np.stack(frames), rewards=np.array(rewards, np.float32),
               terminated=np.array(term), truncated=np.array(trunc), props=np.stack(props))
    return rec


# Stated replay tolerances. Free-running comparison of a sensitive system: frames agree to
# `early` (relative, floor 1e-2) for the first `early_steps` steps, to `late` afterwards; the episode ends
# within one step of the recorded one.
REPLAY_TOLERANCE = {"fp64": dict(early_steps=300, early=1e-5, late=1e-1, reward=2e-5),
                    "fp32": dict(early_steps=100, early=2e-3, late=None, reward=1e-3)}


def replay_trace(trace, step_fn, reset_fn, mode="fp64"):
    """Replay a trace through an implementation given as two callables:
    reset_fn(goal (3,) f32) -> obs (10,15); step_fn(action (4,) f32) -> (frame (15,), reward, done, truncated).
    Returns a report dict; raises AssertionError on a violation of REPLAY_TOLERANCE[mode]."""
    tol = REPLAY_TOLERANCE[mode]
    obs = reset_fn(trace["goal"])
    e0 = float(np.abs(np.asarray(obs, np.float64) - trace["reset_obs"]).max())
    assert e0 <= 1e-6 * max(1.0, float(np.abs(trace["reset_obs"]).max())), "reset observation differs by %.3g" % e0
    n = len(trace["actions"])
    e_early = e_late = 0.0
    ended = None
    for k in range(n):
        frame, reward, done, trunc = step_fn(trace["actions"][k])
        f = trace["frames"][k]
        e = float((np.abs(np.asarray(frame[:12], np.float64) - f[:12]) / np.maximum(np.abs(f[:12]), 1e-2)).max())
        if k < tol["early_steps"]:
            e_early = max(e_early, e)
            assert e <= tol["early"], "frame %d differs by %.3g" % (k, e)
            assert abs(float(reward) - float(trace["rewards"][k])) <= tol["reward"], "reward %d" % k
        else:
            e_late = max(e_late, e)
            if tol["late"] is not None:
                assert e <= tol["late"], "frame %d differs by %.3g" % (k, e)
        want_done = bool(trace["terminated"][k] or trace["truncated"][k])
        if done or want_done:
            ended = k
            if tol["late"] is not None:
                assert k >= n - 2, "episode ended at step %d, recorded at %d" % (k, n - 1)
                if done and want_done:
                    assert bool(trunc) == bool(trace["truncated"][k])
            break
    return dict(steps=n, ended_at=ended, max_err_early=e_early, max_err_late=e_late, reset_err=e0,
                producer=trace["header"]["producer"])
