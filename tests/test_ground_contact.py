"""Ground reactions (aircraft/f16/f16.xml:85-215): the kernel source's cold path
(f16_ground.cuh, compiled with g++ by tests/hostsim) against the oracle's FGLGear / FGGroundReactions /
friction-solver restatement, teacher-forced over one env-step (4 frames).

Contact can only happen inside the last env-step of a crashing episode (termination at h < 10 m,
jsbsim_gym.py:240; no contact point is further than 24.5 ft from the CG), so the cases are (1) crash
steps of random-action episodes - the radome of a steep dive - and (2) synthetic states placed a few
feet above the ellipsoid at chosen attitudes and speeds so that every STRUCTURE contact, several
contacts at once and both friction branches (dynamic: one multiplier per contact; static: two) are
exercised.

Tolerances: FP64 <= 1e-8 relative per state field (floors of conftest; measured 8e-10 - the compression
is a difference of 2e7-ft numbers, so its last bits differ between the two codings), FP32 <= 1e-3.
The same cases run on the GPU in test_gpu_parity.py.
"""
import numpy as np
import pytest

from conftest import state_floors

NF = 53
A_FT = 20925646.32546          # WGS84 semi-major axis, ft
OMEGA = 0.00007292115          # earth rate, rad/s


def rel_err(a, ref, floors):
    return np.abs(a - ref) / np.maximum(np.abs(ref), floors)


def quat_from_matrix(T):
    """FGMatrix33::GetQuaternion, largest-diagonal branch."""
    tq = [1 + T[0, 0] + T[1, 1] + T[2, 2], 1 + T[0, 0] - T[1, 1] - T[2, 2],
          1 - T[0, 0] + T[1, 1] - T[2, 2], 1 - T[0, 0] - T[1, 1] + T[2, 2]]
    i = int(np.argmax(tq))
    q = np.zeros(4)
    if i == 0:
        q[0] = 0.5 * np.sqrt(tq[0]); q[1] = 0.25 * (T[1, 2] - T[2, 1]) / q[0]
        q[2] = 0.25 * (T[2, 0] - T[0, 2]) / q[0]; q[3] = 0.25 * (T[0, 1] - T[1, 0]) / q[0]
    elif i == 1:
        q[1] = 0.5 * np.sqrt(tq[1]); q[0] = 0.25 * (T[1, 2] - T[2, 1]) / q[1]
        q[2] = 0.25 * (T[0, 1] + T[1, 0]) / q[1]; q[3] = 0.25 * (T[2, 0] + T[0, 2]) / q[1]
    elif i == 2:
        q[2] = 0.5 * np.sqrt(tq[2]); q[0] = 0.25 * (T[2, 0] - T[0, 2]) / q[2]
        q[1] = 0.25 * (T[0, 1] + T[1, 0]) / q[2]; q[3] = 0.25 * (T[1, 2] + T[2, 1]) / q[2]
    else:
        q[3] = 0.5 * np.sqrt(tq[3]); q[0] = 0.25 * (T[0, 1] - T[1, 0]) / q[3]
        q[1] = 0.25 * (T[0, 2] + T[2, 0]) / q[3]; q[2] = 0.25 * (T[1, 2] + T[2, 1]) / q[3]
    return q


def synthetic_state(base, fields, h_ft, phi, theta, psi, uvw, pqr=(0.0, 0.0, 0.0)):
    """Packed state at lat = lon = 0, epa = 0, geodetic altitude h_ft, Euler attitude, body velocity.
    Everything not kinematic (FCS memories, engine) is kept from `base` (a state in flight)."""
    ix = {n: i for i, n in enumerate(fields)}
    s = np.array(base, dtype=np.float64)
    cph, sph, cth, sth, cps, sps = np.cos(phi), np.sin(phi), np.cos(theta), np.sin(theta), np.cos(psi), np.sin(psi)
    Tl2b = np.array([[cth * cps, cth * sps, -sth],
                     [sph * sth * cps - cph * sps, sph * sth * sps + cph * cps, sph * cth],
                     [cph * sth * cps + sph * sps, cph * sth * sps - sph * cps, cph * cth]])
    Ti2l = np.array([[0.0, 0.0, 1.0], [0.0, 1.0, 0.0], [-1.0, 0.0, 0.0]])     # N, E, D at lat = lon = 0, epa = 0
    Ti2b = Tl2b @ Ti2l
    q = quat_from_matrix(Ti2b)
    ri = np.array([A_FT + h_ft, 0.0, 0.0])
    wp = np.array([0.0, 0.0, OMEGA])
    vi = Ti2b.T @ np.asarray(uvw, float) + np.cross(wp, ri)
    wi = np.asarray(pqr, float) + Ti2b @ wp
    s[ix["Q0"]:ix["Q0"] + 4] = q
    s[ix["WI_X"]:ix["WI_X"] + 3] = wi
    s[ix["RI_X"]:ix["RI_X"] + 3] = ri
    s[ix["VI_X"]:ix["VI_X"] + 3] = vi
    s[ix["EPA"]] = 0.0
    s[ix["VI1_X"]:ix["VI1_X"] + 3] = vi
    s[ix["VI2_X"]:ix["VI2_X"] + 3] = vi
    for nm in ("AI0_X", "AI1_X", "WDOT_X", "ABODY_X"):
        s[ix[nm]:ix[nm] + 3] = 0.0
    s[ix["PQR_X"]:ix["PQR_X"] + 3] = pqr
    V = float(np.linalg.norm(uvw))
    s[ix["ALPHA"]] = np.arctan2(uvw[2], uvw[0]) if V > 1e-3 else 0.0
    s[ix["MACH"]] = V / 1116.45
    s[ix["VC_KTS"]] = V * 0.5924838
    s[ix["VG"]] = V
    return s


def flying_oracle_env(oracle):
    """Oracle env in flight configuration (gear up, tanks at 1000 lb, engine running, steady CG)."""
    env = oracle.OracleEnv()
    goal = oracle.sample_goal(0)
    env.reset(goal)
    for _ in range(3):
        env.step(np.array([0, 0, 0, 0.5], np.float32))
    return env, goal


# name, h_ft, phi, theta, psi, uvw, pqr, contacts expected to touch in the first frame (file order index)
CASES = [
    ("belly_slow_sink", 2.0, 0.0, 0.0, 0.3, (150.0, 0.0, 8.0), (0.0, 0.0, 0.0), {6, 7, 8}),          # ventral fins + intake
    ("inverted_fin", 9.0, np.pi, 0.05, 1.0, (400.0, 5.0, -20.0), (0.1, 0.0, 0.0), {5}),             # top of the vertical stabiliser
    ("knife_edge_wingtip", 12.0, 1.45, 0.0, 2.0, (500.0, 0.0, 10.0), (0.0, 0.1, 0.0), {4}),          # right wing tip
    ("left_wingtip", 12.0, -1.45, 0.0, 4.0, (500.0, 0.0, 10.0), (0.0, 0.0, 0.05), {3}),
    ("steep_dive_radome", 20.0, 0.2, -1.3, 5.0, (700.0, 0.0, 30.0), (0.0, -0.2, 0.0), {9}),
    ("at_rest_static_friction", 2.0, 0.0, 0.0, 0.0, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), {6, 7, 8}),    # static branch: 2 multipliers each
    ("deep_all_contacts", -16.0, 0.4, -0.3, 0.7, (300.0, 20.0, 40.0), (0.3, 0.2, -0.1), None),
]


def test_unpack_state_round_trip(oracle, golden):
    """unpack(pack(x)) continues the trajectory exactly: the packed fields are all that must survive."""
    t = golden["random0"]
    env, goal = flying_oracle_env(oracle)
    for k in (5, 200, 400):
        env.fdm.unpack_state(t["states"][k])
        env.step(t["actions"][k])
        got = env.fdm.pack_state()
        assert np.array_equal(got, t["states"][k + 1]) or np.abs(got - t["states"][k + 1]).max() < 1e-12 * np.abs(t["states"][k + 1]).max()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("mode,tol", [(0, 1e-8), (1, 1e-3)])
def test_synthetic_contact_step_parity(oracle, hostsim, state_fields, case, mode, tol):
    name, h, phi, th, psi, uvw, pqr, want = case
    floors = state_floors(state_fields)
    env, goal = flying_oracle_env(oracle)
    s0 = synthetic_state(env.fdm.pack_state(), state_fields, h, phi, th, psi, uvw, pqr)
    act = np.array([0.2, -0.1, 0.1, 0.7], np.float32)
    # first frame alone: which contacts touch, and the reaction must push away from the ground
    env.fdm.unpack_state(s0)
    for p, v in (("propulsion/tank/contents-lbs", 1000), ("propulsion/tank[1]/contents-lbs", 1000), ("gear/gear-cmd-norm", 0), ("gear/gear-pos-norm", 0)):
        env.fdm[p] = v
    env.fdm.run()
    wow = {i for i in range(10) if env.fdm["gear/unit[%d]/WOW" % i] > 0.5}
    assert not (wow & {0, 1, 2}), "a retracted BOGEY must not touch"
    if want is not None:
        assert wow == want, wow
    else:
        assert len(wow) == 7
    F = np.array([env.fdm["forces/fb%s-gear-lbs" % a] for a in "xyz"])
    phi_, th_ = env.fdm["attitude/phi-rad"], env.fdm["attitude/theta-rad"]
    down = np.array([-np.sin(th_), np.sin(phi_) * np.cos(th_), np.cos(phi_) * np.cos(th_)])
    assert F @ down < 0.0
    nm = env.fdm["gear/num-friction-multipliers"]
    assert nm == (2 * len(wow) if name == "at_rest_static_friction" else len(wow))
    # whole env-step, teacher-forced on both sides
    env.fdm.unpack_state(s0)
    env.step(act)
    s1 = env.fdm.pack_state()
    hs = hostsim.env(mode)
    hs.reset(goal)
    hs.set_state(s0, current_step=5)
    hs.step(act)
    e = rel_err(hs.get_state(NF), s1, floors)
    if mode == 1 and name == "at_rest_static_friction":
        # stated tolerance of the float mode for a STANDING aircraft: calibrated airspeed within 0.5 kt absolute. float32
        # cannot form pow(1 + 1e-7, 1/3.5) - 1; the value only feeds FCS thresholds at 5 kt and above, and no state of the
        # reference's flights is slower than 100 kt
        i_vc = state_fields.index("VC_KTS")
        assert abs(hs.get_state(NF)[i_vc] - s1[i_vc]) <= 0.5
        e[i_vc] = min(e[i_vc], tol / 2)
    assert e.max() < tol, (name, state_fields[int(e.argmax())], float(e.max()))
    # the contact must have mattered: without it the accelerations are completely different
    assert np.abs(s1[state_fields.index("WDOT_X"):state_fields.index("WDOT_X") + 3]).max() > 1e-3


def crash_steps(oracle, n_episodes=40, first_seed=5000):
    """(goal, state before the terminal step, action, state after, frame, contact flags) of random-action crashes."""
    out = []
    for ep in range(first_seed, first_seed + n_episodes):
        rng = np.random.default_rng(ep)
        env = oracle.OracleEnv()
        goal = oracle.sample_goal(ep)
        env.reset(goal)
        for t in range(1200):
            a = rng.uniform([-1, -1, -1, 0], [1, 1, 1, 1]).astype(np.float32)
            s0 = env.fdm.pack_state()
            obs, r, term, trunc = env.step(a)
            if term or trunc:
                break
        if term and obs[-1][2] < 10.0:
            wow = [env.fdm["gear/unit[%d]/WOW" % i] > 0.5 for i in range(10)]
            out.append((goal, s0, t, a, env.fdm.pack_state(), obs[-1].copy(), np.float32(r), wow))
    return out


@pytest.fixture(scope="module")
def crashes(oracle):
    return crash_steps(oracle)


@pytest.mark.parametrize("mode,tol", [(0, 1e-8), (1, 1e-3)])
def test_crash_step_parity_random_episodes(oracle, hostsim, state_fields, crashes, mode, tol):
    floors = state_floors(state_fields)
    touched = 0
    for goal, s0, t, a, s1, frame, reward, wow in crashes:
        hs = hostsim.env(mode)
        hs.reset(goal)
        hs.set_state(s0, current_step=t)
        obs, r, fl, _ = hs.step(a)
        e = rel_err(hs.get_state(NF), s1, floors)
        assert e.max() < tol, (t, state_fields[int(e.argmax())], float(e.max()), wow)
        assert fl & 8 and fl & 32                      # done, crash
        if mode == 0:
            assert np.array_equal(obs[-1][:12], frame[:12])
        touched += any(wow)
    assert len(crashes) >= 20 and touched >= 2         # steep impacts do reach the ground within the step
