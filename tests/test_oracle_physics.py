"""First-principles checks of the restated 6-DOF core - pins that need no JSBSim.

The oracle cannot be compared with the real `jsbsim` package here, but a rigid-body simulation has to
satisfy identities that do not depend on whose code it is. Over 600 frames of one random-action flight (rates up to
1.4 rad/s, accelerations up to 4 g) the properties the oracle serves are checked against each other:

  * Euler-angle kinematics: d(phi, theta, psi)/dt from consecutive frames = the textbook functions of (p, q, r, phi, theta)
    - quaternion propagation, the local frame, Euler extraction and the definition of p, q, r are mutually consistent
    (residual: the earth rate and the transport rate, ~7e-5 rad/s);
  * position kinematics: d(h, lat, lon)/dt = the NED velocity obtained by rotating (u, v, w) with the Euler angles;
  * rotational dynamics: (w[k+1] - w[k]) / dt = J^-1 (M - w x J w) with the moments and inertias of frame k
    (rectangular Euler, so this holds frame by frame; it also fixes the sign convention of Ixz);
  * translational dynamics: the change of the NED velocity = the Adams-Bashforth-2 combination of (F_aero + F_prop) / m +
    gravity of the last two frames (residual: the rotating-earth terms, ~0.25 ft/s2).

The CUDA kernel is held to the oracle by the parity tests, so these identities carry over to it.
"""
import numpy as np
import pytest

DT = 1.0 / 120.0
R_EQUATOR_FT = 20925646.32546

NAMES = ["attitude/phi-rad", "attitude/theta-rad", "attitude/psi-rad", "velocities/p-rad_sec", "velocities/q-rad_sec",
         "velocities/r-rad_sec", "velocities/u-fps", "velocities/v-fps", "velocities/w-fps", "position/h-sl-ft", "position/lat-gc-rad",
         "position/long-gc-rad", "forces/fbx-aero-lbs", "forces/fby-aero-lbs", "forces/fbz-aero-lbs", "forces/fbx-prop-lbs",
         "moments/l-aero-lbsft", "moments/m-aero-lbsft", "moments/n-aero-lbsft", "moments/m-prop-lbsft", "inertia/mass-slugs",
         "inertia/ixx-slugs_ft2", "inertia/iyy-slugs_ft2", "inertia/izz-slugs_ft2", "inertia/ixz-slugs_ft2",
         "accelerations/gravity-ft_sec2", "aero/alpha-rad", "aero/beta-rad", "velocities/mach", "velocities/vt-fps", "atmosphere/a-fps",
         "atmosphere/rho-slugs_ft3", "aero/qbar-psf", "atmosphere/T-R", "atmosphere/P-psf",
         "accelerations/n-pilot-y-norm", "accelerations/n-pilot-z-norm", "inertia/cg-x-in", "inertia/cg-y-in", "inertia/cg-z-in"]


@pytest.fixture(scope="module")
def flight(oracle):
    env = oracle.OracleEnv()
    env.reset(np.array([5000.0, 0.0, 2000.0], np.float32))
    f = env.fdm
    rng = np.random.default_rng(1)
    hist = []
    for k in range(600):
        if k % 4 == 0:          # a new action every env-step, as the env does (jsbsim_gym.py:216-232)
            a = rng.uniform([-1, -1, -1, 0], [1, 1, 1, 1])
            for n, v in zip(("fcs/aileron-cmd-norm", "fcs/elevator-cmd-norm", "fcs/rudder-cmd-norm", "fcs/throttle-cmd-norm"), a):
                f.set_property_value(n, float(v))
        f.set_property_value("propulsion/tank/contents-lbs", 1000.0)
        f.set_property_value("propulsion/tank[1]/contents-lbs", 1000.0)
        f.set_property_value("gear/gear-cmd-norm", 0.0)
        f.set_property_value("gear/gear-pos-norm", 0.0)
        f.run()
        hist.append({n: f[n] for n in NAMES})
    assert max(abs(h["velocities/p-rad_sec"]) for h in hist) > 0.5          # the flight is not a trivial one
    return hist


def _wrap(x):
    return (x + np.pi) % (2 * np.pi) - np.pi


def _tb2l(h):
    ph, th, ps = h["attitude/phi-rad"], h["attitude/theta-rad"], h["attitude/psi-rad"]
    cph, sph, cth, sth, cps, sps = np.cos(ph), np.sin(ph), np.cos(th), np.sin(th), np.cos(ps), np.sin(ps)
    return np.array([[cth * cps, sph * sth * cps - cph * sps, cph * sth * cps + sph * sps],
                     [cth * sps, sph * sth * sps + cph * cps, cph * sth * sps - sph * cps],
                     [-sth, sph * cth, cph * cth]])


def _uvw(h):
    return np.array([h["velocities/u-fps"], h["velocities/v-fps"], h["velocities/w-fps"]])


def _pqr(h):
    return np.array([h["velocities/p-rad_sec"], h["velocities/q-rad_sec"], h["velocities/r-rad_sec"]])


def test_euler_angle_kinematics(flight):
    worst = np.zeros(3)
    for a, b in zip(flight[2:-1], flight[3:]):
        ph = 0.5 * (a["attitude/phi-rad"] + b["attitude/phi-rad"])
        th = 0.5 * (a["attitude/theta-rad"] + b["attitude/theta-rad"])
        if abs(np.cos(th)) < 0.2:
            continue
        p, q, r = _pqr(a)
        rates = np.array([p + np.tan(th) * (q * np.sin(ph) + r * np.cos(ph)), q * np.cos(ph) - r * np.sin(ph),
                          (q * np.sin(ph) + r * np.cos(ph)) / np.cos(th)])
        fd = np.array([_wrap(b["attitude/phi-rad"] - a["attitude/phi-rad"]), b["attitude/theta-rad"] - a["attitude/theta-rad"],
                       _wrap(b["attitude/psi-rad"] - a["attitude/psi-rad"])]) / DT
        worst = np.maximum(worst, np.abs(fd - rates))
    assert worst.max() < 2e-4, worst           # measured 5e-5: the earth rate


def test_position_kinematics(flight):
    worst = np.zeros(3)
    for a, b in zip(flight[2:-1], flight[3:]):
        v = 0.5 * (_tb2l(a) @ _uvw(a) + _tb2l(b) @ _uvw(b))          # NED, ft/s
        r = R_EQUATOR_FT + a["position/h-sl-ft"]
        fd = np.array([(b["position/lat-gc-rad"] - a["position/lat-gc-rad"]) / DT * r,
                       (b["position/long-gc-rad"] - a["position/long-gc-rad"]) / DT * r * np.cos(a["position/lat-gc-rad"]),
                       -(b["position/h-sl-ft"] - a["position/h-sl-ft"]) / DT])
        worst = np.maximum(worst, np.abs(fd - v))
    assert worst.max() < 1.0, worst            # ft/s of ~900: measured 0.3 (trapezoid against AB3, geocentric against geodetic)


def test_rotational_dynamics(flight):
    worst, biggest = 0.0, 0.0
    for a, b in zip(flight[2:-1], flight[3:]):
        ixz = a["inertia/ixz-slugs_ft2"]
        J = np.array([[a["inertia/ixx-slugs_ft2"], 0, -ixz], [0, a["inertia/iyy-slugs_ft2"], 0], [-ixz, 0, a["inertia/izz-slugs_ft2"]]])
        w = _pqr(a)
        M = np.array([a["moments/l-aero-lbsft"], a["moments/m-aero-lbsft"] + a["moments/m-prop-lbsft"], a["moments/n-aero-lbsft"]])
        wdot = np.linalg.solve(J, M - np.cross(w, J @ w))
        worst = max(worst, np.abs((_pqr(b) - w) / DT - wdot).max())
        biggest = max(biggest, np.abs(wdot).max())
    assert biggest > 1.0 and worst < 5e-4, (worst, biggest)      # rad/s2; measured 6e-5 against accelerations of several rad/s2


def test_translational_dynamics(flight):
    def accel(h):        # specific force + gravity, body axes
        F = np.array([h["forces/fbx-aero-lbs"] + h["forces/fbx-prop-lbs"], h["forces/fby-aero-lbs"], h["forces/fbz-aero-lbs"]]) / h["inertia/mass-slugs"]
        ph, th, g = h["attitude/phi-rad"], h["attitude/theta-rad"], h["accelerations/gravity-ft_sec2"]
        return F + g * np.array([-np.sin(th), np.sin(ph) * np.cos(th), np.cos(ph) * np.cos(th)])

    worst, biggest = 0.0, 0.0
    for z, a, b in zip(flight[2:-2], flight[3:-1], flight[4:]):
        fd = (_tb2l(b) @ _uvw(b) - _tb2l(a) @ _uvw(a)) / DT
        pred = 1.5 * (_tb2l(a) @ accel(a)) - 0.5 * (_tb2l(z) @ accel(z))        # Adams-Bashforth 2 (FGPropagate's default for the velocity)
        worst = max(worst, np.abs(fd - pred).max())
        biggest = max(biggest, np.abs(pred).max())
    assert biggest > 60.0 and worst < 0.5, (worst, biggest)       # ft/s2; measured 0.15: Coriolis + centrifugal + transport terms


def test_air_data_identities(flight):
    """alpha = atan2(w, u), beta = atan2(v, sqrt(u2 + w2)), Vt = |uvw| (no wind), Mach = Vt / a, qbar = rho Vt2 / 2,
    a = sqrt(gamma R T), rho = P / (R T) - the Auxiliary and atmosphere outputs agree with the state they are made from."""
    R_AIR = 1716.5574933  # ft lbf / (slug R): 8.31432 J/(mol K) / 28.9645 g/mol in engineering units
    for h in flight[1:]:
        u, v, w = _uvw(h)
        vt = np.sqrt(u * u + v * v + w * w)
        assert h["aero/alpha-rad"] == pytest.approx(np.arctan2(w, u), abs=1e-12)
        assert h["aero/beta-rad"] == pytest.approx(np.arctan2(v, np.hypot(u, w)), abs=1e-12)
        assert h["velocities/vt-fps"] == pytest.approx(vt, rel=1e-12)
        assert h["velocities/mach"] == pytest.approx(vt / h["atmosphere/a-fps"], rel=1e-12)
        assert h["aero/qbar-psf"] == pytest.approx(0.5 * h["atmosphere/rho-slugs_ft3"] * vt * vt, rel=1e-12)
        assert h["atmosphere/a-fps"] == pytest.approx(np.sqrt(1.4 * R_AIR * h["atmosphere/T-R"]), rel=1e-6)
        assert h["atmosphere/rho-slugs_ft3"] == pytest.approx(h["atmosphere/P-psf"] / (R_AIR * h["atmosphere/T-R"]), rel=1e-6)


def test_gravity_at_the_initial_condition(oracle):
    """J2 gravity on the equator, 5 000 ft above the WGS84 ellipsoid: GM / r2 * (1 + 1.5 J2 (a / r)2), evaluated here from the
    published constants (GM = 3.986004418e14 m3/s2, a = 6 378 137 m, J2 = 1.08262982e-3)."""
    env = oracle.OracleEnv()
    env.reset(np.array([5000.0, 0.0, 2000.0], np.float32))
    ft = 0.3048
    gm, a, j2 = 3.986004418e14 / ft ** 3, 6378137.0 / ft, 1.08262982e-3
    r = a + 5000.0
    assert env.fdm["accelerations/gravity-ft_sec2"] == pytest.approx(gm / r ** 2 * (1 + 1.5 * j2 * (a / r) ** 2), rel=1e-9)


def test_pilot_accelerations(flight):
    """The load factors the pitch and yaw control laws feed on (aircraft/f16/f16.xml:502-761) are the specific force at the
    eyepoint (f16.xml:50-54): n = [F / m + wdot x r + w x (w x r)] / g0 with r from the CG to the eyepoint in body axes -
    formed, as FGAuxiliary does, from the forces and moments of the PREVIOUS frame and this frame's rates. z points down, so
    lift (a negative body-z force) gives a negative n-pilot-z: the identity below fixes the sign as well."""
    eye = np.array([-336.2, 0.0, 29.5])                           # structural frame, inches
    g0 = 9.80665 / 0.3048
    worst = 0.0
    for z, a in zip(flight[2:-1], flight[3:]):
        cg = np.array([z["inertia/cg-x-in"], z["inertia/cg-y-in"], z["inertia/cg-z-in"]])
        r = (eye - cg) / 12.0 * np.array([-1.0, 1.0, -1.0])       # structural -> body axes, feet
        ixz = z["inertia/ixz-slugs_ft2"]
        J = np.array([[z["inertia/ixx-slugs_ft2"], 0, -ixz], [0, z["inertia/iyy-slugs_ft2"], 0], [-ixz, 0, z["inertia/izz-slugs_ft2"]]])
        M = np.array([z["moments/l-aero-lbsft"], z["moments/m-aero-lbsft"] + z["moments/m-prop-lbsft"], z["moments/n-aero-lbsft"]])
        wz = _pqr(z)
        wdot = np.linalg.solve(J, M - np.cross(wz, J @ wz))
        F = np.array([z["forces/fbx-aero-lbs"] + z["forces/fbx-prop-lbs"], z["forces/fby-aero-lbs"], z["forces/fbz-aero-lbs"]]) / z["inertia/mass-slugs"]
        w = _pqr(a)
        n = (F + np.cross(wdot, r) + np.cross(w, np.cross(w, r))) / g0
        worst = max(worst, abs(a["accelerations/n-pilot-z-norm"] - n[2]), abs(a["accelerations/n-pilot-y-norm"] - n[1]))
    assert worst < 1e-4, worst                                    # g; measured 2e-5 (inertial against earth-relative rates)
