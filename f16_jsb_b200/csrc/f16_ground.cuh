// f16_ground.cuh - ground reactions of the F-16's seven STRUCTURE contacts (aircraft/f16/f16.xml:137-214)
// and the friction solve that follows them: FGGroundReactions::Run / FGLGear::GetBodyForces and
// FGAccelerations::CalculateFrictionForces of one FGFDMExec::Run(), against JSBSim's default terrain
// (WGS84 ellipsoid at elevation 0, at rest in ECEF).
//
// Where it sits: the reference terminates an episode when position/h-sl-meters < 10
// (jsbsim_gym/jsbsim_gym.py:240), and no contact point is further than 24.5 ft from the CG, so a
// contact can only touch inside the last env-step of a crashing episode - in roughly one crash out of
// fifteen under uniform random actions (steep or inverted impacts). fdm_frame() therefore calls this
// only when the CG is within kGroundReach of the ellipsoid; it is one cold, non-inlined function, in
// double for both precision modes, that redoes the FGAircraft + FGAccelerations block of the frame
// with the contact forces and the friction multipliers included. The three BOGEY contacts
// (f16.xml:86-136) are retractable and the reference keeps gear/gear-pos-norm at 0
// (jsbsim_gym.py:230-231): FGLGear's "gear up" branch, no force.
//
// Friction multipliers persist from frame to frame while a contact stays compressed (warm start of
// the projected Gauss-Seidel sweeps) and are zeroed when it is not; since contact is confined to
// one env-step they live in GroundMem, local to env_step_one, and cost no HBM state.
// Included by f16_model.cuh just ahead of fdm_frame (it uses that header's constants and MassSetT).
#pragma once

namespace f16 {

#if defined(__CUDACC__)
#define F16_COLD __host__ __device__ __noinline__ inline
#else
#define F16_COLD inline
#endif

constexpr int kNumStructure = f16data::n_structure;
static_assert(kNumStructure == 7, "MassSetT::r_ct holds seven contacts");
constexpr double kGroundReach = 26.0;   // ft; > max |StructuralToBody(contact)| = 24.5 ft (checked in build_mass_sets)
constexpr double kContactMargin = 0.05; // ft; flat-earth estimate of a contact's height: off by < 1e-4 ft within kGroundReach

struct GroundMem {
  double lm[3 * kNumStructure];   // per contact: roll, side (static friction), dynamic
  int started;                    // 0 until the first frame of this env-step that came within reach
};

struct GroundIn {
  double q[4], ri[3], vi[3], wi[3], epa;   // kinematic state after this frame's Propagate
  double F[3], M[3];                       // aerodynamic + propulsive force and moment about the CG, body axes
  double g_ec[3];                          // gravity, ECEF
  double dt;
};
struct GroundOut {
  double wdot[3], abody[3], ai0[3];        // vPQRidot, vBodyAccel, vUVWidot
};

struct GeoPoint { double h, n[3]; };       // geodetic altitude and the ellipsoid's outward unit normal (ECEF)

// FGLocation::ComputeDerivedUnconditional (Fukushima 2006) for one ECEF point: what
// FGDefaultGroundCallback::GetAGLevel returns for it.
F16_HD void geodetic_point(double x, double y, double z, GeoPoint& o) {
  const double rxy = sqrt(x * x + y * y);
  double sinLon = 0.0, cosLon = 1.0;
  if (rxy != 0.0) { sinLon = y / rxy; cosLon = x / rxy; }
  const double s0 = fabs(z);
  const double c = kEarthA * kE2;
  const double zc = kEc * s0, c0 = kEc * rxy;
  const double c02 = c0 * c0, s02 = s0 * s0;
  const double a02 = c02 + s02;
  const double a0 = sqrt(a02);
  const double a03 = a02 * a0;
  double s1 = zc * a03 + c * s02 * s0;
  const double c1 = rxy * a03 - c * c02 * c0;
  const double cs0c0 = c * c0 * s0;
  const double b0 = 1.5 * cs0c0 * ((rxy * s0 - zc * c0) * a0 - cs0c0);
  s1 = s1 * a03 - b0 * s0;
  const double cc = kEc * (c1 * a03 - b0 * c0);
  const double s12 = s1 * s1, cc2 = cc * cc;
  const double hyp = sqrt(s12 + cc2);
  o.h = (rxy * cc + s0 * s1 - kEarthA * sqrt(kEc2 * s12 + cc2)) / hyp;
  const double sinLat = (z >= 0.0 ? 1.0 : -1.0) * (s1 / hyp), cosLat = cc / hyp;
  o.n[0] = cosLat * cosLon; o.n[1] = cosLat * sinLon; o.n[2] = sinLat;
}

F16_HD double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
F16_HD void cross3(const double* a, const double* b, double* o) {
  o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
F16_HD void mat3_mul(const double (*m)[3], const double* a, double* o) {
  for (int i = 0; i < 3; ++i) o[i] = m[i][0] * a[0] + m[i][1] * a[1] + m[i][2] * a[2];
}
F16_HD void mat3_tmul(const double (*m)[3], const double* a, double* o) {   // m^T a
  for (int i = 0; i < 3; ++i) o[i] = m[0][i] * a[0] + m[1][i] * a[1] + m[2][i] * a[2];
}
F16_HD void normalize3(double* a) {   // FGColumnVector3::Normalize
  double m = sqrt(dot3(a, a));
  if (m != 0.0) { double t = 1.0 / m; a[0] *= t; a[1] *= t; a[2] *= t; }
}

// Returns false (and leaves `out` alone) when no contact is compressed in this frame: the frame's own
// accelerations stand.
F16_COLD bool ground_accelerations(const GroundIn& in, const MassSetT<double>& ms, GroundMem& mem, GroundOut& out) {
  const double k_spring[kNumStructure] = F16_STRUCT_SPRING, k_damp[kNumStructure] = F16_STRUCT_DAMPING;
  const double mu_s[kNumStructure] = F16_STRUCT_STATIC_F, mu_d[kNumStructure] = F16_STRUCT_DYNAMIC_F;
  if (!mem.started) {
    for (int i = 0; i < 3 * kNumStructure; ++i) mem.lm[i] = 0.0;
    mem.started = 1;
  }
  // ---- frames: ECEF position, Ti2b, Tec2b, Tec2l at the CG
  double se, ce;
  sincos(in.epa, &se, &ce);
  const double pe[3] = {ce * in.ri[0] + se * in.ri[1], -se * in.ri[0] + ce * in.ri[1], in.ri[2]};
  double b[3][3];
  {
    const double q0 = in.q[0], q1 = in.q[1], q2 = in.q[2], q3 = in.q[3];
    b[0][0] = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3; b[0][1] = 2.0 * (q1 * q2 + q0 * q3); b[0][2] = 2.0 * (q1 * q3 - q0 * q2);
    b[1][0] = 2.0 * (q1 * q2 - q0 * q3); b[1][1] = q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3; b[1][2] = 2.0 * (q2 * q3 + q0 * q1);
    b[2][0] = 2.0 * (q1 * q3 + q0 * q2); b[2][1] = 2.0 * (q2 * q3 - q0 * q1); b[2][2] = q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3;
  }
  double e2b[3][3];   // Tec2b = Ti2b * Tec2i
  for (int i = 0; i < 3; ++i) {
    e2b[i][0] = b[i][0] * ce + b[i][1] * se;
    e2b[i][1] = -b[i][0] * se + b[i][1] * ce;
    e2b[i][2] = b[i][2];
  }
  GeoPoint gc;
  geodetic_point(pe[0], pe[1], pe[2], gc);
  // local "down" axis of the CG's frame in ECEF = third row of Tec2l = -normal
  const double down_ec[3] = {-gc.n[0], -gc.n[1], -gc.n[2]};
  // body velocity and rates relative to the rotating earth
  const double wp[3] = {0.0, 0.0, kEarthOmega};
  double uvw[3], pqr[3], wpb[3];
  {
    double wxr[3];
    cross3(wp, in.ri, wxr);
    const double rel[3] = {in.vi[0] - wxr[0], in.vi[1] - wxr[1], in.vi[2] - wxr[2]};
    mat3_mul(b, rel, uvw);
    mat3_mul(b, wp, wpb);
    for (int i = 0; i < 3; ++i) pqr[i] = in.wi[i] - wpb[i];
  }

  // ---- FGLGear::GetBodyForces for each STRUCTURE contact
  double Fg[3] = {0, 0, 0}, Mg[3] = {0, 0, 0};
  double U[2 * kNumStructure][3], arm[2 * kNumStructure][3], lo[2 * kNumStructure], hi[2 * kNumStructure];
  int slot[2 * kNumStructure];
  int n = 0;
  for (int i = 0; i < kNumStructure; ++i) {
    const double rb[3] = {ms.r_ct[i][0], ms.r_ct[i][1], ms.r_ct[i][2]};   // Ts2b (vXYZn - vXYZcg), feet
    double off[3];
    mat3_tmul(e2b, rb, off);   // Tb2ec * r
    GeoPoint gp;
    gp.h = 1.0;
    // the exact height (FGLocation of the contact point) only for contacts the flat-earth estimate puts within
    // twice the margin of the surface; the others are in the air whichever way it is computed
    if (gc.h - dot3(down_ec, off) < 2.0 * kContactMargin) geodetic_point(pe[0] + off[0], pe[1] + off[1], pe[2] + off[2], gp);
    if (!(gp.h < 0.0)) {
      mem.lm[3 * i] = 0.0; mem.lm[3 * i + 1] = 0.0; mem.lm[3 * i + 2] = 0.0;
      continue;
    }
    double nb[3];
    mat3_mul(e2b, gp.n, nb);                       // vGroundNormal
    const double normalZ = dot3(down_ec, gp.n);
    const double comp = gp.h * normalZ / dot3(gp.n, gp.n);
    const double rc[3] = {rb[0] + comp * nb[0], rb[1] + comp * nb[1], rb[2] + comp * nb[2]};
    double vb[3];
    cross3(pqr, rc, vb);
    vb[0] += uvw[0]; vb[1] += uvw[1]; vb[2] += uvw[2];
    // ground frame: roll = body x projected on the ground plane, side = normal x body x
    const double nx = nb[0];
    double roll[3] = {1.0 - nx * nb[0], -nx * nb[1], -nx * nb[2]};
    double side[3] = {0.0, nb[2], -nb[1]};
    normalize3(roll);
    normalize3(side);
    const double vgx = dot3(roll, vb), vgy = dot3(side, vb), vgz = dot3(nb, vb);
    const double cspeed = -vgz;
    double strut = -comp * k_spring[i] + -cspeed * k_damp[i];
    if (strut > 0.0) strut = 0.0;
    const double N = -strut;
    const double fb[3] = {N * nb[0], N * nb[1], N * nb[2]};
    double mb[3];
    cross3(rc, fb, mb);
    for (int k = 0; k < 3; ++k) { Fg[k] += fb[k]; Mg[k] += mb[k]; }
    const double slip = sqrt(vgx * vgx + vgy * vgy);
    if (slip > 1e-3) {
      const double dx = vgx / slip, dy = vgy / slip;
      for (int k = 0; k < 3; ++k) { U[n][k] = roll[k] * dx + side[k] * dy; arm[n][k] = rc[k]; }
      hi[n] = 0.0; lo[n] = -fabs(mu_d[i] * N);
      slot[n] = 3 * i + 2;
      ++n;
    } else {
      const double mx = fabs(mu_s[i] * N);
      for (int k = 0; k < 3; ++k) { U[n][k] = roll[k]; arm[n][k] = rc[k]; U[n + 1][k] = side[k]; arm[n + 1][k] = rc[k]; }
      hi[n] = mx; lo[n] = -mx; slot[n] = 3 * i;
      hi[n + 1] = mx; lo[n + 1] = -mx; slot[n + 1] = 3 * i + 1;
      n += 2;
    }
  }
  if (n == 0) return false;   // a compressed contact always registers a multiplier
  for (int i = 0; i < n; ++i) {
    double v = mem.lm[slot[i]];
    mem.lm[slot[i]] = v < lo[i] ? lo[i] : (v > hi[i] ? hi[i] : v);
  }

  // ---- FGAircraft + FGAccelerations::CalculatePQRdot / CalculateUVWdot
  const double F[3] = {in.F[0] + Fg[0], in.F[1] + Fg[1], in.F[2] + Fg[2]};
  const double M[3] = {in.M[0] + Mg[0], in.M[1] + Mg[1], in.M[2] + Mg[2]};
  const double (*J)[3] = reinterpret_cast<const double (*)[3]>(ms.J);
  const double (*Ji)[3] = reinterpret_cast<const double (*)[3]>(ms.Jinv);
  double Jw[3], wJw[3], t[3], wdot_i[3], wdot_b[3], tmp[3];
  mat3_mul(J, in.wi, Jw);
  cross3(in.wi, Jw, wJw);
  for (int k = 0; k < 3; ++k) t[k] = M[k] - wJw[k];
  mat3_mul(Ji, t, wdot_i);
  cross3(in.wi, wpb, tmp);
  for (int k = 0; k < 3; ++k) wdot_b[k] = wdot_i[k] - tmp[k];
  double ab[3] = {F[0] / ms.mass, F[1] / ms.mass, F[2] / ms.mass};
  double vdot_b[3];
  {
    const double w2[3] = {pqr[0] + 2.0 * wpb[0], pqr[1] + 2.0 * wpb[1], pqr[2] + 2.0 * wpb[2]};
    double cor[3], wxr[3], wxwxr[3], cen[3], gb[3];
    cross3(w2, uvw, cor);
    cross3(wp, in.ri, wxr);
    cross3(wp, wxr, wxwxr);
    mat3_mul(b, wxwxr, cen);
    mat3_mul(e2b, in.g_ec, gb);
    for (int k = 0; k < 3; ++k) vdot_b[k] = ab[k] - cor[k] - cen[k] + gb[k];
  }

  // ---- FGAccelerations::CalculateFrictionForces: projected Gauss-Seidel, <= 50 sweeps
  {
    double a[2 * kNumStructure][2 * kNumStructure], rhs[2 * kNumStructure];
    for (int i = 0; i < n; ++i) {
      const double v1[3] = {U[i][0] / ms.mass, U[i][1] / ms.mass, U[i][2] / ms.mass};
      double rxu[3], v2[3];
      cross3(arm[i], U[i], rxu);
      mat3_mul(Ji, rxu, v2);
      for (int j = 0; j < i; ++j) a[i][j] = a[j][i];
      for (int j = i; j < n; ++j) {
        double v2xr[3];
        cross3(v2, arm[j], v2xr);
        const double s[3] = {v1[0] + v2xr[0], v1[1] + v2xr[1], v1[2] + v2xr[2]};
        a[i][j] = dot3(U[j], s);
      }
    }
    double vd[3] = {vdot_b[0], vdot_b[1], vdot_b[2]}, wd[3] = {wdot_b[0], wdot_b[1], wdot_b[2]};
    if (in.dt > 0.0)
      for (int k = 0; k < 3; ++k) { vd[k] += uvw[k] / in.dt; wd[k] += pqr[k] / in.dt; }
    for (int i = 0; i < n; ++i) {
      const double d = a[i][i];
      double wxr[3];
      cross3(wd, arm[i], wxr);
      const double s[3] = {vd[0] + wxr[0], vd[1] + wxr[1], vd[2] + wxr[2]};
      rhs[i] = -dot3(U[i], s) / d;
      for (int j = 0; j < n; ++j) a[i][j] /= d;
    }
    for (int iter = 0; iter < 50; ++iter) {
      double norm = 0.0;
      for (int i = 0; i < n; ++i) {
        const double lambda0 = mem.lm[slot[i]];
        double dl = rhs[i];
        for (int j = 0; j < n; ++j) dl -= a[i][j] * mem.lm[slot[j]];
        double v = lambda0 + dl;
        v = v < lo[i] ? lo[i] : (v > hi[i] ? hi[i] : v);
        mem.lm[slot[i]] = v;
        norm += fabs(v - lambda0);
      }
      if (norm < 1e-5) break;
    }
    double Ff[3] = {0, 0, 0}, Mf[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i) {
      const double lam = mem.lm[slot[i]];
      const double f[3] = {lam * U[i][0], lam * U[i][1], lam * U[i][2]};
      double m[3];
      cross3(arm[i], f, m);
      for (int k = 0; k < 3; ++k) { Ff[k] += f[k]; Mf[k] += m[k]; }
    }
    double od[3];
    mat3_mul(Ji, Mf, od);
    for (int k = 0; k < 3; ++k) { ab[k] += Ff[k] / ms.mass; wdot_i[k] += od[k]; }
  }
  // vUVWidot = Tb2i a_body + Tec2i g (the friction share of a_body included)
  double ai[3];
  mat3_tmul(b, ab, ai);
  ai[0] += ce * in.g_ec[0] - se * in.g_ec[1];
  ai[1] += se * in.g_ec[0] + ce * in.g_ec[1];
  ai[2] += in.g_ec[2];
  for (int k = 0; k < 3; ++k) { out.wdot[k] = wdot_i[k]; out.abody[k] = ab[k]; out.ai0[k] = ai[k]; }
  return true;
}

}  // namespace f16
