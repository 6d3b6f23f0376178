// =====================================================================================
// f16_oracle.cpp - CPU restatement of the reference's F-16 env-step path.
//
// TEST INFRASTRUCTURE ONLY. Nothing under f16_jsb_b200/ may include, link or call this file;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
//
// PARITY STATUS: **parity unpinned**. The arithmetic of the reference path lives in the
// third-party `jsbsim` PyPI package (reference requirements.txt:4, unpinned; imported at
// jsbsim_gym/jsbsim_gym.py:1), which is neither vendored under /root/reference nor installable in
// the build container, and the reference ships no test, golden vector or fixture for this path.
// What follows restates JSBSim's published algorithm (v1.1 - v1.2 line: FGFDMExec::Run and the
// models it schedules) for the call sites the reference actually uses:
//   jsbsim_gym/jsbsim_gym.py:151-155  FGFDMExec(root, None) / load_model('f16') / run_ic()
//   jsbsim_gym/jsbsim_gym.py:166-170  propulsion/set-running, ic/u-fps, ic/h-sl-ft
//   jsbsim_gym/jsbsim_gym.py:181-182  the 12 STATE_FORMAT property reads
//   jsbsim_gym/jsbsim_gym.py:219-232  fcs/*-cmd-norm, tank contents, gear cmd/pos, run()
//   jsbsim_gym/jsbsim_gym.py:305-306  run_ic() + propulsion/set-running on reset
// and the aircraft data it loads: aircraft/f16/f16.xml, aircraft/f16/Engines/F100-PW-229.xml
// (rendered into f16_oracle_gen.inc by tools/gen_model.py, component by component in file order).
// Pins that do exist: analytic known-answer tests (ISA-1976, table lookups, mass properties, IC
// flight condition; tests/test_oracle_kat.py) and the reference's own Python env layer executed
// unmodified on top of this FDM (tools/make_golden.py -> tests/golden/).
//
// Structure mirrors JSBSim: one struct per model with an explicit `in` copy filled by
// load_inputs(model) right before that model runs, so the frame-to-frame staleness pattern of
// FGFDMExec::LoadInputs is reproduced by construction.
// =====================================================================================
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

// ------------------------------------------------------------------ FGJSBBase constants
constexpr double fttom = 0.3048;
constexpr double inchtoft = 1.0 / 12.0;
constexpr double slugtolb = 32.174049;
constexpr double lbtoslug = 1.0 / slugtolb;
constexpr double kgtoslug = 0.06852168;
constexpr double radtodeg = 180.0 / M_PI;
constexpr double degtorad = M_PI / 180.0;
constexpr double ktstofps = 1852.0 / (3600.0 * fttom);
constexpr double fpstokts = 1.0 / ktstofps;

// ------------------------------------------------------------------ FGColumnVector3 / FGMatrix33 subset
struct V3 {
  double v[3];
  V3() : v{0, 0, 0} {}
  V3(double a, double b, double c) : v{a, b, c} {}
  double& operator()(int i) { return v[i - 1]; }        // 1-based like JSBSim
  double operator()(int i) const { return v[i - 1]; }
  V3 operator+(const V3& o) const { return V3(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
  V3 operator-(const V3& o) const { return V3(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
  V3& operator+=(const V3& o) { v[0] += o.v[0]; v[1] += o.v[1]; v[2] += o.v[2]; return *this; }
  V3& operator-=(const V3& o) { v[0] -= o.v[0]; v[1] -= o.v[1]; v[2] -= o.v[2]; return *this; }
  V3 operator*(double s) const { return V3(s * v[0], s * v[1], s * v[2]); }
  V3 operator/(double s) const { double t = 1.0 / s; return V3(v[0] * t, v[1] * t, v[2] * t); }  // FGColumnVector3::operator/
  // cross product (FGColumnVector3::operator*(const FGColumnVector3&))
  V3 operator*(const V3& o) const {
    return V3(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]);
  }
  double Magnitude() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
  double Magnitude(int a, int b) const { return std::sqrt(v[a - 1] * v[a - 1] + v[b - 1] * v[b - 1]); }
};
inline V3 operator*(double s, const V3& a) { return a * s; }

struct M33 {
  double m[3][3];
  M33() { std::memset(m, 0, sizeof(m)); }
  M33(double a, double b, double c, double d, double e, double f, double g, double h, double i)
      : m{{a, b, c}, {d, e, f}, {g, h, i}} {}
  double& operator()(int r, int c) { return m[r - 1][c - 1]; }
  double operator()(int r, int c) const { return m[r - 1][c - 1]; }
  V3 operator*(const V3& a) const {
    return V3(m[0][0] * a.v[0] + m[0][1] * a.v[1] + m[0][2] * a.v[2],
              m[1][0] * a.v[0] + m[1][1] * a.v[1] + m[1][2] * a.v[2],
              m[2][0] * a.v[0] + m[2][1] * a.v[1] + m[2][2] * a.v[2]);
  }
  M33 operator*(const M33& o) const {
    M33 r;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) r.m[i][j] = m[i][0] * o.m[0][j] + m[i][1] * o.m[1][j] + m[i][2] * o.m[2][j];
    return r;
  }
  M33& operator+=(const M33& o) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) m[i][j] += o.m[i][j];
    return *this;
  }
  M33 Transposed() const {
    return M33(m[0][0], m[1][0], m[2][0], m[0][1], m[1][1], m[2][1], m[0][2], m[1][2], m[2][2]);
  }
};

// FGQuaternion subset: data[0..3] = (q0, q1, q2, q3), scalar first.
struct Quat {
  double d[4];
  Quat() : d{1, 0, 0, 0} {}
  Quat(double a, double b, double c, double e) : d{a, b, c, e} {}
  double Magnitude() const { return std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + d[3] * d[3]); }
  // FGQuaternion::Normalize: leaves the value alone when |q| is within 1e-10 of one.
  void Normalize() {
    double norm = Magnitude();
    if (norm == 0.0 || std::fabs(norm - 1.000) < 1e-10) return;
    double rnorm = 1.0 / norm;
    for (double& x : d) x *= rnorm;
  }
  // Stevens & Lewis eqn 1.3-32 (FGQuaternion::ComputeDerivedUnconditional)
  M33 GetT() const {
    double q0 = d[0], q1 = d[1], q2 = d[2], q3 = d[3];
    double q0q0 = q0 * q0, q1q1 = q1 * q1, q2q2 = q2 * q2, q3q3 = q3 * q3;
    double q0q1 = q0 * q1, q0q2 = q0 * q2, q0q3 = q0 * q3, q1q2 = q1 * q2, q1q3 = q1 * q3, q2q3 = q2 * q3;
    return M33(q0q0 + q1q1 - q2q2 - q3q3, 2.0 * (q1q2 + q0q3), 2.0 * (q1q3 - q0q2),
               2.0 * (q1q2 - q0q3), q0q0 - q1q1 + q2q2 - q3q3, 2.0 * (q2q3 + q0q1),
               2.0 * (q1q3 + q0q2), 2.0 * (q2q3 - q0q1), q0q0 - q1q1 - q2q2 + q3q3);
  }
  // FGQuaternion::GetQDot
  Quat GetQDot(const V3& PQR) const {
    return Quat(-0.5 * (d[1] * PQR(1) + d[2] * PQR(2) + d[3] * PQR(3)),
                0.5 * (d[0] * PQR(1) - d[3] * PQR(2) + d[2] * PQR(3)),
                0.5 * (d[3] * PQR(1) + d[0] * PQR(2) - d[1] * PQR(3)),
                0.5 * (-d[2] * PQR(1) + d[1] * PQR(2) + d[0] * PQR(3)));
  }
  // quaternion product (FGQuaternion::operator*)
  Quat operator*(const Quat& q) const {
    return Quat(d[0] * q.d[0] - d[1] * q.d[1] - d[2] * q.d[2] - d[3] * q.d[3],
                d[0] * q.d[1] + d[1] * q.d[0] + d[2] * q.d[3] - d[3] * q.d[2],
                d[0] * q.d[2] - d[1] * q.d[3] + d[2] * q.d[0] + d[3] * q.d[1],
                d[0] * q.d[3] + d[1] * q.d[2] - d[2] * q.d[1] + d[3] * q.d[0]);
  }
};

// FGMatrix33::GetQuaternion (largest-diagonal branch selection)
Quat MatrixToQuat(const M33& T) {
  double tq[4];
  tq[0] = 1.0 + T(1, 1) + T(2, 2) + T(3, 3);
  tq[1] = 1.0 + T(1, 1) - T(2, 2) - T(3, 3);
  tq[2] = 1.0 - T(1, 1) + T(2, 2) - T(3, 3);
  tq[3] = 1.0 - T(1, 1) - T(2, 2) + T(3, 3);
  int idx = 0;
  for (int i = 1; i < 4; i++)
    if (tq[i] > tq[idx]) idx = i;
  Quat Q;
  switch (idx) {
    case 0:
      Q.d[0] = 0.50 * std::sqrt(tq[0]);
      Q.d[1] = 0.25 * (T(2, 3) - T(3, 2)) / Q.d[0];
      Q.d[2] = 0.25 * (T(3, 1) - T(1, 3)) / Q.d[0];
      Q.d[3] = 0.25 * (T(1, 2) - T(2, 1)) / Q.d[0];
      break;
    case 1:
      Q.d[1] = 0.50 * std::sqrt(tq[1]);
      Q.d[0] = 0.25 * (T(2, 3) - T(3, 2)) / Q.d[1];
      Q.d[2] = 0.25 * (T(1, 2) + T(2, 1)) / Q.d[1];
      Q.d[3] = 0.25 * (T(3, 1) + T(1, 3)) / Q.d[1];
      break;
    case 2:
      Q.d[2] = 0.50 * std::sqrt(tq[2]);
      Q.d[0] = 0.25 * (T(3, 1) - T(1, 3)) / Q.d[2];
      Q.d[1] = 0.25 * (T(1, 2) + T(2, 1)) / Q.d[2];
      Q.d[3] = 0.25 * (T(2, 3) + T(3, 2)) / Q.d[2];
      break;
    default:
      Q.d[3] = 0.50 * std::sqrt(tq[3]);
      Q.d[0] = 0.25 * (T(1, 2) - T(2, 1)) / Q.d[3];
      Q.d[1] = 0.25 * (T(1, 3) + T(3, 1)) / Q.d[3];
      Q.d[2] = 0.25 * (T(2, 3) + T(3, 2)) / Q.d[3];
      break;
  }
  return Q;
}

// FGMatrix33::GetEuler -> (phi, theta, psi), psi in [0, 2pi)
V3 MatrixToEuler(const M33& T) {
  V3 e;
  bool GimbalLock = false;
  if (T(1, 3) <= -1.0) { e(2) = 0.5 * M_PI; GimbalLock = true; }
  else if (1.0 <= T(1, 3)) { e(2) = -0.5 * M_PI; GimbalLock = true; }
  else e(2) = std::asin(-T(1, 3));
  if (GimbalLock) e(1) = std::atan2(-T(3, 2), T(2, 2));
  else e(1) = std::atan2(T(2, 3), T(3, 3));
  if (GimbalLock) e(3) = 0.0;
  else {
    double psi = std::atan2(T(1, 2), T(1, 1));
    if (psi < 0.0) psi += 2 * M_PI;
    e(3) = psi;
  }
  return e;
}

// ------------------------------------------------------------------ FGFCSComponent / FGTable helpers
inline double constrain(double mn, double v, double mx) { return v < mn ? mn : (v > mx ? mx : v); }

// FGTable::GetValue(key) for a 1-D table; stateless breakpoint search (JSBSim's cached row index
// only matters when the key sits exactly on a breakpoint).
double table1d(const double* x, const double* y, int n, double key) {
  if (key <= x[0]) return y[0];
  if (key >= x[n - 1]) return y[n - 1];
  int r = 1;
  while (r < n - 1 && x[r] < key) r++;
  double Span = x[r] - x[r - 1];
  double Factor;
  if (Span != 0.0) {
    Factor = (key - x[r - 1]) / Span;
    if (Factor > 1.0) Factor = 1.0;
  } else {
    Factor = 1.0;
  }
  return Factor * (y[r] - y[r - 1]) + y[r - 1];
}

// FGTable::GetValue(rowKey, colKey): out-of-range keys clamp through the interpolation factors.
double table2d(const double* rk, const double* ck, const double* v, int nr, int nc, double rowKey, double colKey) {
  int r = 1, c = 1;
  while (r < nr - 1 && rk[r] < rowKey) r++;
  while (c < nc - 1 && ck[c] < colKey) c++;
  double rFactor = (rowKey - rk[r - 1]) / (rk[r] - rk[r - 1]);
  double cFactor = (colKey - ck[c - 1]) / (ck[c] - ck[c - 1]);
  if (rFactor > 1.0) rFactor = 1.0; else if (rFactor < 0.0) rFactor = 0.0;
  if (cFactor > 1.0) cFactor = 1.0; else if (cFactor < 0.0) cFactor = 0.0;
  double col1temp = rFactor * (v[r * nc + c - 1] - v[(r - 1) * nc + c - 1]) + v[(r - 1) * nc + c - 1];
  double col2temp = rFactor * (v[r * nc + c] - v[(r - 1) * nc + c]) + v[(r - 1) * nc + c];
  return col1temp + cFactor * (col2temp - col1temp);
}

// FGGain::Run, AEROSURFACE_SCALE branch, zero-centred (the default), gain 1
double aerosurface_scale(double Input, double InMin, double InMax, double OutMin, double OutMax) {
  double Output;
  if (Input == 0.0) Output = 0.0;
  else if (Input > 0) Output = (Input / InMax) * OutMax;
  else Output = (Input / InMin) * OutMin;
  Output *= 1.0;
  return Output;
}

inline bool EqualToRoundoff(double a, double b) {
  double eps = 2.0 * DBL_EPSILON;
  return std::fabs(a - b) <= eps * std::max(std::fabs(a), std::fabs(b));
}

// FGKinematic::Run. `Output` is the component's current output (re-read from its <output>
// property when it has one), scale-by-last-detent is on (no <noscale/> in f16.xml).
double kinematic_run(const double* Detents, const double* TransitionTimes, int n, double Input, double Output, double dt) {
  double dt0 = dt;
  Input *= Detents[n - 1];
  Input = constrain(Detents[0], Input, Detents[n - 1]);
  while (dt0 > 0.0 && !EqualToRoundoff(Input, Output)) {
    int ind;
    for (ind = 1; (Input < Output) ? Detents[ind] < Output : Detents[ind] <= Output; ++ind)
      if (ind >= n - 1) break;
    if (ind > n - 1) ind = n - 1;
    if (TransitionTimes[ind] <= 0.0) {
      Output = Input;
      break;
    } else {
      double Rate = (Detents[ind] - Detents[ind - 1]) / TransitionTimes[ind];
      double ThisInput = constrain(Detents[ind - 1], Input, Detents[ind]);
      double ThisDt = std::fabs((ThisInput - Output) / Rate);
      if (dt0 < ThisDt) {
        ThisDt = dt0;
        if (Output < Input) Output += ThisDt * Rate;
        else Output -= ThisDt * Rate;
      } else {
        Output = ThisInput;
      }
      dt0 -= ThisDt;
    }
  }
  return Output;
}

struct PidMem { double Input_prev = 0.0, Input_prev2 = 0.0, I_out_total = 0.0; };

// FGPID::Run, non-"standard" form, default Adams-Bashforth-2 integrator. The integrator only
// accumulates while |trigger| < 1e-6; a negative trigger resets it.
double pid_run(PidMem& m, double Input, double test, double Kp, double Ki, double Kd, double dt) {
  double I_out_delta = 0.0;
  double Dval = (Input - m.Input_prev) / dt;
  if (std::fabs(test) < 0.000001) I_out_delta = 1.5 * Input - 0.5 * m.Input_prev;
  if (test < 0.0) m.I_out_total = 0.0;
  m.I_out_total += Ki * dt * I_out_delta;
  double Output = Kp * Input + m.I_out_total + Kd * Dval;
  m.Input_prev2 = test < 0.0 ? 0.0 : m.Input_prev;
  m.Input_prev = Input;
  return Output;
}

#include "f16_oracle_gen.inc"

// ------------------------------------------------------------------ FGLocation subset (WGS84 ellipse set)
struct Earth {
  double a = 20925646.32546;   // WGS84 semimajor axis, ft
  double b = 20855486.5951;    // WGS84 semiminor axis, ft
  double GM = 14.0764417572E15;
  double J2 = 1.08262982E-03;
  double RotationRate = 0.00007292115;
  double ec, ec2, e2, c;
  Earth() { ec = b / a; ec2 = ec * ec; e2 = 1.0 - ec2; c = a * e2; }
};
const Earth kEarth;

struct Location {
  V3 mECLoc;
  double mRadius = 0, mLon = 0, mLat = 0, mGeodLat = 0, GeodeticAltitude = 0;
  M33 mTec2l, mTl2ec;
  // FGLocation::ComputeDerivedUnconditional (v1.1+: Fukushima 2006 geodetic latitude)
  void ComputeDerived() {
    const Earth& E = kEarth;
    mRadius = mECLoc.Magnitude();
    double rxy = mECLoc.Magnitude(1, 2);
    double sinLon, cosLon;
    if (rxy == 0.0) { sinLon = 0.0; cosLon = 1.0; mLon = 0.0; }
    else { sinLon = mECLoc(2) / rxy; cosLon = mECLoc(1) / rxy; mLon = std::atan2(mECLoc(2), mECLoc(1)); }
    if (mRadius == 0.0) mLat = 0.0;
    else mLat = std::atan2(mECLoc(3), rxy);
    double s0 = std::fabs(mECLoc(3));
    double zc = E.ec * s0;
    double c0 = E.ec * rxy;
    double c02 = c0 * c0;
    double s02 = s0 * s0;
    double a02 = c02 + s02;
    double a0 = std::sqrt(a02);
    double a03 = a02 * a0;
    double s1 = zc * a03 + E.c * s02 * s0;
    double c1 = rxy * a03 - E.c * c02 * c0;
    double cs0c0 = E.c * c0 * s0;
    double b0 = 1.5 * cs0c0 * ((rxy * s0 - zc * c0) * a0 - cs0c0);
    s1 = s1 * a03 - b0 * s0;
    double cc = E.ec * (c1 * a03 - b0 * c0);
    mGeodLat = (mECLoc(3) >= 0.0 ? 1.0 : -1.0) * std::atan(s1 / cc);
    double s12 = s1 * s1;
    double cc2 = cc * cc;
    GeodeticAltitude = (rxy * cc + s0 * s1 - E.a * std::sqrt(E.ec2 * s12 + cc2)) / std::sqrt(s12 + cc2);
    double cosLat = std::cos(mGeodLat), sinLat = std::sin(mGeodLat);
    mTec2l = M33(-cosLon * sinLat, -sinLon * sinLat, cosLat,
                 -sinLon, cosLon, 0.0,
                 -cosLon * cosLat, -sinLon * cosLat, -sinLat);
    mTl2ec = mTec2l.Transposed();
  }
  double GetSeaLevelRadius() const {
    double cosLat = std::cos(mLat);
    return kEarth.a * kEarth.ec / std::sqrt(1.0 - kEarth.e2 * cosLat * cosLat);
  }
  void SetPositionGeodetic(double lon, double lat, double height) {
    double slat = std::sin(lat), clat = std::cos(lat);
    double RN = kEarth.a / std::sqrt(1.0 - kEarth.e2 * slat * slat);
    mECLoc(1) = (RN + height) * clat * std::cos(lon);
    mECLoc(2) = (RN + height) * clat * std::sin(lon);
    mECLoc(3) = ((1 - kEarth.e2) * RN + height) * slat;
    ComputeDerived();
  }
  void SetRadius(double radius) {
    double rold = mECLoc.Magnitude();
    if (rold == 0.0) mECLoc(1) = radius;
    else mECLoc = mECLoc * (radius / rold);   // FGColumnVector3::operator*=(double)
    ComputeDerived();
  }
};

// ------------------------------------------------------------------ FGStandardAtmosphere (1976)
struct Atmosphere {
  static constexpr double Rstar = 8.31432 * kgtoslug / (1.8 * (fttom * fttom));
  static constexpr double Mair = 28.9645 * kgtoslug / 1000.0;
  static constexpr double g0 = 9.80665 / fttom;
  static constexpr double Reng0 = Rstar / Mair;
  static constexpr double SHRatio = 1.4;
  static constexpr double StdDaySLtemperature = 518.67;
  static constexpr double StdDaySLpressure = 2116.228;
  static constexpr double EarthRadius = 6356766.0 / fttom;
  static constexpr int NR = 9;
  double H[NR] = {0.0000, 36089.2388, 65616.7979, 104986.8766, 154199.4751, 167322.8346, 232939.6325, 278385.8268, 298556.4304};
  double T[NR] = {518.67, 389.97, 389.97, 411.57, 487.17, 487.17, 386.37, 336.5028, 336.5028};
  double LapseRates[NR - 1];
  double PressureBreakpoints[NR];
  double StdDensityBreakpoints[NR];
  double StdDaySLsoundspeed, SLdensity;
  // outputs
  double Temperature = 0, Pressure = 0, Density = 0, Soundspeed = 0, DensityAltitude = 0;
  struct { double altitudeASL = 0; } in;

  Atmosphere() {
    for (int b = 0; b < NR - 1; b++) LapseRates[b] = (T[b + 1] - T[b]) / (H[b + 1] - H[b]) - 0.0;
    PressureBreakpoints[0] = StdDaySLpressure;
    for (int b = 0; b < NR - 1; b++) {
      double BaseTemp = T[b], BaseAlt = H[b], UpperAlt = H[b + 1];
      double deltaH = UpperAlt - BaseAlt;
      double Tmb = BaseTemp;
      if (LapseRates[b] != 0.00) {
        double Lmb = LapseRates[b];
        double Exp = g0 / (Reng0 * Lmb);
        double factor = Tmb / (Tmb + Lmb * deltaH);
        PressureBreakpoints[b + 1] = PressureBreakpoints[b] * std::pow(factor, Exp);
      } else {
        PressureBreakpoints[b + 1] = PressureBreakpoints[b] * std::exp(-g0 * deltaH / (Reng0 * Tmb));
      }
    }
    for (int i = 0; i < NR; i++) StdDensityBreakpoints[i] = PressureBreakpoints[i] / (Reng0 * T[i]);
    StdDaySLsoundspeed = std::sqrt(SHRatio * Reng0 * StdDaySLtemperature);
    SLdensity = StdDaySLpressure / (Reng0 * StdDaySLtemperature);
  }
  static double GeopotentialAltitude(double geometalt) { return (geometalt * EarthRadius) / (EarthRadius + geometalt); }
  static double GeometricAltitude(double geopotalt) { return (geopotalt * EarthRadius) / (EarthRadius - geopotalt); }
  double GetTemperature(double altitude) const {
    double GeoPotAlt = GeopotentialAltitude(altitude);
    double Tk;
    if (GeoPotAlt >= 0.0) Tk = table1d(H, T, NR, GeoPotAlt);
    else Tk = T[0] + GeoPotAlt * LapseRates[0];
    return Tk;
  }
  double GetPressure(double altitude) const {
    double GeoPotAlt = GeopotentialAltitude(altitude);
    double BaseAlt = H[0];
    int b;
    for (b = 0; b < NR - 2; ++b) {
      double testAlt = H[b + 1];
      if (GeoPotAlt < testAlt) break;
      BaseAlt = testAlt;
    }
    double Tmb = GetTemperature(GeometricAltitude(BaseAlt));
    double deltaH = GeoPotAlt - BaseAlt;
    double Lmb = LapseRates[b];
    if (Lmb != 0.0) {
      double Exp = g0 / (Reng0 * Lmb);
      double factor = Tmb / (Tmb + Lmb * deltaH);
      return PressureBreakpoints[b] * std::pow(factor, Exp);
    }
    return PressureBreakpoints[b] * std::exp(-g0 * deltaH / (Reng0 * Tmb));
  }
  double CalculateDensityAltitude(double density) const {
    int b = 0;
    for (; b < NR - 2; b++)
      if (density >= StdDensityBreakpoints[b + 1]) break;
    double Tmb = T[b], Hb = H[b], Lmb = LapseRates[b], pb = StdDensityBreakpoints[b];
    double density_altitude;
    if (Lmb != 0.0) {
      double Exp = -1.0 / (1.0 + g0 / (Reng0 * Lmb));
      density_altitude = Hb + (Tmb / Lmb) * (std::pow(density / pb, Exp) - 1);
    } else {
      double Factor = -Reng0 * Tmb / g0;
      density_altitude = Hb + Factor * std::log(density / pb);
    }
    return GeometricAltitude(density_altitude);
  }
  void Run() {
    double altitude = in.altitudeASL;
    Temperature = GetTemperature(altitude);
    Pressure = GetPressure(altitude);
    Density = Pressure / (Reng0 * Temperature);
    Soundspeed = std::sqrt(SHRatio * Reng0 * Temperature);
    DensityAltitude = CalculateDensityAltitude(Density);
  }
};

// ------------------------------------------------------------------ the executive and its models
enum Phase { tpOff, tpRun, tpSpinUp, tpStart, tpStall, tpSeize, tpTrim };

struct FDM {
  // ---- FGFDMExec
  double dT = 1.0 / 120.0, saved_dT = 1.0 / 120.0;
  long Frame = 0;
  // ---- FGInitialCondition (only what the reference sets: jsbsim_gym.py:169-170)
  double ic_u_fps = 0.0, ic_h_sl_ft = 0.0;

  // ---- FGPropagate
  struct {
    Location vLocation;
    V3 vUVW, vPQR, vPQRi, vInertialVelocity, vInertialPosition;
    Quat qAttitudeLocal, qAttitudeECI, vQtrndot;
    V3 dqPQRidot[5], dqUVWidot[5], dqInertialVelocity[5];
    Quat dqQtrndot[5];
  } VState;
  double epa = 0.0;
  M33 Ti2ec, Tec2i, Tl2ec, Tec2l, Ti2l, Tl2i, Ti2b, Tb2i, Tl2b, Tb2l, Tec2b, Tb2ec;
  V3 vVel, vEuler;   // NED velocity; (phi, theta, psi) of qAttitudeLocal
  struct { V3 vPQRidot, vUVWidot; double DeltaT = 0; V3 vOmegaPlanet; } prop_in;

  // ---- FGInertial
  V3 vGravAccel;
  // ---- FGAtmosphere
  Atmosphere atm;
  // ---- FGFCS (Props doubles as the property tree for everything the F-16 files reference)
  Props P;
  FcsMem M;
  double ThrottleCmd = 0.0, ThrottlePos = 0.0;
  // ---- FGMassBalance
  V3 vXYZcg, vLastXYZcg;
  double Weight = 0, Mass = 0;
  M33 mJ, mJinv, baseJ;
  struct { double TanksWeight = 0; V3 TanksMoment; M33 TankInertia; } mb_in;
  // ---- FGPropulsion / FGTank / FGTurbine
  double tank_contents[4];
  struct {
    double ThrottlePos = 0, TotalDeltaT = 1.0 / 120.0, DensityRatio = 1.0;
  } pp_in;
  Phase phase = tpOff;
  bool Running = false, Cutoff = true, Starved = false, Augmentation = false;
  double N1 = 0, N2 = 0, N2norm = 0, AugmentCmd = 0, turb_ThrottlePos = 0, Thrust = 0;
  V3 prop_vForces, prop_vMoments;
  // ---- FGAuxiliary
  struct {
    double Pressure = 0, Density = 0, Temperature = 0, SoundSpeed = 0, DistanceAGL = 0;
    M33 Tl2b, Tb2l;
    V3 vPQR, vPQRi, vPQRidot, vUVW, vUVWdot, vVel, vBodyAccel, ToEyePt, RPBody;
  } aux_in;
  double alpha = 0, beta = 0, Vt = 0, qbar = 0, Mach = 0, Vground = 0, vcas = 0, pt = 0, hoverbmac = 0;
  V3 vAeroPQR, vAeroUVW, vPilotAccelN;
  M33 mTw2b, mTb2w;
  // ---- FGAerodynamics
  struct { double Alpha = 0, Beta = 0, Qbar = 0, Vt = 0; M33 Tw2b; V3 RPBody; } aero_in;
  double bi2vel = 0, ci2vel = 0;
  V3 aero_vFw, aero_vForces, aero_vMoments;
  double aero_fval[64];
  // ---- FGGroundReactions / FGLGear (contacts of f16.xml:85-215)
  struct LagrangeMultiplier { V3 ForceJacobian, LeverArm; double Min = 0, Max = 0, value = 0; };
  enum { ftRoll = 0, ftSide = 1, ftDynamic = 2 };
  struct Contact {
    bool WOW = false, lastWOW = false, StaticFriction = false;
    double compressLength = 0, compressSpeed = 0, StrutForce = 0;
    V3 vFn, vGroundNormal, vGroundWhlVel, vWhlVelVec, vActingXYZn;
    M33 mT;
    LagrangeMultiplier LMultiplier[3];
  } gear[modelk::n_contacts];
  struct {
    double TotalDeltaT = 0;
    M33 Tb2l, Tec2l, Tec2b;
    V3 PQR, UVW, vXYZcg;
    Location location;
  } gr_in;
  V3 gr_vForces, gr_vMoments;
  LagrangeMultiplier* multipliers[3 * modelk::n_contacts];
  int n_multipliers = 0;
  // ---- FGAircraft
  V3 ac_vForces, ac_vMoments;
  // ---- FGAccelerations
  V3 vPQRidot, vPQRdot, vUVWidot, vUVWdot, vBodyAccel, vFrictionForces, vFrictionMoments;

  FDM() {
    for (int i = 0; i < 4; i++) tank_contents[i] = modelk::tank_contents0[i];
    // FGFCS ctor: gear defaults down
    P.gear_gear_cmd_norm = 1.0;
    P.gear_gear_pos_norm = 1.0;
    P.metrics_Sw_sqft = modelk::Sw;
    P.metrics_bw_ft = modelk::bw;
    P.metrics_cbarw_ft = modelk::cbar;
    prop_in.vOmegaPlanet = V3(0, 0, kEarth.RotationRate);
    // FGMassBalance::ReadInertiaMatrix
    double bixx = modelk::ixx, biyy = modelk::iyy, bizz = modelk::izz, bixy = modelk::ixy, bixz = modelk::ixz, biyz = modelk::iyz;
    if (!modelk::negated_crossproduct_inertia)
      baseJ = M33(bixx, bixy, -bixz, bixy, biyy, biyz, -bixz, biyz, bizz);
    else
      baseJ = M33(bixx, -bixy, bixz, -bixy, biyy, -biyz, bixz, -biyz, bizz);
    Ti2ec = Tec2i = M33(1, 0, 0, 0, 1, 0, 0, 0, 1);
    mTw2b = mTb2w = M33(1, 0, 0, 0, 1, 0, 0, 0, 1);
  }

  // ================================================================ FGMassBalance helpers
  V3 StructuralToBody(const V3& r) const {
    return V3(inchtoft * (vXYZcg(1) - r(1)), inchtoft * (r(2) - vXYZcg(2)), inchtoft * (vXYZcg(3) - r(3)));
  }
  M33 GetPointmassInertia(double mass_sl, const V3& r) const {
    V3 v = StructuralToBody(r);
    V3 sv = mass_sl * v;
    double xx = sv(1) * v(1), yy = sv(2) * v(2), zz = sv(3) * v(3);
    double xy = -sv(1) * v(2), xz = -sv(1) * v(3), yz = -sv(2) * v(3);
    return M33(yy + zz, xy, xz, xy, xx + zz, yz, xz, yz, xx + yy);
  }

  // ================================================================ FGFDMExec::LoadInputs
  void load_propagate() {
    prop_in.vPQRidot = vPQRidot;
    prop_in.vUVWidot = vUVWidot;
    prop_in.DeltaT = dT;
  }
  double GetAltitudeASL() const { return VState.vLocation.mRadius - VState.vLocation.GetSeaLevelRadius(); }
  double GetDistanceAGL() const { return VState.vLocation.GeodeticAltitude - 0.0; }  // FGDefaultGroundCallback, terrain elevation 0

  // ================================================================ FGPropagate
  void UpdateLocationMatrices() {
    Tl2ec = VState.vLocation.mTl2ec;
    Tec2l = Tl2ec.Transposed();
    Ti2l = Tec2l * Ti2ec;
    Tl2i = Ti2l.Transposed();
  }
  void UpdateBodyMatrices() {
    Ti2b = VState.qAttitudeECI.GetT();
    Tb2i = Ti2b.Transposed();
    Tl2b = Ti2b * Tl2i;
    Tb2l = Tl2b.Transposed();
    Tec2b = Ti2b * Tec2i;
    Tb2ec = Tec2b.Transposed();
  }
  void SetLocalAttitude(const Quat& q) {
    VState.qAttitudeLocal = q;
    vEuler = MatrixToEuler(q.GetT());   // FGQuaternion derives its Euler angles from its own T
  }
  void SetInitialState() {
    // FGInitialCondition defaults: geocentric lat = lon = 0, identity local attitude, zero rates, no wind
    Location pos;
    pos.SetPositionGeodetic(0.0, 0.0, 0.0);
    double slr = pos.GetSeaLevelRadius();          // SetAltitudeASLFtIC, lastLatitudeSet == setgeoc
    pos.SetRadius(slr + ic_h_sl_ft);
    VState.vLocation = pos;
    epa = 0.0;
    Ti2ec = M33(std::cos(epa), std::sin(epa), 0.0, -std::sin(epa), std::cos(epa), 0.0, 0.0, 0.0, 1.0);
    Tec2i = Ti2ec.Transposed();
    VState.vInertialPosition = Tec2i * VState.vLocation.mECLoc;
    UpdateLocationMatrices();
    SetLocalAttitude(Quat(1, 0, 0, 0));
    VState.qAttitudeECI = MatrixToQuat(Ti2l) * VState.qAttitudeLocal;
    UpdateBodyMatrices();
    VState.vUVW = V3(ic_u_fps, 0, 0);
    vVel = Tb2l * VState.vUVW;
    VState.vPQR = V3(0, 0, 0);
    VState.vPQRi = VState.vPQR + Ti2b * prop_in.vOmegaPlanet;
    VState.vInertialVelocity = Tb2i * VState.vUVW + (prop_in.vOmegaPlanet * VState.vInertialPosition);
    VState.vQtrndot = VState.qAttitudeECI.GetQDot(VState.vPQRi);
  }
  template <class T>
  static void push(T* dq, const T& v) { for (int i = 4; i > 0; --i) dq[i] = dq[i - 1]; dq[0] = v; }
  void run_propagate() {
    double dt = prop_in.DeltaT;
    // attitude: rectangular Euler on the stored quaternion derivative, then normalise
    push(VState.dqQtrndot, VState.vQtrndot);
    for (int i = 0; i < 4; i++) VState.qAttitudeECI.d[i] += dt * VState.dqQtrndot[0].d[i];
    VState.qAttitudeECI.Normalize();
    // angular rate: rectangular Euler
    push(VState.dqPQRidot, prop_in.vPQRidot);
    VState.vPQRi += dt * VState.dqPQRidot[0];
    // position: Adams-Bashforth 3 on the inertial velocity history (velocity before this frame's update)
    push(VState.dqInertialVelocity, VState.vInertialVelocity);
    VState.vInertialPosition += (1 / 12.0) * dt * (23.0 * VState.dqInertialVelocity[0] - 16.0 * VState.dqInertialVelocity[1] + 5.0 * VState.dqInertialVelocity[2]);
    // velocity: Adams-Bashforth 2
    push(VState.dqUVWidot, prop_in.vUVWidot);
    VState.vInertialVelocity += dt * (1.5 * VState.dqUVWidot[0] - 0.5 * VState.dqUVWidot[1]);

    epa += prop_in.vOmegaPlanet(3) * dt;
    double cos_epa = std::cos(epa), sin_epa = std::sin(epa);
    Ti2ec = M33(cos_epa, sin_epa, 0.0, -sin_epa, cos_epa, 0.0, 0.0, 0.0, 1.0);
    Tec2i = Ti2ec.Transposed();
    VState.vLocation.mECLoc = Ti2ec * VState.vInertialPosition;
    VState.vLocation.ComputeDerived();
    UpdateLocationMatrices();
    UpdateBodyMatrices();
    VState.vUVW = Ti2b * (VState.vInertialVelocity - (prop_in.vOmegaPlanet * VState.vInertialPosition));
    VState.vPQR = VState.vPQRi - Ti2b * prop_in.vOmegaPlanet;
    VState.vQtrndot = VState.qAttitudeECI.GetQDot(VState.vPQRi);
    SetLocalAttitude(MatrixToQuat(Tl2b));
    vVel = Tb2l * VState.vUVW;
  }
  void InitializeDerivatives() {
    for (int i = 0; i < 5; i++) {
      VState.dqPQRidot[i] = prop_in.vPQRidot;     // NB: Propagate's `in` copy = what was loaded before the last Run()
      VState.dqUVWidot[i] = prop_in.vUVWidot;
      VState.dqInertialVelocity[i] = VState.vInertialVelocity;
      VState.dqQtrndot[i] = VState.vQtrndot;
    }
  }

  // ================================================================ FGInertial (gtWGS84, J2)
  void run_inertial() {
    const V3& pos = VState.vLocation.mECLoc;
    double r = VState.vLocation.mRadius;
    double sinLat = std::sin(VState.vLocation.mLat);
    double adivr = kEarth.a / r;
    double preCommon = 1.5 * kEarth.J2 * adivr * adivr;
    double xy = 1.0 - 5.0 * (sinLat * sinLat);
    double z = 3.0 - 5.0 * (sinLat * sinLat);
    double GMOverr2 = kEarth.GM / (r * r);
    vGravAccel(1) = -GMOverr2 * ((1.0 + (preCommon * xy)) * pos(1) / r);
    vGravAccel(2) = -GMOverr2 * ((1.0 + (preCommon * xy)) * pos(2) / r);
    vGravAccel(3) = -GMOverr2 * ((1.0 + (preCommon * z)) * pos(3) / r);
  }

  // ================================================================ FGFCS
  void run_fcs() {
    // properties owned by other models, as the property tree would serve them right now:
    // Propagate-owned ones are fresh (this frame), Auxiliary-owned ones still hold last frame's values.
    P.attitude_pitch_rad = vEuler(2);
    P.attitude_roll_rad = vEuler(1);
    P.velocities_u_fps = VState.vUVW(1);
    P.velocities_v_fps = VState.vUVW(2);
    P.velocities_vc_kts = vcas * fpstokts;
    P.velocities_mach = Mach;
    P.velocities_vg_fps = Vground;
    P.velocities_p_aero_rad_sec = vAeroPQR(1);
    P.velocities_q_aero_rad_sec = vAeroPQR(2);
    P.velocities_r_aero_rad_sec = vAeroPQR(3);
    P.aero_alpha_rad = alpha;
    P.aero_alpha_deg = alpha * radtodeg;
    P.accelerations_n_pilot_y_norm = vPilotAccelN(2);
    P.accelerations_n_pilot_z_norm = vPilotAccelN(3);
    ThrottlePos = ThrottleCmd = P.fcs_throttle_cmd_norm;   // FGFCS::Run: ThrottlePos[i] = ThrottleCmd[i]
    // system channels (Systems/pushback.xml, Systems/hook.xml) are inert in flight: their outputs
    // (external_reactions/pushback/magnitude, systems/hook/force) stay 0, see SURVEY.md B.1.
    fcs_channels_run(P, M, 1.0 / 120.0);   // component dt is latched at load time (FGFCSComponent ctor)
    ThrottlePos = P.fcs_throttle_pos_norm;
    P.fcs_speedbrake_pos_rad = P.fcs_speedbrake_pos_deg * degtorad;   // FGFCS::SetDsbPos(ofDeg)
  }

  // ================================================================ FGMassBalance
  void load_massbalance() {
    mb_in.TanksWeight = 0.0;
    mb_in.TanksMoment = V3();
    mb_in.TankInertia = M33();
    for (int i = 0; i < modelk::n_tanks; i++) {
      V3 loc(modelk::tank_loc[i][0], modelk::tank_loc[i][1], modelk::tank_loc[i][2]);
      mb_in.TanksWeight += tank_contents[i];
      mb_in.TanksMoment += loc * tank_contents[i];
      mb_in.TankInertia += GetPointmassInertia(lbtoslug * tank_contents[i], loc);   // uses the CG of the previous frame
    }
  }
  void run_massbalance() {
    double pmw = 0.0;
    V3 pmm;
    for (int i = 0; i < modelk::n_pointmass; i++) {
      pmw += modelk::pointmass_w[i];
      pmm += modelk::pointmass_w[i] * V3(modelk::pointmass_loc[i][0], modelk::pointmass_loc[i][1], modelk::pointmass_loc[i][2]);
    }
    V3 vbaseXYZcg(modelk::base_cg[0], modelk::base_cg[1], modelk::base_cg[2]);
    Weight = modelk::emptywt + mb_in.TanksWeight + pmw;
    Mass = lbtoslug * Weight;
    vXYZcg = (modelk::emptywt * vbaseXYZcg + pmm + mb_in.TanksMoment) / Weight;
    if (vLastXYZcg.Magnitude() == 0.0) vLastXYZcg = vXYZcg;
    vLastXYZcg = vXYZcg;
    mJ = baseJ;
    mJ += GetPointmassInertia(lbtoslug * modelk::emptywt, vbaseXYZcg);
    M33 pmJ;
    for (int i = 0; i < modelk::n_pointmass; i++)
      pmJ += GetPointmassInertia(lbtoslug * modelk::pointmass_w[i], V3(modelk::pointmass_loc[i][0], modelk::pointmass_loc[i][1], modelk::pointmass_loc[i][2]));
    mJ += pmJ;
    mJ += mb_in.TankInertia;
    double Ixx = mJ(1, 1), Iyy = mJ(2, 2), Izz = mJ(3, 3), Ixy = -mJ(1, 2), Ixz = -mJ(1, 3), Iyz = -mJ(2, 3);
    double k1 = (Iyy * Izz - Iyz * Iyz);
    double k2 = (Iyz * Ixz + Ixy * Izz);
    double k3 = (Ixy * Iyz + Iyy * Ixz);
    double denom = 1.0 / (Ixx * k1 - Ixy * k2 - Ixz * k3);
    k1 = k1 * denom; k2 = k2 * denom; k3 = k3 * denom;
    double k4 = (Izz * Ixx - Ixz * Ixz) * denom;
    double k5 = (Ixy * Ixz + Iyz * Ixx) * denom;
    double k6 = (Ixx * Iyy - Ixy * Ixy) * denom;
    mJinv = M33(k1, k2, k3, k2, k4, k5, k3, k5, k6);
  }

  // ================================================================ FGAuxiliary
  void load_auxiliary() {
    aux_in.Pressure = atm.Pressure;
    aux_in.Density = atm.Density;
    aux_in.Temperature = atm.Temperature;
    aux_in.SoundSpeed = atm.Soundspeed;
    aux_in.DistanceAGL = GetDistanceAGL();
    aux_in.Tl2b = Tl2b;
    aux_in.Tb2l = Tb2l;
    aux_in.vPQR = VState.vPQR;
    aux_in.vPQRi = VState.vPQRi;
    aux_in.vPQRidot = vPQRidot;        // Accelerations of the previous frame
    aux_in.vUVW = VState.vUVW;
    aux_in.vUVWdot = vUVWdot;
    aux_in.vVel = vVel;
    aux_in.vBodyAccel = vBodyAccel;    // previous frame
    aux_in.ToEyePt = StructuralToBody(V3(modelk::EYEPOINT[0], modelk::EYEPOINT[1], modelk::EYEPOINT[2]));
    aux_in.RPBody = StructuralToBody(V3(modelk::AERORP[0], modelk::AERORP[1], modelk::AERORP[2]));
  }
  static double PitotTotalPressure(double mach, double p) {
    if (mach < 0) return p;
    if (mach < 1) return p * std::pow((1 + 0.2 * mach * mach), 3.5);
    return p * 166.92158009316827 * std::pow(mach, 7.0) / std::pow(7 * mach * mach - 1, 2.5);
  }
  static double MachFromImpactPressure(double qc, double p) {
    double A = qc / p + 1;
    double Mc = std::sqrt(5.0 * (std::pow(A, 1. / 3.5) - 1));
    if (Mc > 1.0)
      for (unsigned int i = 0; i < 10; i++) Mc = 0.8812848543473311 * std::sqrt(A * std::pow(1 - 1.0 / (7.0 * Mc * Mc), 2.5));
    return Mc;
  }
  void run_auxiliary() {
    const double StandardGravity = 9.80665 / fttom;
    vAeroPQR = aux_in.vPQR;                 // no turbulence
    vAeroUVW = aux_in.vUVW;                 // no wind
    alpha = beta = 0;
    double AeroU2 = vAeroUVW(1) * vAeroUVW(1);
    double AeroV2 = vAeroUVW(2) * vAeroUVW(2);
    double AeroW2 = vAeroUVW(3) * vAeroUVW(3);
    double mUW = AeroU2 + AeroW2;
    double Vt2 = mUW + AeroV2;
    Vt = std::sqrt(Vt2);
    if (Vt > 0.001) {
      beta = std::atan2(vAeroUVW(2), std::sqrt(mUW));
      if (mUW >= 1E-6) alpha = std::atan2(vAeroUVW(3), vAeroUVW(1));
    }
    double ca = std::cos(alpha), sa = std::sin(alpha), cb = std::cos(beta), sb = std::sin(beta);
    mTw2b = M33(ca * cb, -ca * sb, -sa, sb, cb, 0.0, sa * cb, -sa * sb, ca);
    mTb2w = mTw2b.Transposed();
    double densityD2 = 0.5 * aux_in.Density;
    qbar = densityD2 * Vt2;
    Mach = Vt / aux_in.SoundSpeed;
    Vground = std::sqrt(aux_in.vVel(1) * aux_in.vVel(1) + aux_in.vVel(2) * aux_in.vVel(2));
    pt = PitotTotalPressure(Mach, aux_in.Pressure);
    if (std::fabs(Mach) > 0.0) {
      double qc = PitotTotalPressure(Mach, aux_in.Pressure) - aux_in.Pressure;   // VcalibratedFromMach
      vcas = atm.StdDaySLsoundspeed * MachFromImpactPressure(qc, Atmosphere::StdDaySLpressure);
    } else {
      vcas = 0.0;
    }
    V3 vPilotAccel = aux_in.vBodyAccel + aux_in.vPQRidot * aux_in.ToEyePt;
    vPilotAccel += aux_in.vPQRi * (aux_in.vPQRi * aux_in.ToEyePt);
    vPilotAccelN = vPilotAccel / StandardGravity;
    V3 vMac = aux_in.Tb2l * aux_in.RPBody;
    hoverbmac = (aux_in.DistanceAGL - vMac(3)) / modelk::bw;
  }

  // ================================================================ FGPropulsion / FGTurbine
  void load_propulsion() {
    pp_in.DensityRatio = atm.Density / atm.SLdensity;
    pp_in.TotalDeltaT = dT;
    pp_in.ThrottlePos = ThrottlePos;
  }
  double eng_table(const double* rk, int nr, const double* v) const {
    return table2d(rk, T_ENG_IdleThrust_c, v, nr, 8, Mach, atm.DensityAltitude);
  }
  double IdleThrustLookup() const { return eng_table(T_ENG_IdleThrust_r, 6, &T_ENG_IdleThrust_v[0][0]); }
  double MilThrustLookup() const { return eng_table(T_ENG_MilThrust_r, 8, &T_ENG_MilThrust_v[0][0]); }
  double MaxThrustLookup() const { return eng_table(T_ENG_AugThrust_r, 14, &T_ENG_AugThrust_v[0][0]); }
  double Seek(double var, double target, double accel, double decel) const {
    double v = var;
    if (v > target) { v -= pp_in.TotalDeltaT * decel; if (v < target) v = target; }
    else if (v < target) { v += pp_in.TotalDeltaT * accel; if (v > target) v = target; }
    return v;
  }
  double SpoolUp(double factor) const {   // FGSpoolUp::GetValue
    double delay = factor * 90.0 / (modelk::bypassratio + 3.0);
    double n = std::min(1.0, N2norm + 0.1);
    return delay / (1 + 3 * (1 - n) * (1 - n) * (1 - n) + (1 - pp_in.DensityRatio));
  }
  double turbine_off() {
    Running = false;
    N1 = Seek(N1, qbar / 10.0, N1 / 2.0, N1 / (modelk::maxn1 - modelk::idlen1));
    N2 = Seek(N2, qbar / 15.0, N2 / 2.0, N2 / (modelk::maxn2 - modelk::idlen2));
    Augmentation = false;
    return 0.0;
  }
  double turbine_trim() {
    const double N2_factor = modelk::maxn2 - modelk::idlen2;
    double idlethrust = modelk::milthrust * IdleThrustLookup();
    double milthrust = (modelk::milthrust - idlethrust) * MilThrustLookup();
    double N2l = modelk::idlen2 + turb_ThrottlePos * N2_factor;
    double N2n = (N2l - modelk::idlen2) / N2_factor;
    double thrust = (idlethrust + (milthrust * N2n * N2n)) * (1.0 - modelk::bleed);
    if (AugmentCmd > 0.0) {   // AugMethod == 2
      double tdiff = (modelk::maxthrust * MaxThrustLookup()) - thrust;
      thrust += (tdiff * std::min(AugmentCmd, 1.0));
    }
    return thrust;
  }
  double turbine_run() {
    const double N1_factor = modelk::maxn1 - modelk::idlen1, N2_factor = modelk::maxn2 - modelk::idlen2;
    double idlethrust = modelk::milthrust * IdleThrustLookup();
    double milthrust = (modelk::milthrust - idlethrust) * MilThrustLookup();
    Running = true;
    double n2up = SpoolUp(1.0), n2dn = SpoolUp(3.0), n1up = SpoolUp(1.0), n1dn = SpoolUp(2.4);
    N2 = Seek(N2, modelk::idlen2 + turb_ThrottlePos * N2_factor, n2up, n2dn);
    N1 = Seek(N1, modelk::idlen1 + turb_ThrottlePos * N1_factor, n1up, n1dn);
    N2norm = (N2 - modelk::idlen2) / N2_factor;
    double thrust = idlethrust + (milthrust * N2norm * N2norm);
    if (!Augmentation) thrust = thrust * (1.0 - modelk::bleed);
    if (AugmentCmd > 0.0) {   // AugMethod == 2
      Augmentation = true;
      double tdiff = (modelk::maxthrust * MaxThrustLookup()) - thrust;
      thrust += (tdiff * std::min(AugmentCmd, 1.0));
    } else {
      Augmentation = false;
    }
    return thrust;
  }
  void turbine_calculate() {
    turb_ThrottlePos = pp_in.ThrottlePos;
    if (turb_ThrottlePos > 1.0) { AugmentCmd = turb_ThrottlePos - 1.0; turb_ThrottlePos -= AugmentCmd; }
    else AugmentCmd = 0.0;
    if ((phase == tpTrim) && (pp_in.TotalDeltaT > 0)) {
      if (Running && !Starved) {
        phase = tpRun;
        N2 = modelk::idlen2 + turb_ThrottlePos * (modelk::maxn2 - modelk::idlen2);
        N1 = modelk::idlen1 + turb_ThrottlePos * (modelk::maxn1 - modelk::idlen1);
        Cutoff = false;
      } else {
        phase = tpOff;
        Cutoff = true;
      }
    }
    if (Cutoff && (phase != tpSpinUp)) phase = tpOff;
    if (pp_in.TotalDeltaT == 0) phase = tpTrim;
    if (Starved) phase = tpOff;
    double thrust;
    switch (phase) {
      case tpRun: thrust = turbine_run(); break;
      case tpTrim: thrust = turbine_trim(); break;
      default: thrust = turbine_off(); break;
    }
    Thrust = thrust;   // FGThruster::Calculate, "direct" thruster, no reverser
  }
  void engine_forces() {
    V3 vFb(Thrust, 0.0, 0.0);
    V3 vDXYZ = StructuralToBody(V3(modelk::thruster_loc[0], modelk::thruster_loc[1], modelk::thruster_loc[2]));
    prop_vForces = V3() + vFb;
    prop_vMoments = V3() + vDXYZ * vFb;
  }
  void run_propulsion() {
    turbine_calculate();
    engine_forces();
    // ConsumeFuel: the reference refills both internal tanks before every run() (jsbsim_gym.py:227-228)
    // and MassBalance reads the contents before Propulsion runs, so the burn never feeds back.
  }
  // FGPropulsion::InitRunning(-1) -> FGTurbine::InitRunning + GetSteadyState (property propulsion/set-running)
  void InitRunning() {
    pp_in.ThrottlePos = 1;
    Cutoff = false;
    Running = true;
    N2 = 16.0;
    turbine_calculate();
    double TimeStep = dT;
    pp_in.TotalDeltaT = 0.5;
    double currentThrust = 0, lastThrust = -1;
    int steady_count = 0, j = 0;
    bool steady = false;
    while (!steady && j < 6000) {
      turbine_calculate();
      lastThrust = currentThrust;
      currentThrust = Thrust;
      if (std::fabs(lastThrust - currentThrust) < 0.0001) {
        steady_count++;
        if (steady_count > 120) steady = true;
      } else {
        steady_count = 0;
      }
      j++;
    }
    engine_forces();
    pp_in.TotalDeltaT = TimeStep;
  }

  // ================================================================ FGAerodynamics
  void load_aerodynamics() {
    aero_in.Alpha = alpha;
    aero_in.Beta = beta;
    aero_in.Qbar = qbar;
    aero_in.Vt = Vt;
    aero_in.Tw2b = mTw2b;
    aero_in.RPBody = StructuralToBody(V3(modelk::AERORP[0], modelk::AERORP[1], modelk::AERORP[2]));
  }
  void run_aerodynamics() {
    const double twovel = 2 * aero_in.Vt;
    if (twovel != 0) { bi2vel = modelk::bw / twovel; ci2vel = modelk::cbar / twovel; }
    P.aero_bi2vel = bi2vel;
    P.aero_ci2vel = ci2vel;
    P.aero_qbar_psf = qbar;
    P.aero_alpha_rad = alpha;         // fresh now (Auxiliary already ran this frame)
    P.aero_beta_rad = beta;
    P.velocities_mach = Mach;
    P.velocities_p_aero_rad_sec = vAeroPQR(1);
    P.velocities_q_aero_rad_sec = vAeroPQR(2);
    P.velocities_r_aero_rad_sec = vAeroPQR(3);
    P.aero_h_b_mac_ft = hoverbmac;
    double axis[6];
    aero_functions_run(P, axis, aero_fval);
    V3 vFnative(axis[0], axis[1], axis[2]);
    aero_vFw = vFnative;              // atLiftDrag
    aero_vFw(1) *= -1;
    aero_vFw(3) *= -1;
    aero_vForces = aero_in.Tw2b * aero_vFw;
    V3 vMomentsMRC(axis[3], axis[4], axis[5]);
    aero_vMoments = vMomentsMRC + aero_in.RPBody * aero_vForces;   // M = r x F about the CG
  }

  // ================================================================ FGGroundReactions / FGLGear
  // [UPSTREAM] FGLGear::GetBodyForces and helpers for the contacts of f16.xml:85-215, terrain =
  // FGDefaultGroundCallback (WGS84 ellipsoid, elevation 0, at rest in ECEF). The three BOGEY contacts
  // are retractable and the reference forces gear/gear-pos-norm to 0 before every run()
  // (jsbsim_gym.py:230-231), so in flight only their "gear up" branch executes; a BOGEY that is
  // down AND compressed (tyre slip, brakes, steering) cannot occur on the reference's path (gear is
  // only down during run_ic at 5000 ft) and is reported instead of modelled.
  void load_groundreactions() {
    gr_in.TotalDeltaT = dT;
    gr_in.Tb2l = Tb2l;
    gr_in.Tec2l = Tec2l;
    gr_in.Tec2b = Tec2b;
    gr_in.PQR = VState.vPQR;
    gr_in.UVW = VState.vUVW;
    gr_in.vXYZcg = vXYZcg;
    gr_in.location = VState.vLocation;
  }
  static V3 mat_col_mul(const M33& m, const V3& a) { return m * a; }
  // FGLGear::GetBodyForces for contact i; returns the body force, *moment receives FGForce::GetMoments
  V3 gear_body_forces(int i, V3* moment) {
    const modelk::ContactDef& def = modelk::contacts[i];
    Contact& c = gear[i];
    const V3 vXYZn(def.loc[0], def.loc[1], def.loc[2]);
    double gearPos = 1.0;
    c.vFn = V3();
    if (def.retractable) gearPos = P.gear_gear_pos_norm;     // GetGearUnitPos
    c.vActingXYZn = vXYZn;
    if (gearPos > 0.99) {   // gear down (always, for a STRUCTURE contact)
      // Ts2b * (vXYZn - vXYZcg): structural inches -> body feet
      V3 d = vXYZn - gr_in.vXYZcg;
      V3 vWhlBodyVec(-inchtoft * d(1), inchtoft * d(2), -inchtoft * d(3));
      V3 vLocalGear = gr_in.Tb2l * vWhlBodyVec;
      // FGLocation::LocalToLocation + FGDefaultGroundCallback::GetAGLevel
      Location gearLoc;
      gearLoc.mECLoc = gr_in.location.mTl2ec * vLocalGear + gr_in.location.mECLoc;
      gearLoc.ComputeDerived();
      double cosLat = std::cos(gearLoc.mGeodLat);
      V3 normal(cosLat * std::cos(gearLoc.mLon), cosLat * std::sin(gearLoc.mLon), std::sin(gearLoc.mGeodLat));
      double height = gearLoc.GeodeticAltitude - 0.0;
      V3 vWhlDisplVec;
      if (height < 0.0) {
        c.WOW = true;
        c.vGroundNormal = gr_in.Tec2b * normal;
        double normalZ = (gr_in.Tec2l * normal)(3);
        if (def.bogey) {
          std::fprintf(stderr, "f16 oracle: BOGEY contact %s compressed - not on the reference's path\n", def.name);
          c.WOW = false;
        } else {
          double nn = normal(1) * normal(1) + normal(2) * normal(2) + normal(3) * normal(3);
          c.compressLength = height * normalZ / nn;
          vWhlDisplVec = c.compressLength * c.vGroundNormal;
        }
      } else {
        c.WOW = false;
      }
      if (c.WOW) {
        V3 vWhlContactVec = vWhlBodyVec + vWhlDisplVec;
        // vActingXYZn = vXYZn + Tb2s * vWhlDisplVec (body feet -> structural inches)
        c.vActingXYZn = vXYZn + V3(-12.0 * vWhlDisplVec(1), 12.0 * vWhlDisplVec(2), -12.0 * vWhlDisplVec(3));
        V3 vBodyWhlVel = gr_in.PQR * vWhlContactVec;
        vBodyWhlVel += gr_in.UVW;                            // terrain velocity is zero
        c.vWhlVelVec = vBodyWhlVel;                          // mTGear = identity (no strut angles in the file)
        // ComputeGroundFrame, SteerAngle = 0
        V3 roll(1.0, 0.0, 0.0);
        V3 side = c.vGroundNormal * roll;
        double rn = roll(1) * c.vGroundNormal(1) + roll(2) * c.vGroundNormal(2) + roll(3) * c.vGroundNormal(3);
        roll -= rn * c.vGroundNormal;
        { double m = roll.Magnitude(); if (m != 0.0) roll = roll / m; }      // FGColumnVector3::Normalize
        { double m = side.Magnitude(); if (m != 0.0) side = side / m; }
        c.mT = M33(roll(1), side(1), c.vGroundNormal(1),
                   roll(2), side(2), c.vGroundNormal(2),
                   roll(3), side(3), c.vGroundNormal(3));
        c.vGroundWhlVel = c.mT.Transposed() * vBodyWhlVel;
        c.compressSpeed = -c.vGroundWhlVel(3);               // ctSTRUCTURE: along the ground normal
        // ComputeVerticalStrutForce: linear spring, linear damping (same coefficient on rebound)
        double springForce = -c.compressLength * def.spring;
        double dampForce = -c.compressSpeed * def.damping;
        c.StrutForce = std::min(springForce + dampForce, 0.0);
        c.vFn(3) = -c.StrutForce;                            // ctSTRUCTURE: normal to the ground
        // ComputeJacobian
        const double staticFFactor = 1.0;                    // FGSurface defaults
        if (c.vGroundWhlVel.Magnitude(1, 2) > 1E-3) {
          V3 velocityDirection = c.vGroundWhlVel;
          c.StaticFriction = false;
          velocityDirection(3) = 0.0;
          { double m = velocityDirection.Magnitude(); if (m != 0.0) velocityDirection = velocityDirection / m; }
          LagrangeMultiplier& L = c.LMultiplier[ftDynamic];
          L.ForceJacobian = c.mT * velocityDirection;
          L.LeverArm = vWhlContactVec;
          L.Max = 0.0;
          L.Min = -std::fabs(staticFFactor * def.dynamic_f * c.vFn(3));
          L.value = constrain(L.Min, L.value, L.Max);
          multipliers[n_multipliers++] = &L;
        } else {
          c.StaticFriction = true;
          LagrangeMultiplier& Lr = c.LMultiplier[ftRoll];
          LagrangeMultiplier& Ls = c.LMultiplier[ftSide];
          Lr.ForceJacobian = c.mT * V3(1.0, 0.0, 0.0);
          Ls.ForceJacobian = c.mT * V3(0.0, 1.0, 0.0);
          Lr.LeverArm = vWhlContactVec;
          Ls.LeverArm = vWhlContactVec;
          Lr.Max = std::fabs(staticFFactor * def.static_f * c.vFn(3));
          Ls.Max = Lr.Max;
          Lr.Min = -Lr.Max;
          Ls.Min = -Ls.Max;
          Lr.value = constrain(Lr.Min, Lr.value, Lr.Max);
          Ls.value = constrain(Ls.Min, Ls.value, Ls.Max);
          multipliers[n_multipliers++] = &Lr;
          multipliers[n_multipliers++] = &Ls;
        }
      } else {   // not compressed
        c.compressLength = 0.0;
        c.compressSpeed = 0.0;
        c.StrutForce = 0.0;
        c.LMultiplier[ftRoll].value = 0.0;
        c.LMultiplier[ftSide].value = 0.0;
        c.LMultiplier[ftDynamic].value = 0.0;
        c.vWhlVelVec(1) -= 13.0 * gr_in.TotalDeltaT;         // wheel spin-down (no effect on forces)
        if (c.vWhlVelVec(1) < 0.0) c.vWhlVelVec(1) = 0.0;
      }
    } else if (gearPos < 0.01) {   // gear up
      c.WOW = false;
      c.vWhlVelVec = V3();
    }
    c.lastWOW = c.WOW;             // CrashDetect only reports
    // FGForce::GetBodyForces: vFb = mT * vFn; vM = StructuralToBody(vActingXYZn) x vFb
    V3 vFb = c.WOW ? c.mT * c.vFn : V3();
    *moment = StructuralToBody(c.vActingXYZn) * vFb;
    return vFb;
  }
  void run_groundreactions() {
    gr_vForces = V3();
    gr_vMoments = V3();
    n_multipliers = 0;
    for (int i = 0; i < modelk::n_contacts; i++) {
      V3 m;
      gr_vForces += gear_body_forces(i, &m);
      gr_vMoments += m;
    }
  }

  // ================================================================ FGAircraft / FGAccelerations
  void run_aircraft() {
    ac_vForces = V3();
    ac_vMoments = V3();
    ac_vForces += aero_vForces;
    ac_vForces += prop_vForces;
    ac_vForces += gr_vForces;        // external (pushback, hook: magnitude 0) and buoyant reactions are zero
    ac_vMoments += aero_vMoments;
    ac_vMoments += prop_vMoments;
    ac_vMoments += gr_vMoments;
  }
  // [UPSTREAM] FGAccelerations::CalculateFrictionForces: projected Gauss-Seidel on the Lagrange
  // multipliers registered by the contacts, <= 50 sweeps, stops when the summed change is < 1e-5
  void resolve_friction_forces(double dt) {
    const int n = n_multipliers;
    vFrictionForces = V3();
    vFrictionMoments = V3();
    if (!n) return;
    std::vector<double> a((size_t)n * n), rhs(n);
    auto dot = [](const V3& x, const V3& y) { return x(1) * y(1) + x(2) * y(2) + x(3) * y(3); };
    for (int i = 0; i < n; i++) {
      V3 U = multipliers[i]->ForceJacobian;
      V3 r = multipliers[i]->LeverArm;
      V3 v1 = U / Mass;
      V3 v2 = mJinv * (r * U);
      for (int j = 0; j < i; j++) a[i * n + j] = a[j * n + i];
      for (int j = i; j < n; j++) {
        U = multipliers[j]->ForceJacobian;
        r = multipliers[j]->LeverArm;
        a[i * n + j] = dot(U, v1 + v2 * r);
      }
    }
    V3 vdot = vUVWdot;
    if (dt > 0.) vdot += VState.vUVW / dt;          // terrain at rest
    V3 wdot = vPQRdot;
    if (dt > 0.) wdot += VState.vPQR / dt;
    for (int i = 0; i < n; i++) {
      double d = a[i * n + i];
      V3 U = multipliers[i]->ForceJacobian;
      V3 r = multipliers[i]->LeverArm;
      rhs[i] = -dot(U, vdot + wdot * r) / d;
      for (int j = 0; j < n; j++) a[i * n + j] /= d;
    }
    for (int iter = 0; iter < 50; iter++) {
      double norm = 0.;
      for (int i = 0; i < n; i++) {
        double lambda0 = multipliers[i]->value;
        double dlambda = rhs[i];
        for (int j = 0; j < n; j++) dlambda -= a[i * n + j] * multipliers[j]->value;
        multipliers[i]->value = constrain(multipliers[i]->Min, lambda0 + dlambda, multipliers[i]->Max);
        dlambda = multipliers[i]->value - lambda0;
        norm += std::fabs(dlambda);
      }
      if (norm < 1E-5) break;
    }
    for (int i = 0; i < n; i++) {
      double lambda = multipliers[i]->value;
      V3 F = lambda * multipliers[i]->ForceJacobian;
      vFrictionForces += F;
      vFrictionMoments += multipliers[i]->LeverArm * F;
    }
    V3 accel = vFrictionForces / Mass;
    V3 omegadot = mJinv * vFrictionMoments;
    vBodyAccel += accel;
    vUVWdot += accel;
    vUVWidot += Tb2i * accel;
    vPQRdot += omegadot;
    vPQRidot += omegadot;
  }
  void run_accelerations() {
    const V3& wp = prop_in.vOmegaPlanet;
    vPQRidot = mJinv * (ac_vMoments - VState.vPQRi * (mJ * VState.vPQRi));
    vPQRdot = vPQRidot - VState.vPQRi * (Ti2b * wp);
    vBodyAccel = ac_vForces / Mass;
    vUVWdot = vBodyAccel - (VState.vPQR + 2.0 * (Ti2b * wp)) * VState.vUVW;
    vUVWdot -= Ti2b * (wp * (wp * VState.vInertialPosition));
    vUVWdot += Tec2b * vGravAccel;
    vUVWidot = Tb2i * vBodyAccel + Tec2i * vGravAccel;
    resolve_friction_forces(dT);
  }

  // ================================================================ FGFDMExec::Run / RunIC
  bool Run() {
    if (dT != 0.0) Frame++;          // IncrementTime
    load_propagate();  run_propagate();
    run_inertial();
    atm.in.altitudeASL = GetAltitudeASL();  atm.Run();
    run_fcs();
    load_massbalance();  run_massbalance();
    load_auxiliary();  run_auxiliary();
    load_propulsion();  run_propulsion();
    load_aerodynamics();  run_aerodynamics();
    load_groundreactions();  run_groundreactions();
    run_aircraft();
    run_accelerations();
    return true;
  }
  bool RunIC() {
    saved_dT = dT;  dT = 0.0;        // SuspendIntegration
    SetInitialState();               // Initialize(IC): SetInitialState + Run
    Run();
    Run();
    InitializeDerivatives();
    dT = saved_dT;                   // ResumeIntegration
    return true;
  }

  // ================================================================ property tree (the 23 names the reference touches + extras for tests)
  bool set_property(const std::string& name, double value) {
    if (name == "propulsion/set-running") { InitRunning(); return true; }
    if (name == "ic/u-fps") { ic_u_fps = value; return true; }
    if (name == "ic/h-sl-ft") { ic_h_sl_ft = value; return true; }
    if (name == "propulsion/tank/contents-lbs" || name == "propulsion/tank[0]/contents-lbs") { tank_contents[0] = std::min(value, modelk::tank_capacity[0]); return true; }
    if (name == "propulsion/tank[1]/contents-lbs") { tank_contents[1] = std::min(value, modelk::tank_capacity[1]); return true; }
    for (int i = 0; i < kNumProps; i++)
      if (name == kPropNames[i]) { (&P.velocities_vc_kts)[i] = value; return true; }
    return false;
  }
  bool get_property(const std::string& name, double* out) const {
    const Location& L = VState.vLocation;
    if (name == "position/lat-gc-rad") { *out = L.mLat; return true; }
    if (name == "position/long-gc-rad") { *out = L.mLon; return true; }
    if (name == "position/h-sl-meters") { *out = GetAltitudeASL() * fttom; return true; }
    if (name == "position/h-sl-ft") { *out = GetAltitudeASL(); return true; }
    if (name == "position/h-agl-ft") { *out = GetDistanceAGL(); return true; }
    if (name == "position/lat-geod-rad") { *out = L.mGeodLat; return true; }
    if (name == "velocities/mach") { *out = Mach; return true; }
    if (name == "aero/alpha-rad") { *out = alpha; return true; }
    if (name == "aero/beta-rad") { *out = beta; return true; }
    if (name == "aero/qbar-psf") { *out = qbar; return true; }
    if (name == "velocities/p-rad_sec") { *out = VState.vPQR(1); return true; }
    if (name == "velocities/q-rad_sec") { *out = VState.vPQR(2); return true; }
    if (name == "velocities/r-rad_sec") { *out = VState.vPQR(3); return true; }
    if (name == "velocities/u-fps") { *out = VState.vUVW(1); return true; }
    if (name == "velocities/v-fps") { *out = VState.vUVW(2); return true; }
    if (name == "velocities/w-fps") { *out = VState.vUVW(3); return true; }
    if (name == "velocities/vt-fps") { *out = Vt; return true; }
    if (name == "velocities/vc-kts") { *out = vcas * fpstokts; return true; }
    if (name == "velocities/vg-fps") { *out = Vground; return true; }
    if (name == "attitude/phi-rad") { *out = vEuler(1); return true; }
    if (name == "attitude/theta-rad") { *out = vEuler(2); return true; }
    if (name == "attitude/psi-rad") { *out = vEuler(3); return true; }
    if (name == "atmosphere/T-R") { *out = atm.Temperature; return true; }
    if (name == "atmosphere/P-psf") { *out = atm.Pressure; return true; }
    if (name == "atmosphere/rho-slugs_ft3") { *out = atm.Density; return true; }
    if (name == "atmosphere/a-fps") { *out = atm.Soundspeed; return true; }
    if (name == "atmosphere/density-altitude") { *out = atm.DensityAltitude; return true; }
    if (name == "inertia/weight-lbs") { *out = Weight; return true; }
    if (name == "inertia/mass-slugs") { *out = Mass; return true; }
    if (name == "inertia/cg-x-in") { *out = vXYZcg(1); return true; }
    if (name == "inertia/cg-y-in") { *out = vXYZcg(2); return true; }
    if (name == "inertia/cg-z-in") { *out = vXYZcg(3); return true; }
    if (name == "inertia/ixx-slugs_ft2") { *out = mJ(1, 1); return true; }
    if (name == "inertia/iyy-slugs_ft2") { *out = mJ(2, 2); return true; }
    if (name == "inertia/izz-slugs_ft2") { *out = mJ(3, 3); return true; }
    if (name == "inertia/ixz-slugs_ft2") { *out = -mJ(1, 3); return true; }
    if (name == "propulsion/engine/n1") { *out = N1; return true; }
    if (name == "propulsion/engine/n2") { *out = N2; return true; }
    if (name == "propulsion/engine/thrust-lbs") { *out = Thrust; return true; }
    if (name == "propulsion/engine/augmentation") { *out = Augmentation ? 1.0 : 0.0; return true; }
    if (name == "accelerations/n-pilot-x-norm") { *out = vPilotAccelN(1); return true; }
    if (name == "accelerations/n-pilot-y-norm") { *out = vPilotAccelN(2); return true; }
    if (name == "accelerations/n-pilot-z-norm") { *out = vPilotAccelN(3); return true; }
    if (name == "accelerations/gravity-ft_sec2") { *out = vGravAccel.Magnitude(); return true; }
    if (name == "forces/fbx-aero-lbs") { *out = aero_vForces(1); return true; }
    if (name == "forces/fby-aero-lbs") { *out = aero_vForces(2); return true; }
    if (name == "forces/fbz-aero-lbs") { *out = aero_vForces(3); return true; }
    if (name == "forces/fbx-prop-lbs") { *out = prop_vForces(1); return true; }
    if (name == "moments/l-aero-lbsft") { *out = aero_vMoments(1); return true; }
    if (name == "moments/m-aero-lbsft") { *out = aero_vMoments(2); return true; }
    if (name == "moments/n-aero-lbsft") { *out = aero_vMoments(3); return true; }
    if (name == "moments/m-prop-lbsft") { *out = prop_vMoments(2); return true; }
    if (name == "forces/fbx-gear-lbs") { *out = gr_vForces(1); return true; }
    if (name == "forces/fby-gear-lbs") { *out = gr_vForces(2); return true; }
    if (name == "forces/fbz-gear-lbs") { *out = gr_vForces(3); return true; }
    if (name == "moments/l-gear-lbsft") { *out = gr_vMoments(1); return true; }
    if (name == "moments/m-gear-lbsft") { *out = gr_vMoments(2); return true; }
    if (name == "moments/n-gear-lbsft") { *out = gr_vMoments(3); return true; }
    if (name == "gear/num-friction-multipliers") { *out = (double)n_multipliers; return true; }   // not a JSBSim property
    if (name.rfind("gear/unit[", 0) == 0) {
      int i = std::atoi(name.c_str() + 10);
      if (i >= 0 && i < modelk::n_contacts) {
        std::string tail = name.substr(name.find(']') + 1);
        if (tail == "/WOW") { *out = gear[i].WOW ? 1.0 : 0.0; return true; }
        if (tail == "/compression-ft") { *out = gear[i].compressLength; return true; }
        if (tail == "/compression-velocity-fps") { *out = gear[i].compressSpeed; return true; }
      }
    }
    if (name == "simulation/frame") { *out = (double)Frame; return true; }
    if (name == "simulation/epa-rad") { *out = epa; return true; }
    if (name.rfind("aero/coefficient/", 0) == 0) {
      for (int i = 0; i < kNumAeroFunctions; i++)
        if (name == kAeroFunctionNames[i]) { *out = aero_fval[i]; return true; }
    }
    for (int i = 0; i < kNumProps; i++)
      if (name == kPropNames[i]) { *out = (&P.velocities_vc_kts)[i]; return true; }
    return false;
  }
};

// ------------------------------------------------------------------ packed state (shared field order with the CUDA library)
// Order is defined by include/f16_state_fields.h; the oracle fills it from its JSBSim-shaped members
// so tests can teacher-force the kernel from any oracle state and compare afterwards.
#include "../include/f16_state_fields.h"

void pack_state(const FDM& f, double* s) {
  for (int i = 0; i < F16_NUM_STATE_FIELDS; i++) s[i] = 0.0;
  const auto& V = f.VState;
  for (int i = 0; i < 4; i++) s[F16S_Q0 + i] = V.qAttitudeECI.d[i];
  for (int i = 0; i < 3; i++) {
    s[F16S_WI_X + i] = V.vPQRi.v[i];
    s[F16S_RI_X + i] = V.vInertialPosition.v[i];
    s[F16S_VI_X + i] = V.vInertialVelocity.v[i];
    s[F16S_VI1_X + i] = V.dqInertialVelocity[0].v[i];
    s[F16S_VI2_X + i] = V.dqInertialVelocity[1].v[i];
    s[F16S_AI0_X + i] = f.vUVWidot.v[i];
    s[F16S_AI1_X + i] = V.dqUVWidot[0].v[i];
    s[F16S_WDOT_X + i] = f.vPQRidot.v[i];
    s[F16S_ABODY_X + i] = f.vBodyAccel.v[i];
    s[F16S_PQR_X + i] = f.vAeroPQR.v[i];
  }
  s[F16S_EPA] = f.epa;
  s[F16S_ALPHA] = f.alpha;
  s[F16S_MACH] = f.Mach;
  s[F16S_VC_KTS] = f.vcas * fpstokts;
  s[F16S_VG] = f.Vground;
  s[F16S_NPY] = f.vPilotAccelN.v[1];
  s[F16S_NPZ] = f.vPilotAccelN.v[2];
  s[F16S_TEF] = f.P.fcs_tef_control;
  s[F16S_AIL] = f.P.fcs_left_aileron_pos_norm;
  s[F16S_ELEV] = f.P.fcs_elevator_pos_norm;
  s[F16S_SB_DEG] = f.P.fcs_speedbrake_pos_deg;
  s[F16S_ROLL_INPREV] = f.M.fcs_roll_rate_pid.Input_prev;
  s[F16S_ROLL_I] = f.M.fcs_roll_rate_pid.I_out_total;
  s[F16S_PITCH_INPREV] = f.M.fcs_g_load_pid.Input_prev;
  s[F16S_PITCH_I] = f.M.fcs_g_load_pid.I_out_total;
  s[F16S_YAW_INPREV] = f.M.fcs_yaw_load_pid.Input_prev;
  s[F16S_YAW_I] = f.M.fcs_yaw_load_pid.I_out_total;
  s[F16S_N2] = f.N2;
  s[F16S_AUG] = f.Augmentation ? 1.0 : 0.0;
}

// Inverse of pack_state, for tests that need the oracle in a synthetic state (ground contact at a chosen
// attitude and speed). The FDM must already be in flight configuration (reset + at least one env-step:
// engine running, gear up, tanks at 1000 lb, steady CG); only the packed fields are overwritten and the
// values Propagate derives from them (vQtrndot) are refreshed. Everything else is recomputed by the next Run().
void unpack_state(FDM& f, const double* s) {
  auto& V = f.VState;
  for (int i = 0; i < 4; i++) V.qAttitudeECI.d[i] = s[F16S_Q0 + i];
  for (int i = 0; i < 3; i++) {
    V.vPQRi.v[i] = s[F16S_WI_X + i];
    V.vInertialPosition.v[i] = s[F16S_RI_X + i];
    V.vInertialVelocity.v[i] = s[F16S_VI_X + i];
    V.dqInertialVelocity[0].v[i] = s[F16S_VI1_X + i];
    V.dqInertialVelocity[1].v[i] = s[F16S_VI2_X + i];
    f.vUVWidot.v[i] = s[F16S_AI0_X + i];
    V.dqUVWidot[0].v[i] = s[F16S_AI1_X + i];
    f.vPQRidot.v[i] = s[F16S_WDOT_X + i];
    f.vBodyAccel.v[i] = s[F16S_ABODY_X + i];
    f.vAeroPQR.v[i] = s[F16S_PQR_X + i];
  }
  f.epa = s[F16S_EPA];
  V.vQtrndot = V.qAttitudeECI.GetQDot(V.vPQRi);
  f.alpha = s[F16S_ALPHA];
  f.Mach = s[F16S_MACH];
  f.vcas = s[F16S_VC_KTS] * ktstofps;
  f.Vground = s[F16S_VG];
  f.vPilotAccelN.v[1] = s[F16S_NPY];
  f.vPilotAccelN.v[2] = s[F16S_NPZ];
  f.P.fcs_tef_control = s[F16S_TEF];
  f.P.fcs_left_aileron_pos_norm = s[F16S_AIL];
  f.P.fcs_elevator_pos_norm = s[F16S_ELEV];
  f.P.fcs_speedbrake_pos_deg = s[F16S_SB_DEG];
  f.M.fcs_roll_rate_pid.Input_prev = s[F16S_ROLL_INPREV];
  f.M.fcs_roll_rate_pid.I_out_total = s[F16S_ROLL_I];
  f.M.fcs_g_load_pid.Input_prev = s[F16S_PITCH_INPREV];
  f.M.fcs_g_load_pid.I_out_total = s[F16S_PITCH_I];
  f.M.fcs_yaw_load_pid.Input_prev = s[F16S_YAW_INPREV];
  f.M.fcs_yaw_load_pid.I_out_total = s[F16S_YAW_I];
  f.N2 = s[F16S_N2];
  f.N2norm = (f.N2 - modelk::idlen2) / (modelk::maxn2 - modelk::idlen2);
  f.Augmentation = s[F16S_AUG] > 0.5;
  for (int i = 0; i < modelk::n_contacts; i++)
    for (int k = 0; k < 3; k++) f.gear[i].LMultiplier[k].value = 0.0;
}

// ------------------------------------------------------------------ env layer (jsbsim_gym/jsbsim_gym.py restated)
// float32 cast chain of JSBSimEnv._get_current_single_observation (jsbsim_gym.py:172-197) under
// NumPy >= 2 scalar rules (float32 array element op python float -> float32 arithmetic).
float np_mod_f32(float a, float b) {   // numpy remainder for float32 (npy_divmodf)
  float mod = std::fmod(a, b);
  if (mod != 0.0f) { if ((b < 0) != (mod < 0)) mod += b; }
  else mod = std::copysign(0.0f, b);
  return mod;
}
float normalize_angle_mpi_pi_f32(float angle) {   // jsbsim_gym.py:60-78
  if (std::isnan(angle) || std::isinf(angle)) return 0.0f;
  const float two_pi = (float)(2 * M_PI), pi = (float)M_PI;
  angle = np_mod_f32(angle, two_pi);
  if (angle >= pi) angle -= two_pi;
  return angle;
}

struct Env {
  FDM sim;
  int current_step = 0;
  float goal[3] = {0, 0, 0};
  float last_distance = 0.0f;
  float frames[10][15];

  Env() {   // JSBSimEnv.__init__ (jsbsim_gym.py:151-155): construct, set-running, ic, run_ic
    sim.set_property("propulsion/set-running", -1);
    sim.set_property("ic/u-fps", 900.0);
    sim.set_property("ic/h-sl-ft", 5000.0);
    sim.RunIC();
    std::memset(frames, 0, sizeof(frames));
  }
  void single_obs(float* o) const {   // jsbsim_gym.py:172-197
    static const char* fmt[12] = {"position/lat-gc-rad", "position/long-gc-rad", "position/h-sl-meters", "velocities/mach",
                                  "aero/alpha-rad", "aero/beta-rad", "velocities/p-rad_sec", "velocities/q-rad_sec",
                                  "velocities/r-rad_sec", "attitude/phi-rad", "attitude/theta-rad", "attitude/psi-rad"};
    for (int i = 0; i < 12; i++) { double v = 0; sim.get_property(fmt[i], &v); o[i] = (float)v; }
    o[9] = normalize_angle_mpi_pi_f32(o[9]);
    o[10] = normalize_angle_mpi_pi_f32(o[10]);
    o[11] = normalize_angle_mpi_pi_f32(o[11]);
    o[0] *= 6.3781e6f;
    o[1] *= 6.3781e6f;
    for (int i = 0; i < 3; i++) o[12 + i] = goal[i];
  }
  // PositionReward: np.linalg.norm(goal - pos) on float32 (jsbsim_gym.py:499-500) = sqrt(x.dot(x)); NumPy's
  // float32 dot goes to BLAS sdot, whose scalar tail (n < 32) adds the float32 products in a double
  // accumulator and rounds the sum to float32 once (OpenBLAS kernel/x86_64/sdot.c).
  static float dist3(const float* o) {
    float dx = o[12] - o[0], dy = o[13] - o[1], dz = o[14] - o[2];
    float xx = dx * dx, yy = dy * dy, zz = dz * dz;
    double dot = 0.0;
    dot += xx; dot += yy; dot += zz;
    return std::sqrt((float)dot);
  }
  void reset(const float g[3]) {   // jsbsim_gym.py:289-331 + PositionReward.reset :511-519
    current_step = 0;
    sim.RunIC();
    sim.set_property("propulsion/set-running", -1);
    for (int i = 0; i < 3; i++) goal[i] = g[i];
    float o[15];
    single_obs(o);
    for (int r = 0; r < 10; r++) std::memcpy(frames[r], o, sizeof(o));
    last_distance = dist3(o);
  }
  // returns flags: bit0 terminated, bit1 truncated
  int step(const float a[4], float* reward_out) {   // jsbsim_gym.py:199-287 + PositionReward.step :487-509
    current_step += 1;
    sim.P.fcs_aileron_cmd_norm = (double)a[0];
    sim.P.fcs_elevator_cmd_norm = (double)a[1];
    sim.P.fcs_rudder_cmd_norm = (double)a[2];
    sim.P.fcs_throttle_cmd_norm = (double)a[3];
    for (int k = 0; k < 4; k++) {
      sim.tank_contents[0] = 1000.0;
      sim.tank_contents[1] = 1000.0;
      sim.P.gear_gear_cmd_norm = 0.0;
      sim.P.gear_gear_pos_norm = 0.0;
      sim.Run();
    }
    float o[15];
    single_obs(o);
    std::memmove(frames[0], frames[1], 9 * 15 * sizeof(float));
    std::memcpy(frames[9], o, sizeof(o));
    float reward = 0.0f;
    bool terminated = false, truncated = false;
    float altitude_m = o[2];
    if (altitude_m < 10.0f) { reward = -10.0f; terminated = true; }
    float d2 = (o[0] - goal[0]) * (o[0] - goal[0]) + (o[1] - goal[1]) * (o[1] - goal[1]);
    if (!terminated && std::sqrt(d2) < 100.0f && std::fabs(altitude_m - goal[2]) < 100.0f) { reward = 10.0f; terminated = true; }
    if (!terminated && current_step >= 1200) truncated = true;
    float d = dist3(o);
    reward += 0.01f * (last_distance - d);
    last_distance = d;
    *reward_out = reward;
    return (terminated ? 1 : 0) | (truncated ? 2 : 0);
  }
};

thread_local std::string g_err;

}  // namespace

// ===================================================================================== C ABI
extern "C" {

// ---- FGFDMExec-shaped handle (what a `jsbsim` stub module binds)
void* f16o_fdm_create(void) { return new FDM(); }
void f16o_fdm_destroy(void* h) { delete (FDM*)h; }
int f16o_fdm_set_property(void* h, const char* name, double v) { return ((FDM*)h)->set_property(name, v) ? 0 : -1; }
int f16o_fdm_get_property(void* h, const char* name, double* out) { return ((FDM*)h)->get_property(name, out) ? 0 : -1; }
int f16o_fdm_run_ic(void* h) { return ((FDM*)h)->RunIC() ? 0 : -1; }
int f16o_fdm_run(void* h) { return ((FDM*)h)->Run() ? 0 : -1; }
int f16o_num_state_fields(void) { return F16_NUM_STATE_FIELDS; }
void f16o_fdm_pack_state(void* h, double* out) { pack_state(*(FDM*)h, out); }
void f16o_fdm_unpack_state(void* h, const double* in) { unpack_state(*(FDM*)h, in); }

// ---- env-shaped handle (JSBSimEnv + PositionReward restated)
void* f16o_env_create(void) { return new Env(); }
void f16o_env_destroy(void* h) { delete (Env*)h; }
void* f16o_env_fdm(void* h) { return &((Env*)h)->sim; }
void f16o_env_reset(void* h, const float* goal, float* obs_out /*150*/) {
  Env* e = (Env*)h;
  e->reset(goal);
  if (obs_out) std::memcpy(obs_out, e->frames, sizeof(e->frames));
}
int f16o_env_step(void* h, const float* action, float* obs_out /*150*/, float* reward) {
  Env* e = (Env*)h;
  int fl = e->step(action, reward);
  if (obs_out) std::memcpy(obs_out, e->frames, sizeof(e->frames));
  return fl;
}

// ---- bounded CPU rollout for bench.py's cpu_baseline / --impl reference legs.
// n_envs independent envs, n_steps each, random actions from a splitmix/xorshift stream, auto-reset on
// done; work split contiguously over n_threads. Returns total env-steps executed.
static inline uint64_t splitmix64(uint64_t& x) {
  uint64_t z = (x += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline float u01(uint64_t& s) { return (float)((splitmix64(s) >> 40) * (1.0 / 16777216.0)); }

long long f16o_rollout(int n_envs, int n_steps, uint64_t seed, int n_threads, double* checksum_out) {
  if (n_threads < 1) n_threads = 1;
  std::vector<std::thread> th;
  std::vector<double> sums(n_threads, 0.0);
  std::vector<long long> counts(n_threads, 0);
  auto work = [&](int t) {
    int lo = (int)((long long)n_envs * t / n_threads), hi = (int)((long long)n_envs * (t + 1) / n_threads);
    for (int e = lo; e < hi; e++) {
      uint64_t s = seed * 0x100000001B3ull + (uint64_t)e;
      Env env;
      auto new_goal = [&](float* g) {
        float d = 1000.0f + 9000.0f * u01(s), b = 6.2831853f * u01(s), alt = 1000.0f + 3000.0f * u01(s);
        g[0] = d * std::cos(b); g[1] = d * std::sin(b); g[2] = alt;
      };
      float g[3];
      new_goal(g);
      env.reset(g);
      for (int k = 0; k < n_steps; k++) {
        float a[4] = {2 * u01(s) - 1, 2 * u01(s) - 1, 2 * u01(s) - 1, u01(s)};
        float r;
        int fl = env.step(a, &r);
        sums[t] += r;
        counts[t]++;
        if (fl) { new_goal(g); env.reset(g); }
      }
    }
  };
  for (int t = 0; t < n_threads; t++) th.emplace_back(work, t);
  for (auto& x : th) x.join();
  double cs = 0;
  long long n = 0;
  for (int t = 0; t < n_threads; t++) { cs += sums[t]; n += counts[t]; }
  if (checksum_out) *checksum_out = cs;
  return n;
}

// ---- batch trajectories for the parity tests: n_envs envs, each reset with goals[e] and driven by
// actions[k][e][:] for k < n_steps with no auto-reset; stepping of an env stops after its first done
// (later rows repeat the done flags with zero reward). frames: [n_steps][n_envs][15] newest frame,
// rewards: [n_steps][n_envs], flags: [n_steps][n_envs] (bit0 terminated, bit1 truncated, bit7 = env
// was already finished). final_states: [n_envs][F16_NUM_STATE_FIELDS] or NULL.
void f16o_batch_trajectory(int n_envs, int n_steps, const float* goals, const float* actions, float* frames,
                           float* rewards, uint8_t* flags, double* final_states, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  std::vector<std::thread> th;
  auto work = [&](int t) {
    int lo = (int)((long long)n_envs * t / n_threads), hi = (int)((long long)n_envs * (t + 1) / n_threads);
    for (int e = lo; e < hi; e++) {
      Env env;
      env.reset(goals + 3 * e);
      bool finished = false;
      int last = 0;
      for (int k = 0; k < n_steps; k++) {
        size_t idx = (size_t)k * n_envs + e;
        if (!finished) {
          float r;
          int fl = env.step(actions + idx * 4, &r);
          rewards[idx] = r;
          flags[idx] = (uint8_t)fl;
          if (fl) { finished = true; last = fl; }
        } else {
          rewards[idx] = 0.0f;
          flags[idx] = (uint8_t)(last | 0x80);
        }
        std::memcpy(frames + idx * 15, env.frames[9], 15 * sizeof(float));
      }
      if (final_states) pack_state(env.sim, final_states + (size_t)e * F16_NUM_STATE_FIELDS);
    }
  };
  for (int t = 0; t < n_threads; t++) th.emplace_back(work, t);
  for (auto& x : th) x.join();
}

}  // extern "C"
