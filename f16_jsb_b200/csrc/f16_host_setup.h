// f16_host_setup.h - host-side, double-precision preparation of everything the kernels treat as
// constant: the four mass-property sets (FGMassBalance::Run restated), the shared-memory table image
// (from f16_model_data.h) and the kinematic initial condition (FGPropagate::SetInitialState for the
// reference's IC, jsbsim_gym/jsbsim_gym.py:166-170). Included by f16_b200.cu (product) and by
// tests/hostsim (CPU debugging harness of the kernel source).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "f16_env.cuh"

namespace f16 {
namespace host {

// ---- FGMassBalance::Run for one (tank contents, previous-frame CG) configuration, in double
struct V3h { double x, y, z; };
V3h s2b(const V3h& cg, const V3h& r) { return {(cg.x - r.x) / 12.0, (r.y - cg.y) / 12.0, (cg.z - r.z) / 12.0}; }
void add_pm_inertia(double J[9], double mass_sl, const V3h& cg, const V3h& r) {
  V3h v = s2b(cg, r);
  V3h sv = {mass_sl * v.x, mass_sl * v.y, mass_sl * v.z};
  double xx = sv.x * v.x, yy = sv.y * v.y, zz = sv.z * v.z;
  double xy = -sv.x * v.y, xz = -sv.x * v.z, yz = -sv.y * v.z;
  J[0] += yy + zz; J[1] += xy; J[2] += xz;
  J[3] += xy; J[4] += xx + zz; J[5] += yz;
  J[6] += xz; J[7] += yz; J[8] += xx + yy;
}
// returns the new CG; tank inertia is evaluated about `cg_prev` (LoadInputs happens before Run)
V3h compute_mass_set(const double contents[4], const V3h& cg_prev, MassSet* ms) {
  using namespace f16data;
  const double lbtoslug = 1.0 / kSlugToLb;
  double tanksJ[9] = {0};
  double tw = 0;
  V3h tm = {0, 0, 0};
  for (int i = 0; i < n_tanks; ++i) {
    V3h loc = {tank_loc[i][0], tank_loc[i][1], tank_loc[i][2]};
    tw += contents[i];
    tm.x += loc.x * contents[i]; tm.y += loc.y * contents[i]; tm.z += loc.z * contents[i];
    add_pm_inertia(tanksJ, lbtoslug * contents[i], cg_prev, loc);
  }
  double weight = emptywt + tw + pilot_w;
  ms->mass = lbtoslug * weight;
  ms->inv_mass = 1.0 / ms->mass;
  V3h cg = {(emptywt * base_cg[0] + pilot_w * pilot_loc[0] + tm.x) / weight,
            (emptywt * base_cg[1] + pilot_w * pilot_loc[1] + tm.y) / weight,
            (emptywt * base_cg[2] + pilot_w * pilot_loc[2] + tm.z) / weight};
  double J[9];
  if (!negated_crossproduct_inertia) {
    double t[9] = {ixx, ixy, -ixz, ixy, iyy, iyz, -ixz, iyz, izz};
    memcpy(J, t, sizeof(J));
  } else {
    double t[9] = {ixx, -ixy, ixz, -ixy, iyy, -iyz, ixz, -iyz, izz};
    memcpy(J, t, sizeof(J));
  }
  add_pm_inertia(J, lbtoslug * emptywt, cg, {base_cg[0], base_cg[1], base_cg[2]});
  add_pm_inertia(J, lbtoslug * pilot_w, cg, {pilot_loc[0], pilot_loc[1], pilot_loc[2]});
  for (int i = 0; i < 9; ++i) J[i] += tanksJ[i];
  memcpy(ms->J, J, sizeof(J));
  double Ixx = J[0], Iyy = J[4], Izz = J[8], Ixy = -J[1], Ixz = -J[2], Iyz = -J[5];
  double k1 = (Iyy * Izz - Iyz * Iyz), k2 = (Iyz * Ixz + Ixy * Izz), k3 = (Ixy * Iyz + Iyy * Ixz);
  double denom = 1.0 / (Ixx * k1 - Ixy * k2 - Ixz * k3);
  k1 *= denom; k2 *= denom; k3 *= denom;
  double k4 = (Izz * Ixx - Ixz * Ixz) * denom, k5 = (Ixy * Ixz + Iyz * Ixx) * denom, k6 = (Ixx * Iyy - Ixy * Ixy) * denom;
  double Ji[9] = {k1, k2, k3, k2, k4, k5, k3, k5, k6};
  memcpy(ms->Jinv, Ji, sizeof(Ji));
  V3h rp = s2b(cg, {AERORP[0], AERORP[1], AERORP[2]});
  V3h ey = s2b(cg, {EYEPOINT[0], EYEPOINT[1], EYEPOINT[2]});
  V3h th = s2b(cg, {thruster_loc[0], thruster_loc[1], thruster_loc[2]});
  ms->r_rp[0] = rp.x; ms->r_rp[1] = rp.y; ms->r_rp[2] = rp.z;
  ms->r_eye[0] = ey.x; ms->r_eye[1] = ey.y; ms->r_eye[2] = ey.z;
  ms->r_thr[0] = th.x; ms->r_thr[1] = th.y; ms->r_thr[2] = th.z;
  ms->cg[0] = cg.x; ms->cg[1] = cg.y; ms->cg[2] = cg.z;
  // the ground-reaction path is entered below kGroundReach: every contact must lie inside that sphere
  const double sl[n_structure][3] = F16_STRUCT_LOC;
  for (int i = 0; i < n_structure; ++i) {
    V3h r = s2b(cg, {sl[i][0], sl[i][1], sl[i][2]});
    if (std::sqrt(r.x * r.x + r.y * r.y + r.z * r.z) > kGroundReach - 1.0) { fprintf(stderr, "f16: contact %d is out of kGroundReach\n", i); abort(); }
    ms->r_ct[i][0] = r.x; ms->r_ct[i][1] = r.y; ms->r_ct[i][2] = r.z;
  }
  return cg;
}
inline void convert_mass_set(const MassSetT<double>& a, MassSetT<float>* b) {
  b->mass = (float)a.mass; b->inv_mass = (float)a.inv_mass;
  for (int i = 0; i < 9; ++i) { b->J[i] = (float)a.J[i]; b->Jinv[i] = (float)a.Jinv[i]; }
  for (int i = 0; i < 3; ++i) { b->r_rp[i] = (float)a.r_rp[i]; b->r_eye[i] = (float)a.r_eye[i]; b->r_thr[i] = (float)a.r_thr[i]; b->cg[i] = (float)a.cg[i]; }
  for (int i = 0; i < 7; ++i)
    for (int k = 0; k < 3; ++k) b->r_ct[i][k] = (float)a.r_ct[i][k];
}
void build_mass_sets(MassSet out[MS_COUNT]) {
  double ic[4], fl[4];
  for (int i = 0; i < 4; ++i) { ic[i] = f16data::tank_contents0[i]; fl[i] = ic[i]; }
  fl[0] = 1000.0; fl[1] = 1000.0;                    // jsbsim_gym.py:227-228
  V3h cg0 = {0, 0, 0};                               // FGMassBalance ctor: vXYZcg = 0 before the first Run
  V3h cg_ic = compute_mass_set(ic, cg0, &out[MS_IC_FIRST]);
  compute_mass_set(ic, cg_ic, &out[MS_IC]);
  V3h cg_fl = compute_mass_set(fl, cg_ic, &out[MS_FLIGHT_FIRST]);
  compute_mass_set(fl, cg_fl, &out[MS_FLIGHT]);
}

// ---- 1-D clamped interpolation on the host (FGTable::GetValue) for the Mach union grid
double interp1(const double* x, const double* y, int n, double key) {
  if (key <= x[0]) return y[0];
  if (key >= x[n - 1]) return y[n - 1];
  int r = 1;
  while (r < n - 1 && x[r] < key) r++;
  double f = (key - x[r - 1]) / (x[r] - x[r - 1]);
  return f * (y[r] - y[r - 1]) + y[r - 1];
}

template <typename R>
void build_tables(Tables<R>* T) {
  using namespace f16data;
  constexpr int NA = f16::NA, NDE = f16::NDE, NB7 = f16::NB7, NB13 = f16::NB13;
  memset(T, 0, sizeof(*T));
  for (int i = 0; i < NA; ++i) {
    for (int k = 0; k < A1_N; ++k) T->A1[i][k] = (R)A1[i][k];
    for (int k = A1_N; k < A1_N + 4; ++k) T->A1[i][k] = (R)0;        // row padding (bank spreading, f16_model.cuh)
    for (int j = 0; j < NDE; ++j)
      for (int k = 0; k < 4; ++k) T->AE[i][j][k] = (R)AE[i][j][k];
    for (int j = 0; j < NB7; ++j)
      for (int k = 0; k < 4; ++k) T->AB7[i][j][k] = (R)AB7[i][j][k];
    for (int j = 0; j < NB13; ++j)
      for (int k = 0; k < 2; ++k) T->AB13[i][j][k] = (R)AB13[i][j][k];
  }
  auto fill_seg = [](R (*seg)[2], const double* x, int n) {
    for (int r = 1; r < n; ++r) { seg[r][0] = (R)x[r - 1]; seg[r][1] = (R)(1.0 / (x[r] - x[r - 1])); }
  };
  // packed entries for locate_uniform_packed: {x[i-2], 1/w[i-1], x[i-1], 1/w[i], x[i], 1/w[i+1], 0, 0}, w[j] = x[j] - x[j-1]
  auto fill_segp = [](R (*segp)[8], const double* x, int n) {
    for (int i = 0; i <= n; ++i)
      for (int k = 0; k < 8; ++k) segp[i][k] = (R)0;
    auto inv_w = [&](int j) { return (j >= 1 && j <= n - 1) ? 1.0 / (x[j] - x[j - 1]) : 0.0; };
    for (int i = 1; i <= n - 1; ++i) {
      segp[i][0] = (R)(i >= 2 ? x[i - 2] : x[0]); segp[i][1] = (R)inv_w(i - 1);
      segp[i][2] = (R)x[i - 1];                   segp[i][3] = (R)inv_w(i);
      segp[i][4] = (R)x[i];                       segp[i][5] = (R)inv_w(i + 1);
    }
  };
  fill_segp(T->segp_alpha, alpha_bp, NA);
  fill_segp(T->segp_de, de_bp, NDE);
  fill_seg(T->seg_alpha, alpha_bp, NA);
  fill_seg(T->seg_de, de_bp, NDE);
  fill_seg(T->seg_b7, b7_bp, NB7);
  // the frame derives the 7-point beta segment from the 13-point one: the coarse grid must be every other fine point
  static_assert(NB13 == 2 * NB7 - 1, "beta grids");
  for (int i = 0; i < NB7; ++i)
    if (b7_bp[i] != b13_bp[2 * i]) { std::fprintf(stderr, "f16: beta breakpoint grids are not nested\n"); std::abort(); }
  fill_seg(T->seg_b13, b13_bp, NB13);
  fill_segp(T->segp_b13, b13_bp, NB13);
  // union of the Mach breakpoints of the nine Mach tables; every table is piecewise linear with
  // clamped ends, so resampling it on the union grid reproduces it exactly
  std::vector<double> grid;
  auto add = [&](const double* x, int n) {
    for (int i = 0; i < n; ++i) {
      bool seen = false;
      for (double g : grid) seen |= (g == x[i]);
      if (!seen) grid.push_back(x[i]);
    }
  };
  add(x_CDmach, n_CDmach); add(x_CYb_M, n_CYb_M); add(x_Clb_M, n_Clb_M); add(x_Clda_M, n_Clda_M); add(x_Cldr_M, n_Cldr_M);
  add(x_Cma_M, n_Cma_M); add(x_Cnb_M, n_Cnb_M); add(x_Cnda_M, n_Cnda_M); add(x_Cndr_M, n_Cndr_M);
  std::sort(grid.begin(), grid.end());
  const double gen_grid[NMACH] = F16_MACH_BP;   // the generator's copy (compared as immediates in the kernel)
  if ((int)grid.size() != NMACH) { fprintf(stderr, "f16: Mach union grid has %d points, expected %d\n", (int)grid.size(), NMACH); abort(); }
  for (int i = 0; i < NMACH; ++i)
    if (grid[i] != gen_grid[i]) { fprintf(stderr, "f16: Mach union grid mismatch at %d\n", i); abort(); }
  fill_seg(T->seg_mach, gen_grid, NMACH);
  for (int i = 0; i < NMACH; ++i) {
    double m = grid[i];
    T->MT[i][MT_CDmach] = (R)interp1(x_CDmach, y_CDmach, n_CDmach, m);
    T->MT[i][MT_CYb_M] = (R)interp1(x_CYb_M, y_CYb_M, n_CYb_M, m);
    T->MT[i][MT_Clb_M] = (R)interp1(x_Clb_M, y_Clb_M, n_Clb_M, m);
    T->MT[i][MT_Clda_M] = (R)interp1(x_Clda_M, y_Clda_M, n_Clda_M, m);
    T->MT[i][MT_Cldr_M] = (R)interp1(x_Cldr_M, y_Cldr_M, n_Cldr_M, m);
    T->MT[i][MT_Cma_M] = (R)interp1(x_Cma_M, y_Cma_M, n_Cma_M, m);
    T->MT[i][MT_Cnb_M] = (R)interp1(x_Cnb_M, y_Cnb_M, n_Cnb_M, m);
    T->MT[i][MT_Cnda_M] = (R)interp1(x_Cnda_M, y_Cnda_M, n_Cnda_M, m);
    T->MT[i][MT_Cndr_M] = (R)interp1(x_Cndr_M, y_Cndr_M, n_Cndr_M, m);
  }
  for (int i = 0; i < n_kCLge; ++i) { T->kclge_x[i] = (R)x_kCLge[i]; T->kclge_y[i] = (R)y_kCLge[i]; }
  static_assert(n_idle_mach == 6 && n_mil_mach == 8 && n_aug_mach == 14 && NEH == 8, "engine table shape");
  for (int c = 0; c < NEH; ++c) {
    if (eng_alt_bp[c] != -10000.0 + 10000.0 * c) { fprintf(stderr, "f16: engine altitude grid is not uniform\n"); abort(); }
    for (int r = 0; r < 6; ++r) T->eng_idle[r][c] = (R)idle_tbl[r][c];
    for (int r = 0; r < 8; ++r) T->eng_mil[r][c] = (R)mil_tbl[r][c];
    for (int r = 0; r < 14; ++r) T->eng_aug[r][c] = (R)aug_tbl[r][c];
  }
  for (int r = 0; r < 14; ++r) {
    double want = 0.2 * r;
    if (fabs(aug_mach_bp[r] - want) > 1e-12 || (r < 8 && fabs(mil_mach_bp[r] - want) > 1e-12) || (r < 6 && fabs(idle_mach_bp[r] - want) > 1e-12)) {
      fprintf(stderr, "f16: engine Mach grid is not uniform\n");
      abort();
    }
  }
}

// ---- FGPropagate::SetInitialState for the reference's IC (u = 900 ft/s, h = 5000 ft ASL at
// geocentric lat = lon = 0, level, heading north; jsbsim_gym.py:166-170) -> packed kinematic state
void initial_condition(double u_fps, double h_sl_ft, double* packed) {
  for (int i = 0; i < F16_NUM_STATE_FIELDS; ++i) packed[i] = 0.0;
  // FGLocation::SetPositionGeodetic(0,0,0) -> (a,0,0); SetAltitudeASLFtIC (setgeoc): radius = slr + h
  double x = kEarthA;
  double slr = kEarthA * kEc / std::sqrt(1.0 - kE2 * 1.0);
  double radius = slr + h_sl_ft;
  x *= radius / std::sqrt(x * x);
  // Tec2l at lon = 0, geodetic lat = 0; Ti2ec = I (epa = 0) -> Ti2l = Tec2l; qAttitudeECI = quat(Ti2l) * identity
  double T[3][3] = {{-1.0 * 0.0, -0.0 * 0.0, 1.0}, {-0.0, 1.0, 0.0}, {-1.0 * 1.0, -0.0 * 1.0, -0.0}};
  double tq[4] = {1.0 + T[0][0] + T[1][1] + T[2][2], 1.0 + T[0][0] - T[1][1] - T[2][2], 1.0 - T[0][0] + T[1][1] - T[2][2],
                  1.0 - T[0][0] - T[1][1] + T[2][2]};
  int idx = 0;
  for (int i = 1; i < 4; ++i)
    if (tq[i] > tq[idx]) idx = i;
  double q[4];
  if (idx != 0) { fprintf(stderr, "f16: unexpected IC quaternion branch\n"); abort(); }
  q[0] = 0.50 * std::sqrt(tq[0]);
  q[1] = 0.25 * (T[1][2] - T[2][1]) / q[0];
  q[2] = 0.25 * (T[2][0] - T[0][2]) / q[0];
  q[3] = 0.25 * (T[0][1] - T[1][0]) / q[0];
  // Ti2b from q
  double q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
  double B[3][3] = {{q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, 2.0 * (q1 * q2 + q0 * q3), 2.0 * (q1 * q3 - q0 * q2)},
                    {2.0 * (q1 * q2 - q0 * q3), q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3, 2.0 * (q2 * q3 + q0 * q1)},
                    {2.0 * (q1 * q3 + q0 * q2), 2.0 * (q2 * q3 - q0 * q1), q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3}};
  double uvw[3] = {u_fps, 0.0, 0.0};
  double ri[3] = {x, 0.0, 0.0};
  double wi[3], vi[3];
  for (int i = 0; i < 3; ++i) wi[i] = 0.0 + B[i][2] * kEarthOmega;                       // vPQRi = vPQR + Ti2b w_p
  double wxr[3] = {-kEarthOmega * ri[1], kEarthOmega * ri[0], 0.0};
  for (int i = 0; i < 3; ++i) vi[i] = (B[0][i] * uvw[0] + B[1][i] * uvw[1] + B[2][i] * uvw[2]) + wxr[i];   // Tb2i uvw + w_p x r
  for (int i = 0; i < 4; ++i) packed[F16S_Q0 + i] = q[i];
  for (int i = 0; i < 3; ++i) {
    packed[F16S_WI_X + i] = wi[i];
    packed[F16S_RI_X + i] = ri[i];
    packed[F16S_VI_X + i] = vi[i];
    packed[F16S_VI1_X + i] = vi[i];
    packed[F16S_VI2_X + i] = vi[i];
  }
}


}  // namespace host
}  // namespace f16
