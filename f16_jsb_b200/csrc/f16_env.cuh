// f16_env.cuh - the env layer of one environment: everything JSBSimEnv.step / reset and
// PositionReward.step / reset (jsbsim_gym/jsbsim_gym.py:172-331, 487-519) do around the FDM frames,
// plus the DummyVecEnv auto-reset convention (stable_baselines3/common/vec_env/dummy_vec_env.py:63-72).
// Per-thread code, called by the kernels in f16_b200.cu (and, for CPU-side debugging of this very
// source, by tests/hostsim).
#pragma once
#include "../../include/f16_state_fields.h"
#include "f16_model.cuh"

namespace f16 {

// ------------------------------------------------------------------------------------ SoA field lists
// K fields (always double) and R fields (double in parity mode, float in throughput mode).
#define F16_KFIELDS(X) X(q[0]) X(q[1]) X(q[2]) X(q[3]) X(ri[0]) X(ri[1]) X(ri[2]) X(vi[0]) X(vi[1]) X(vi[2]) X(epa)
#define F16_RFIELDS(X)                                                                                 \
  X(wi[0]) X(wi[1]) X(wi[2]) X(vi1[0]) X(vi1[1]) X(vi1[2]) X(vi2[0]) X(vi2[1]) X(vi2[2])               \
  X(ai0[0]) X(ai0[1]) X(ai0[2]) X(ai1[0]) X(ai1[1]) X(ai1[2]) X(wdot[0]) X(wdot[1]) X(wdot[2])         \
  X(abody[0]) X(abody[1]) X(abody[2]) X(pqr[0]) X(pqr[1]) X(pqr[2]) X(alpha) X(mach) X(vc) X(vg)       \
  X(npy) X(npz) X(tef) X(ail) X(elev) X(sb) X(roll_ip) X(roll_I) X(pitch_ip) X(pitch_I) X(yaw_ip)      \
  X(yaw_I) X(n2) X(aug)
constexpr int NKF = 11, NRF = 42;
enum { EF_GOAL_X = 0, EF_GOAL_Y, EF_GOAL_Z, EF_LAST_DIST, EF_STEP, EF_EP_RET, EF_EP_LEN, EF_EPISODES, NEF };

// packed double[F16_NUM_STATE_FIELDS] (include/f16_state_fields.h) <-> Veh. The enum lists Q, WI, RI,
// VI, EPA first and then every remaining R field in F16_RFIELDS order.
template <typename R>
F16_HD void veh_from_packed(Veh<R>& s, const double* a) {
  double kk[NKF];
  double rr[NRF];
  kk[0] = a[F16S_Q0]; kk[1] = a[F16S_Q1]; kk[2] = a[F16S_Q2]; kk[3] = a[F16S_Q3];
  kk[4] = a[F16S_RI_X]; kk[5] = a[F16S_RI_Y]; kk[6] = a[F16S_RI_Z];
  kk[7] = a[F16S_VI_X]; kk[8] = a[F16S_VI_Y]; kk[9] = a[F16S_VI_Z]; kk[10] = a[F16S_EPA];
  rr[0] = a[F16S_WI_X]; rr[1] = a[F16S_WI_Y]; rr[2] = a[F16S_WI_Z];
  for (int i = 0; i < NRF - 3; ++i) rr[3 + i] = a[F16S_VI1_X + i];
  int f = 0;
#define X(m) s.m = kk[f++];
  F16_KFIELDS(X)
#undef X
  f = 0;
#define X(m) s.m = (R)rr[f++];
  F16_RFIELDS(X)
#undef X
}
template <typename R>
F16_HD void veh_to_packed(const Veh<R>& s, double* a) {
  double kk[NKF];
  double rr[NRF];
  int f = 0;
#define X(m) kk[f++] = s.m;
  F16_KFIELDS(X)
#undef X
  f = 0;
#define X(m) rr[f++] = (double)s.m;
  F16_RFIELDS(X)
#undef X
  a[F16S_Q0] = kk[0]; a[F16S_Q1] = kk[1]; a[F16S_Q2] = kk[2]; a[F16S_Q3] = kk[3];
  a[F16S_RI_X] = kk[4]; a[F16S_RI_Y] = kk[5]; a[F16S_RI_Z] = kk[6];
  a[F16S_VI_X] = kk[7]; a[F16S_VI_Y] = kk[8]; a[F16S_VI_Z] = kk[9]; a[F16S_EPA] = kk[10];
  a[F16S_WI_X] = rr[0]; a[F16S_WI_Y] = rr[1]; a[F16S_WI_Z] = rr[2];
  for (int i = 0; i < NRF - 3; ++i) a[F16S_VI1_X + i] = rr[3 + i];
}

// env-layer scalars of one environment (E fields)
struct EnvScalars {
  float gx, gy, gz;       // goal, metres (jsbsim_gym.py:321-323)
  float last_d;           // PositionReward.last_distance
  int32_t step;           // JSBSimEnv.current_step
  float ep_ret;           // Monitor: running episode return
  int32_t ep_len;         // Monitor: running episode length
  uint32_t episodes;      // episodes started (Philox counter for goal sampling)
};

// ------------------------------------------------------------------------------------ observation frame
// jsbsim_gym.py:172-197: float32 store of each double property, wrap of phi/theta/psi in float32,
// float32 multiply of lat/lon by 6.3781e6.
F16_HD void props_to_frame(const double* p12, float* o) {
  for (int i = 0; i < 12; ++i) o[i] = (float)p12[i];
  o[9] = wrap_mpi_pi_f32(o[9]);
  o[10] = wrap_mpi_pi_f32(o[10]);
  o[11] = wrap_mpi_pi_f32(o[11]);
  o[0] = fmul_rn(o[0], 6.3781e6f);
  o[1] = fmul_rn(o[1], 6.3781e6f);
}
template <typename R>
F16_HD void frame_from_fdm(const Veh<R>& s, const FrameObs<R>& fo, float* o) {
  R phi, tht, psi;
  euler_from_tl2b<R>(fo, phi, tht, psi);
  o[0] = (float)Mx<R>::fatan2_(fo.ze, fo.rxy);   // position/lat-gc-rad
  o[1] = (float)Mx<R>::fatan2_(fo.ye, fo.xe);    // position/long-gc-rad
  o[2] = (float)(fo.h_ft * kFtToM);
  o[3] = (float)s.mach;
  o[4] = (float)s.alpha;
  o[5] = (float)fo.beta;
  o[6] = (float)fo.pqr[0];
  o[7] = (float)fo.pqr[1];
  o[8] = (float)fo.pqr[2];
  o[9] = wrap_mpi_pi_f32((float)phi);
  o[10] = wrap_mpi_pi_f32((float)tht);
  o[11] = wrap_mpi_pi_f32((float)psi);
  o[0] = fmul_rn(o[0], 6.3781e6f);
  o[1] = fmul_rn(o[1], 6.3781e6f);
}
// PositionReward distance (jsbsim_gym.py:499-500): np.linalg.norm of a float32 3-vector = sqrt(x.dot(x)).
// NumPy's float32 dot is BLAS sdot, whose scalar tail adds the float32-rounded products in a double
// accumulator and rounds the sum to float32 once; reproduced here so rewards match bit for bit.
F16_HD float goal_distance(const float* o, float gx, float gy, float gz) {
  float dx = fsub_rn(gx, o[0]), dy = fsub_rn(gy, o[1]), dz = fsub_rn(gz, o[2]);
  double dot = (double)fmul_rn(dx, dx);
  dot += (double)fmul_rn(dy, dy);
  dot += (double)fmul_rn(dz, dz);
  return fsqrt_rn((float)dot);
}
// goal ~ (d cos b, d sin b, alt), d~U[1000,10000), b~U[0,2pi), alt~U[1000,4000) (jsbsim_gym.py:315-323)
F16_HD void sample_goal(uint64_t seed, uint64_t env_id, uint32_t episode, float* g) {
  uint32_t r[4];
  philox4x32_10((uint32_t)env_id, (uint32_t)(env_id >> 32), episode, 0x60A1u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  float d = 1000.0f + 9000.0f * u01_from_u32(r[0]);
  float bearing = 6.2831853071795864769f * u01_from_u32(r[1]);
  float alt = 1000.0f + 3000.0f * u01_from_u32(r[2]);
  float sb, cb;
  sincosf(bearing, &sb, &cb);
  g[0] = d * cb; g[1] = d * sb; g[2] = alt;
}
// action_space.sample()-like uniform action (Box low [-1,-1,-1,0], high [1,1,1,1]; jsbsim_gym.py:143-148)
F16_HD void sample_action(uint64_t seed, uint64_t env_id, uint32_t step_counter, float* act) {
  uint32_t r4[4];
  philox4x32_10((uint32_t)env_id, (uint32_t)(env_id >> 32), step_counter, 0xAC71u, (uint32_t)seed, (uint32_t)(seed >> 32), r4);
  act[0] = 2.0f * u01_from_u32(r4[0]) - 1.0f;
  act[1] = 2.0f * u01_from_u32(r4[1]) - 1.0f;
  act[2] = 2.0f * u01_from_u32(r4[2]) - 1.0f;
  act[3] = u01_from_u32(r4[3]);
}

// ------------------------------------------------------------------------------------ snapshot bring-up
// Canonical post-reset state, always in double: fresh FDM -> constructor's run_ic()
// (two zero-dt frames) -> reset()'s run_ic() (two more) -> propulsion/set-running (N2 = 100 %,
// augmentation off). Gear is still down and both internal tanks still hold their initial 1500 lb
// during these frames; the FCS components tick with dt = 1/120. `out` receives the packed state
// followed by the twelve STATE_FORMAT properties (jsbsim_gym.py:12-25) as doubles.
F16_HD void compute_snapshot(const Tables<double>& T, const MassSetT<double>* msets, const double* ic_state, double* out) {
  Veh<double> s;
  veh_from_packed(s, ic_state);   // kinematic IC, everything else zero (fresh FDM)
  Cmd<double> cmd = {0.0, 0.0, 0.0, 0.0};
  FrameObs<double> fo;
  FrameCfg cfg;
  cfg.dt = 0.0;
  cfg.gear = 1.0;
  for (int ic = 0; ic < 2; ++ic) {
    for (int k = 0; k < 2; ++k) {
      cfg.mass_set = (ic == 0 && k == 0) ? MS_IC_FIRST : MS_IC;
      fdm_frame<double, true>(s, T, msets, cfg, cmd, false, fo);
    }
    for (int i = 0; i < 3; ++i) { s.vi1[i] = s.vi[i]; s.vi2[i] = s.vi[i]; }   // InitializeDerivatives
  }
  s.n2 = f16data::idlen2 + 1.0 * (f16data::maxn2 - f16data::idlen2);          // InitRunning + GetSteadyState
  s.aug = 0.0;
  veh_to_packed(s, out);
  double phi, tht, psi;
  euler_from_tl2b<double>(fo, phi, tht, psi);
  double* p = out + F16_NUM_STATE_FIELDS;
  p[0] = atan2(fo.ze, fo.rxy); p[1] = atan2(fo.ye, fo.xe); p[2] = fo.h_ft * kFtToM; p[3] = s.mach; p[4] = s.alpha; p[5] = fo.beta;
  p[6] = fo.pqr[0]; p[7] = fo.pqr[1]; p[8] = fo.pqr[2]; p[9] = phi; p[10] = tht; p[11] = psi;
}

// ------------------------------------------------------------------------------------ reset / step of one env
// JSBSimEnv.reset + PositionReward.reset (jsbsim_gym.py:289-331, 511-519): restore the snapshot, set
// the goal, build the reset frame (the caller replicates it into all ten rows).
template <typename R>
F16_HD void env_reset_one(Veh<R>& s, EnvScalars& es, const double* snapshot, const double* snapshot_props,
                          const float* goal, float* frame16, bool keep_state = false) {
  // keep_state: carry-over reset in two parts - the flight-dynamics state is left as the episode ended and
  // veh_carryover_reset() brings it up afterwards (the observation after reset is the same either way: its
  // twelve properties only depend on the initial condition)
  if (!keep_state) veh_from_packed(s, snapshot);
  es.gx = goal[0]; es.gy = goal[1]; es.gz = goal[2];
  float o[12];
  props_to_frame(snapshot_props, o);
  for (int i = 0; i < 12; ++i) frame16[i] = o[i];
  frame16[12] = es.gx; frame16[13] = es.gy; frame16[14] = es.gz; frame16[15] = 0.0f;
  es.last_d = goal_distance(o, es.gx, es.gy, es.gz);
  es.step = 0;
  es.ep_ret = 0.0f;
  es.ep_len = 0;
}

// ------------------------------------------------------------------------------------ carry-over reset
// What JSBSimEnv.reset really does to an env object that already exists (jsbsim_gym.py:305-306): run_ic()
// re-initialises only FGPropagate's state from the initial condition and runs two frames with integration
// suspended; propulsion/set-running restarts the engine. Everything else leaks from the episode that just
// ended: the FCS actuator positions and PID memories (they even tick twice more, with dt = 1/120, on the last
// action, whose fcs/*-cmd-norm properties stay set), the Auxiliary outputs the FCS reads one frame late, the
// last accelerations (pilot load factors) and - unlike a fresh construct - the gear is already up and the tanks
// already hold 1000 lb, so the mass properties are the in-flight ones from the very first frame.
// The default reset restores the canonical fresh-env snapshot instead (deviation 1); this is the
// opt-in alternative that reproduces the second and later episodes of one reference env object.
// `used` = the env has been stepped since it was constructed (gear up, tanks refilled to 1000 lb); an env that
// was only ever reset still has its gear down and 1500 lb in the tanks and is brought up like the constructor does.
// Not modelled: the fuel burnt in the last frame of the previous episode (<= 0.13 lb of 19 630 lb), which
// JSBSim's MassBalance would see during these two frames because reset() does not refill the tanks.
constexpr uint32_t kEpisodeUsedBit = 0x80000000u;   // EnvScalars::episodes, top bit: the env was brought up by a carry-over reset
template <typename R>
F16_HD void veh_carryover_reset(Veh<R>& s, const Tables<R>& T, const MassSetT<R>* msets, const double* snapshot, const float* last_act,
                                bool used) {
  Veh<R> ic;
  veh_from_packed(ic, snapshot);
  for (int i = 0; i < 4; ++i) s.q[i] = ic.q[i];                     // FGPropagate::SetInitialState
  for (int i = 0; i < 3; ++i) { s.ri[i] = ic.ri[i]; s.vi[i] = ic.vi[i]; s.wi[i] = ic.wi[i]; }
  s.epa = ic.epa;
  Cmd<R> cmd = {(R)last_act[0], (R)last_act[1], (R)last_act[2], (R)last_act[3]};
  FrameObs<R> fo;
  FrameCfg cfg;
  cfg.dt = 0.0;
  cfg.gear = used ? 0.0 : 1.0;
  cfg.mass_set = used ? MS_FLIGHT : MS_IC;
  for (int k = 0; k < 2; ++k) fdm_frame<R, true, false>(s, T, msets, cfg, cmd, false, fo);
  for (int i = 0; i < 3; ++i) { s.vi1[i] = (R)s.vi[i]; s.vi2[i] = (R)s.vi[i]; }   // InitializeDerivatives
  s.n2 = (R)(f16data::idlen2 + 1.0 * (f16data::maxn2 - f16data::idlen2));       // InitRunning + GetSteadyState
  s.aug = R(0);
}

// Explicit reset() of one env in carry-over mode (f16_reset_carryover): what a second, third, ... call of
// JSBSimEnv.reset does to the same env object. An env that was never reset has no constructor state yet and gets
// the canonical bring-up. last_act: the action of the env's last step (ignored until it has been stepped).
template <typename R>
F16_HD void env_carryover_reset_one(Veh<R>& s, EnvScalars& es, const Tables<R>& T, const MassSetT<R>* msets, const double* snapshot,
                                    const double* snapshot_props, const float* goal, const float* last_act, float* frame16) {
  const uint32_t ep = es.episodes & ~kEpisodeUsedBit;
  if (ep == 0) {
    es.episodes = 1;
    env_reset_one<R>(s, es, snapshot, snapshot_props, goal, frame16);
    return;
  }
  const bool used = (es.episodes & kEpisodeUsedBit) != 0 || es.step > 0;
  const float none[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  veh_carryover_reset<R>(s, T, msets, snapshot, used ? last_act : none, used);
  es.episodes = (ep + 1) | (used ? kEpisodeUsedBit : 0u);
  env_reset_one<R>(s, es, snapshot, snapshot_props, goal, frame16, true);
}

enum { STEP_ACTIVE = 1, STEP_RESET = 2, STEP_TERMINAL = 4, STEP_DONE = 8, STEP_TRUNCATED = 16, STEP_CRASH = 32, STEP_GOAL = 64,
       STEP_NEAR_GROUND = 128 };

// ------------------------------------------------------------------------------------ one env-step
// On return frame16 is the newest row of the env's observation stack (or, if the env auto-reset, the
// reset frame, with tframe16 holding the terminal step's newest row).
//
// Ground reactions (f16_ground.cuh), three instantiations. GROUND_OFF: no ground at all (the FP32
// throughput mode's default: identical code to a model without contacts). GROUND_DETECT (hot): never
// includes ground forces, but if a contact point reached the surface in some frame, and that can have
// changed anything the caller will see, it returns STEP_NEAR_GROUND with s / es / the outputs in an
// unspecified state and the caller redoes the step from the env's saved state with GROUND_FULL (cold: a
// separate, non-inlined function on the device), where every touching frame has its accelerations
// replaced by ground_fix(). "Can have changed": the forces of
// frame k act on the rates from frame k+1 and on the position from frame k+2, so a first contact in the
// last frame only changes the stored accelerations - irrelevant when the env is reset in this very step.
//
// Optional L2 prefetch issued by each lane just before the last FDM frame: the caller's observation
// rows are needed right after that frame, and one frame of compute covers the HBM latency.
struct PrefetchHint { const char* ptr; int count; int stride; };   // count < 0: one bulk (TMA) prefetch of `stride` bytes

// observation frame, reward, termination, auto-reset - everything after the last frame (jsbsim_gym.py:235-263)
template <typename R>
F16_HD int env_step_epilogue(Veh<R>& s, EnvScalars& es, const FrameObs<R>& fo, const double* snapshot, const double* snapshot_props,
                             uint64_t seed, uint64_t env_id, int auto_reset, float* frame16, float* tframe16, float* reward_out,
                             float* ep_ret_out, int32_t* ep_len_out) {
  // all on the float32 frame (jsbsim_gym.py:237-261)
  float o[12];
  frame_from_fdm<R>(s, fo, o);
  float reward = 0.0f;
  bool terminated = false, truncated = false;
  int flags = STEP_ACTIVE;
  const float alt = o[2];
  if (alt < 10.0f) { reward = -10.0f; terminated = true; flags |= STEP_CRASH; }
  {
    float ex = fsub_rn(o[0], es.gx), ey = fsub_rn(o[1], es.gy);
    float d2 = fadd_rn(fmul_rn(ex, ex), fmul_rn(ey, ey));
    if (!terminated && fsqrt_rn(d2) < 100.0f && fabsf(fsub_rn(alt, es.gz)) < 100.0f) { reward = 10.0f; terminated = true; flags |= STEP_GOAL; }
  }
  if (!terminated && es.step >= 1200) truncated = true;
  // PositionReward.step (jsbsim_gym.py:487-509): reward += 1e-2 * (last_distance - distance), float32
  const float d = goal_distance(o, es.gx, es.gy, es.gz);
  reward = fadd_rn(reward, fmul_rn(0.01f, fsub_rn(es.last_d, d)));
  es.last_d = d;
  es.ep_ret += reward;
  es.ep_len += 1;
  *reward_out = reward;
  for (int i = 0; i < 12; ++i) frame16[i] = o[i];
  frame16[12] = es.gx; frame16[13] = es.gy; frame16[14] = es.gz; frame16[15] = 0.0f;
  if (truncated) flags |= STEP_TRUNCATED;
  if (terminated || truncated) {
    flags |= STEP_DONE;
    *ep_ret_out = es.ep_ret;
    *ep_len_out = es.ep_len;
    if (auto_reset) {
      for (int i = 0; i < 16; ++i) tframe16[i] = frame16[i];
      flags |= STEP_RESET | STEP_TERMINAL;
      const uint32_t ep = (es.episodes & ~kEpisodeUsedBit) + 1;
      es.episodes = auto_reset == 2 ? (ep | kEpisodeUsedBit) : ep;
      float g[3];
      sample_goal(seed, env_id, ep, g);
      env_reset_one<R>(s, es, snapshot, snapshot_props, g, frame16, auto_reset == 2);
    }
  }
  return flags;
}

enum { GROUND_OFF = 0, GROUND_DETECT = 1, GROUND_FULL = 2 };
template <typename R, int GMODE, bool FRAME_SYNC = false>
F16_HD int env_step_one(Veh<R>& s, EnvScalars& es, const Tables<R>& T, const MassSetT<R>* msets, const MassSetT<double>* msets_d,
                        const double* snapshot, const double* snapshot_props, const float* act, uint64_t seed, uint64_t env_id,
                        int auto_reset, float* frame16, float* tframe16, float* reward_out, float* ep_ret_out, int32_t* ep_len_out,
                        PrefetchHint pf = PrefetchHint{nullptr, 0, 0}, int sync_threads = 0) {
  // action -> fcs/*-cmd-norm (jsbsim_gym.py:216-222): float32 -> double widening, no clipping
  Cmd<R> cmd = {(R)act[0], (R)act[1], (R)act[2], (R)act[3]};
  es.step += 1;
  // 4 FDM frames (jsbsim_gym.py:225-232); tanks stay at 1000 lb and gear at 0 by construction
  FrameObs<R> fo;
  FrameCfg cfg = {kDt, 0.0, MS_FLIGHT};
  constexpr bool GROUND = GMODE == GROUND_FULL;
  // The reference-detail instantiations (ground reactions on) also carry the carry-over reset (auto_reset == 2): an env
  // that finishes keeps its state, gets its kinematic state re-initialised and runs run_ic()'s two zero-dt frames
  // as iterations 4 and 5 of this very loop (veh_carryover_reset spelled out; the frame takes dt at run time there).
  constexpr bool CARRY = GMODE != GROUND_OFF;
  GroundMem gm;      // friction multipliers of the contacts (GROUND_FULL only; contact is confined to one env-step)
  if (GROUND) gm.started = 0;
  int first_touch = 4;
  int flags = 0;
  // Two passes over ONE copy of the frame loop: the four frames of the step, then - carry-over reset only - run_ic()'s
  // two zero-dt frames. The epilogue sits between the passes, outside the inner loop, so the loop body the warps
  // spend their time in stays as small as in the ground-less instantiation (instruction cache).
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
  for (int pass = 0; pass < (CARRY ? 2 : 1); ++pass) {
    const int k_end = pass == 0 ? 4 : 2;
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (int k = 0; k < k_end; ++k) {
#ifdef __CUDA_ARCH__
#ifndef F16_PREFETCH_AT_FRAME
#define F16_PREFETCH_AT_FRAME 3
#endif
      if (pass == 0 && k == F16_PREFETCH_AT_FRAME) {
        if (pf.count < 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pf.ptr), "r"(pf.stride) : "memory");
        for (int i = 0; i < pf.count; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf.ptr + (size_t)i * pf.stride));
      }
#endif
#ifdef __CUDA_ARCH__
      // every warp of the CTA that steps a tile starts the frame together (shared instruction-cache fills): a named
      // barrier over `sync_threads` threads - the caller guarantees that exactly that many run these four iterations
      if (FRAME_SYNC && pass == 0) asm volatile("bar.sync 1, %0;" ::"r"(sync_threads) : "memory");
#endif
      // the first flight frame after a fresh construct still carries the mass properties of the 1500-lb tanks' CG
      const bool first = es.step == 1 && k == 0 && pass == 0 && !(es.episodes & kEpisodeUsedBit);
      fdm_frame<R, false, GMODE != GROUND_OFF, CARRY>(s, T, msets, cfg, cmd, first, fo);
      if (pass == 0) {
        if (GROUND) {
          if (fo.may_touch) ground_fix<R>(s, fo, msets_d[first ? MS_FLIGHT_FIRST : MS_FLIGHT], gm);
          else if (gm.started) { for (int i = 0; i < 3 * kNumStructure; ++i) gm.lm[i] = 0.0; }   // FGLGear: not compressed
        } else if (GMODE == GROUND_DETECT) {
          if (fo.may_touch && first_touch == 4) first_touch = k;
        }
      }
    }
    if (!CARRY) break;
    if (pass == 1) {
      for (int i = 0; i < 3; ++i) { s.vi1[i] = (R)s.vi[i]; s.vi2[i] = (R)s.vi[i]; }   // InitializeDerivatives
      s.n2 = (R)(f16data::idlen2 + 1.0 * (f16data::maxn2 - f16data::idlen2));       // InitRunning + GetSteadyState
      s.aug = R(0);
      break;
    }
    flags = env_step_epilogue<R>(s, es, fo, snapshot, snapshot_props, seed, env_id, auto_reset, frame16, tframe16, reward_out,
                                 ep_ret_out, ep_len_out);
    // (with the carry-over reset nothing is discarded: a last-frame contact's accelerations feed the next episode)
    if (GMODE == GROUND_DETECT && first_touch < 4 && !(first_touch == 3 && (flags & STEP_RESET) && auto_reset != 2))
      return STEP_ACTIVE | STEP_NEAR_GROUND;
    if (!(auto_reset == 2 && (flags & STEP_RESET))) break;
    Veh<R> ic;
    veh_from_packed(ic, snapshot);
    for (int i = 0; i < 4; ++i) s.q[i] = ic.q[i];                     // FGPropagate::SetInitialState
    for (int i = 0; i < 3; ++i) { s.ri[i] = ic.ri[i]; s.vi[i] = ic.vi[i]; s.wi[i] = ic.wi[i]; }
    s.epa = ic.epa;
    cfg.dt = 0.0;
  }
  if (!CARRY)
    return env_step_epilogue<R>(s, es, fo, snapshot, snapshot_props, seed, env_id, auto_reset, frame16, tframe16, reward_out, ep_ret_out,
                                ep_len_out);
  return flags;
}

}  // namespace f16
