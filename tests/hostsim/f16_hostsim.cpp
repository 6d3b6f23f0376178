// f16_hostsim.cpp - TEST HARNESS ONLY. Compiles the kernel's per-env source (f16_model.cuh,
// f16_env.cuh, f16_host_setup.h) with g++ and runs it serially, one env at a time, so the kernel
// arithmetic can be compared with the oracle on a machine without a GPU (`-m "not gpu"` tests).
// It is never built into, loaded by, or reachable from the f16_jsb_b200 package: the product path
// is libf16b200.so (CUDA) and fails loudly without a device.
#include <cstdint>
#include <cstring>

#include "../../f16_jsb_b200/csrc/f16_host_setup.h"

using namespace f16;

namespace {
struct Consts {
  Tables<double> Td;
  Tables<float> Tf;
  MassSet ms[MS_COUNT];
  MassSetT<float> msf[MS_COUNT];
  double snap[F16_NUM_STATE_FIELDS + 12];
  Consts() {
    host::build_tables<double>(&Td);
    host::build_tables<float>(&Tf);
    host::build_mass_sets(ms);
    for (int i = 0; i < MS_COUNT; ++i) host::convert_mass_set(ms[i], &msf[i]);
    double ic[F16_NUM_STATE_FIELDS];
    host::initial_condition(900.0, 5000.0, ic);
    compute_snapshot(Td, ms, ic, snap);
  }
};
const Consts& consts() { static Consts c; return c; }

struct HsEnv {
  int mode;
  Veh<double> sd;
  Veh<float> sf;
  EnvScalars es;
  float obs[10][15];
  float tobs[10][15];
};
}  // namespace

extern "C" {
void hs_snapshot(double* out) { std::memcpy(out, consts().snap, sizeof(consts().snap)); }
void hs_mass_set(int which, double* out /*mass, J9, Jinv9, rp3, eye3, thr3 = 28*/) { std::memcpy(out, &consts().ms[which], sizeof(MassSet)); }

// the parity kernel's own double atan2 (f16_model.cuh), for the known-answer test against libm
double hs_atan2(double y, double x) { return datan2_fast(y, x); }

void* hs_env_create(int mode) {
  HsEnv* e = new HsEnv();
  std::memset(e, 0, sizeof(*e));
  e->mode = mode;
  return e;
}
void hs_env_destroy(void* h) { delete (HsEnv*)h; }
void hs_env_reset(void* h, const float* goal, float* obs_out) {
  HsEnv* e = (HsEnv*)h;
  const Consts& c = consts();
  float fr[16];
  e->es.episodes = (e->es.episodes & ~kEpisodeUsedBit) + 1;
  if (e->mode == 0) env_reset_one<double>(e->sd, e->es, c.snap, c.snap + F16_NUM_STATE_FIELDS, goal, fr);
  else env_reset_one<float>(e->sf, e->es, c.snap, c.snap + F16_NUM_STATE_FIELDS, goal, fr);
  for (int r = 0; r < 10; ++r) std::memcpy(e->obs[r], fr, 15 * sizeof(float));
  if (obs_out) std::memcpy(obs_out, e->obs, sizeof(e->obs));
}
// JSBSimEnv.reset on the SAME env object (f16_reset_carryover of the C ABI)
void hs_env_reset_carryover(void* h, const float* goal, const float* last_action, float* obs_out) {
  HsEnv* e = (HsEnv*)h;
  const Consts& c = consts();
  float fr[16];
  if (e->mode == 0) env_carryover_reset_one<double>(e->sd, e->es, c.Td, c.ms, c.snap, c.snap + F16_NUM_STATE_FIELDS, goal, last_action, fr);
  else env_carryover_reset_one<float>(e->sf, e->es, c.Tf, c.msf, c.snap, c.snap + F16_NUM_STATE_FIELDS, goal, last_action, fr);
  for (int r = 0; r < 10; ++r) std::memcpy(e->obs[r], fr, 15 * sizeof(float));
  if (obs_out) std::memcpy(obs_out, e->obs, sizeof(e->obs));
}
int hs_env_step(void* h, const float* action, int auto_reset, uint64_t seed, uint64_t env_id, float* obs_out, float* reward,
                float* terminal_obs_out) {
  HsEnv* e = (HsEnv*)h;
  const Consts& c = consts();
  float fr[16], tfr[16], ep_ret = 0;
  int32_t ep_len = 0;
  int flags;
  // as the kernel does: hot instantiation first; if a contact point reached the ground, redo the step from the
  // saved state with the ground-reaction instantiation
  const EnvScalars es0 = e->es;
  if (e->mode == 0) {
    const Veh<double> s0 = e->sd;
    flags = env_step_one<double, GROUND_DETECT>(e->sd, e->es, c.Td, c.ms, c.ms, c.snap, c.snap + F16_NUM_STATE_FIELDS, action, seed, env_id, auto_reset, fr, tfr, reward, &ep_ret, &ep_len);
    if (flags & STEP_NEAR_GROUND) {
      e->sd = s0; e->es = es0;
      flags = env_step_one<double, GROUND_FULL>(e->sd, e->es, c.Td, c.ms, c.ms, c.snap, c.snap + F16_NUM_STATE_FIELDS, action, seed, env_id, auto_reset, fr, tfr, reward, &ep_ret, &ep_len);
    }
  } else {
    const Veh<float> s0 = e->sf;
    flags = env_step_one<float, GROUND_DETECT>(e->sf, e->es, c.Tf, c.msf, c.ms, c.snap, c.snap + F16_NUM_STATE_FIELDS, action, seed, env_id, auto_reset, fr, tfr, reward, &ep_ret, &ep_len);
    if (flags & STEP_NEAR_GROUND) {
      e->sf = s0; e->es = es0;
      flags = env_step_one<float, GROUND_FULL>(e->sf, e->es, c.Tf, c.msf, c.ms, c.snap, c.snap + F16_NUM_STATE_FIELDS, action, seed, env_id, auto_reset, fr, tfr, reward, &ep_ret, &ep_len);
    }
  }
  // (auto_reset == 2: env_step_one itself runs the two zero-dt frames of the carry-over reset)
  // same stack update as warp_write_obs
  if (flags & STEP_TERMINAL) {
    std::memcpy(e->tobs[0], e->obs[1], 9 * 15 * sizeof(float));
    std::memcpy(e->tobs[9], tfr, 15 * sizeof(float));
    if (terminal_obs_out) std::memcpy(terminal_obs_out, e->tobs, sizeof(e->tobs));
  }
  if (flags & STEP_RESET) {
    for (int r = 0; r < 10; ++r) std::memcpy(e->obs[r], fr, 15 * sizeof(float));
  } else {
    std::memmove(e->obs[0], e->obs[1], 9 * 15 * sizeof(float));
    std::memcpy(e->obs[9], fr, 15 * sizeof(float));
  }
  if (obs_out) std::memcpy(obs_out, e->obs, sizeof(e->obs));
  return flags;
}
void hs_env_get_state(void* h, double* out) {
  HsEnv* e = (HsEnv*)h;
  if (e->mode == 0) veh_to_packed(e->sd, out); else veh_to_packed(e->sf, out);
}
void hs_env_set_state(void* h, const double* in, int current_step) {
  HsEnv* e = (HsEnv*)h;
  if (e->mode == 0) veh_from_packed(e->sd, in); else veh_from_packed(e->sf, in);
  e->es.step = current_step;
}
}
