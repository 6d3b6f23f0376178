/* f16_b200.h - C ABI of the B200-native batched F-16 environment (libf16b200.so).
 *
 * This is the drop-in boundary for the reference's env-step path. The reference reaches its
 * flight-dynamics code through the Python binding of JSBSim, one FGFDMExec per environment
 * (jsbsim_gym/jsbsim_gym.py:151-155 construct/load_model/run_ic, :168-170 ic properties,
 * :219-232 command + tank/gear properties + 4 x run(), :181-182 twelve property reads,
 * :305-306 run_ic + propulsion/set-running). One f16_step() call replaces, for N environments at
 * once, everything JSBSimEnv.step + PositionReward.step do (jsbsim_gym.py:199-287, 487-509):
 * action mapping, 4 FDM frames, observation frame, 10-frame stack, reward, termination, truncation,
 * and - with auto_reset - the DummyVecEnv reset-on-done convention
 * (stable_baselines3/common/vec_env/dummy_vec_env.py:56-73).
 *
 * Conventions: opaque handle; device buffers are owned by the caller (PyTorch in this repo) and
 * bound once; every launch goes to the caller-supplied cudaStream_t (passed as void*); no hidden
 * synchronisation except in the *_host entry points and get/set_state; int status returns
 * (0 = ok, negative = error) with f16_last_error() giving the message for the calling thread.
 * There is no CPU fallback: every entry point needs a CUDA device.
 */
#ifndef F16_B200_H
#define F16_B200_H

#include <stddef.h>
#include <stdint.h>

#include "f16_state_fields.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct f16_ctx* f16_handle;

enum { F16_MODE_FP64 = 0, /* parity mode: all model math in double            */
       F16_MODE_FP32 = 1  /* throughput mode: float math, double kinematics    */ };

enum { F16_OBS_FRAMES = 10, F16_OBS_FEATURES = 15, F16_ACTION_DIM = 4, F16_NUM_STATS = 8 };

/* Replaces jsbsim.FGFDMExec(root, None) + load_model('f16') + the constructor's run_ic()
 * (jsbsim_gym.py:151-155) for n_envs environments on CUDA device `device`. Computes the canonical
 * post-reset snapshot on the device (fresh construct -> run_ic -> reset's run_ic + set-running). */
int f16_create(f16_handle* out, int64_t n_envs, int device, int mode);
int f16_destroy(f16_handle h);

/* Bytes of device memory the caller must provide for the structure-of-arrays state. */
size_t f16_state_bytes(f16_handle h);

/* Ground reactions: JSBSim's FGGroundReactions / FGLGear for the seven STRUCTURE contacts of
 * aircraft/f16/f16.xml:137-214 plus the friction solve of FGAccelerations (the three BOGEY contacts are
 * retracted by jsbsim_gym.py:230-231). They can only act inside the last env-step of an episode that ends
 * in a crash (termination below 10 m, jsbsim_gym.py:240): about one crash in fifteen under random actions,
 * where they change the rates and the angles of attack / sideslip of the terminal frame; rewards move by
 * ~1e-5 and done flags do not change. Default: on in F16_MODE_FP64 (parity), off in F16_MODE_FP32
 * (throughput: the step kernel is then the instantiation without any ground code). */
int f16_set_ground_reactions(f16_handle h, int on);
int f16_get_ground_reactions(f16_handle h);

/* Bind caller-owned device buffers. state: f16_state_bytes() bytes, 256-byte aligned.
 * obs: N x 10 x 15 float (row 0 oldest, row 9 newest; jsbsim_gym.py:150,235,263).
 * reward: N float. done/truncated: N uint8. terminal_obs: N x 10 x 15 float or NULL.
 * ep_return / ep_len: N float / N int32, written when an episode ends (Monitor's info["episode"],
 * stable_baselines3/common/monitor.py:96-109), or NULL. */
int f16_bind(f16_handle h, void* state, float* obs, float* reward, uint8_t* done, uint8_t* truncated,
             float* terminal_obs, float* ep_return, int32_t* ep_len);

/* Ring layout of the observations (the ring buffer fused into the step kernel; F16BatchedEnv's default).
 * obs_ring: 20 x N x 15 float, slot-major. Each step writes the newest frame of every env into slot `slot` and slot
 * `slot + 10` (slot cycles 0..9) as two contiguous (N,15) planes; the chronological (10,15) stack of env n is rows
 * obs_ring[first_row + k][n][:], k = 0..9, with first_row reported by f16_obs_window - a strided view (strides N*15, 15, 1
 * floats over k, n, feature), never shifted or copied: 120 B written per env-step instead of 540 B read + 600 B written.
 * Everything else as f16_bind; terminal_obs stays a plain N x 10 x 15 tensor; f16_step_host still returns contiguous
 * N x 10 x 15 host arrays. */
int f16_bind_ring(f16_handle h, void* state, float* obs_ring, float* reward, uint8_t* done, uint8_t* truncated,
                  float* terminal_obs, float* ep_return, int32_t* ep_len);
/* First row (ring layout: first slot) of the current observation window: 0 for the stacked layout, 1..10 for the ring layout. */
int f16_obs_window(f16_handle h, int* first_row);

/* Optional frame layout of the observations: the device keeps no history. obs_frame: N x 15 float; each step
 * writes the newest frame of every env (the reset frame for an env that auto-reset in this step), 60 B per
 * env-step - exactly the bytes that have to cross PCIe when the ten-frame windows live in host memory
 * (f16_hostwin.h). Every env that finishes appends one f16_done_record to done_list (capacity N records, device
 * or mapped host memory; order unspecified) and bumps *done_count (device memory, zeroed by f16_step before each
 * launch); both may be NULL. f16_reset writes each masked env's reset frame to obs_frame. */
typedef struct f16_done_record {
  int32_t env;                 /* local env index */
  int32_t flags;               /* bit 0 truncated (TimeLimit, jsbsim_gym.py:258-261), bit 1 crash, bit 2 goal reached */
  float ep_return;             /* Monitor's info["episode"]["r"] (stable_baselines3/common/monitor.py:96-109) */
  int32_t ep_len;              /* info["episode"]["l"] */
  float terminal_frame[16];    /* newest row of info["terminal_observation"]; [15] is padding */
  float reset_frame[16];       /* the frame the next episode starts from (all ten rows of the returned obs) */
} f16_done_record;
int f16_bind_frames(f16_handle h, void* state, float* obs_frame, float* reward, uint8_t* done, uint8_t* truncated,
                    f16_done_record* done_list, int32_t* done_count);
/* Re-point the done list of a frame-layout env (double buffering by the caller). */
int f16_set_done_list(f16_handle h, f16_done_record* done_list, int32_t* done_count);

/* JSBSimEnv.reset + PositionReward.reset (jsbsim_gym.py:289-331, 511-519) for every env whose mask
 * byte is non-zero (mask == NULL: all). goals: N x 3 float device pointer holding the goal of
 * every env (entries of unmasked envs are ignored), or NULL to sample distance~U[1000,10000) m,
 * bearing~U[0,2pi), altitude~U[1000,4000) m (jsbsim_gym.py:315-317) from Philox4x32-10 keyed by
 * (seed, global env id, episode counter). */
enum { F16_AUTO_RESET_OFF = 0,
       F16_AUTO_RESET_SNAPSHOT = 1,  /* finished envs restart from the canonical fresh-env state (f16_reset)        */
       F16_AUTO_RESET_CARRYOVER = 2  /* finished envs restart as f16_reset_carryover does; needs ground reactions on */ };
int f16_reset(f16_handle h, const uint8_t* mask, const float* goals, uint64_t seed, void* stream);

/* The same call on an env object that ALREADY EXISTS, as the reference really executes it: JSBSimEnv.reset is
 * run_ic() + propulsion/set-running (jsbsim_gym.py:305-306), and JSBSim's run_ic() re-initialises only the
 * kinematic state (FGPropagate) before two frames with integration suspended. Actuator positions, PID memories,
 * the one-frame-late Auxiliary outputs the control laws read, the last accelerations and the last action
 * (fcs/ *-cmd-norm stay set; the FCS ticks twice more on it) carry over from the episode that just ended, and a
 * stepped env keeps its gear up and 1000-lb tanks. f16_reset restores the canonical state of a FRESH env object
 * instead (every episode like the first one of a new env); this entry point reproduces the second and later
 * episodes of one reference env object. The returned observation is the same either way (its twelve properties
 * depend on the initial condition only). last_actions: N x 4 float device pointer, the action of each env's last
 * step (ignored for envs that have not been stepped since they were created), or NULL for zeros. An env that was
 * never reset gets the canonical bring-up. */
int f16_reset_carryover(f16_handle h, const uint8_t* mask, const float* goals, uint64_t seed, const float* last_actions,
                        void* stream);

/* One env-step for all N envs. actions: N x 4 float device pointer
 * [roll, pitch, yaw, throttle] -> fcs/{aileron,elevator,rudder,throttle}-cmd-norm, un-clipped
 * (jsbsim_gym.py:216-222); NULL samples action_space.sample()-like uniform actions in-kernel
 * (Philox, keyed by seed/env/step). auto_reset != 0: an env that finishes is reset in the same
 * launch (new Philox goal), its stacked terminal observation goes to terminal_obs, and obs holds
 * the reset observation, as DummyVecEnv does. auto_reset is one of F16_AUTO_RESET_*: 1 restarts a finished
 * env from the canonical fresh-env state, 2 applies the reference's run_ic() + set-running to the state the
 * episode ended in (see f16_reset_carryover), with this step's action as the last action. */
int f16_step(f16_handle h, const float* actions, int auto_reset, void* stream);

/* One env-step in pieces (frame layout only), so that a caller can pipeline it over several streams: upload of
 * one piece, kernel of the previous, download of the one before. f16_step_begin advances the step counter that
 * keys in-kernel random actions and zeroes the done count on `stream`; every piece's stream must be ordered
 * after it. f16_step_range steps the envs [first, first + count), first a multiple of 32; actions is the full
 * N x 4 device array. Every env has to be covered exactly once between two f16_step_begin calls. */
int f16_step_begin(f16_handle h, void* stream);
int f16_step_range(f16_handle h, const float* actions, int auto_reset, int64_t first, int64_t count, void* stream);

/* Same step through HOST buffers (the reference-facing call: NumPy in, NumPy out). Copies
 * actions host->device, steps, copies obs/reward/done/truncated device->host and synchronises.
 * Pinned host memory makes the copies asynchronous DMA; pageable memory works but is slower.
 * Any output pointer may be NULL to skip that copy. */
int f16_step_host(f16_handle h, const float* actions_host, int auto_reset, float* obs_host,
                  float* reward_host, uint8_t* done_host, uint8_t* truncated_host, void* stream);

/* Offset the global env ids used for Philox keys (multi-GPU: rank r owns ids [base, base+N)). */
int f16_set_env_id_base(f16_handle h, int64_t base);

/* Packed double state of one env (field order: f16_state_fields.h), for parity tests and teacher
 * forcing. n must be F16_NUM_STATE_FIELDS. Synchronises. */
int f16_get_state(f16_handle h, int64_t env, double* out, int n);
int f16_set_state(f16_handle h, int64_t env, const double* in, int n);
/* Batch variants on device memory: packed is N x F16_NUM_STATE_FIELDS doubles (device pointer). */
int f16_pack_states(f16_handle h, double* packed_dev, void* stream);
int f16_unpack_states(f16_handle h, const double* packed_dev, void* stream);
/* Env-layer bookkeeping of one env: current_step (jsbsim_gym.py:215) for teacher forcing. */
int f16_set_env_step(f16_handle h, int64_t env, int32_t current_step);

/* The canonical post-reset state (packed, F16_NUM_STATE_FIELDS doubles) followed by the twelve
 * STATE_FORMAT properties (jsbsim_gym.py:12-25) read right after reset, as doubles. */
int f16_get_snapshot(f16_handle h, double* state_out, double* props12_out);

/* Rollout statistics accumulated on the device since the last call with reset != 0:
 * [0] episodes finished, [1] sum of episode returns, [2] sum of episode lengths, [3] crashes,
 * [4] goals reached, [5] truncations, [6] env-steps, [7] env-steps redone with ground-contact forces (ground
 * reactions on). Synchronises `stream`. */
int f16_get_stats(f16_handle h, double* out8, int reset, void* stream);
/* Device pointer to those 8 doubles (for an NCCL all-reduce by the caller). */
int f16_stats_device_ptr(f16_handle h, double** out);

/* Number of kernels this library has launched so far in this process (bench.py's gpu_launches). */
int64_t f16_launch_count(void);

/* Number of doubles in a packed state (= F16_NUM_STATE_FIELDS of f16_state_fields.h). */
int f16_num_state_fields(void);

const char* f16_last_error(void);
const char* f16_version(void);

#ifdef __cplusplus
}
#endif
#endif /* F16_B200_H */
