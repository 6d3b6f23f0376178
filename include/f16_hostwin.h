/* f16_hostwin.h - C ABI of the host-resident observation windows (libf16b200.so).
 *
 * The reference returns, from every env.step, the whole (10,15) float32 stack of the env
 * (jsbsim_gym/jsbsim_gym.py:150,235,263: deque(maxlen=10) -> np.array), and SB3's DummyVecEnv hands the
 * (N,10,15) array to the algorithm (stable_baselines3/common/vec_env/dummy_vec_env.py:56-73). Nine of the ten
 * rows were already returned by the previous step, so a batched env whose state lives on a GPU only has to
 * move the NEWEST frame of every env across PCIe: 60 B per env-step instead of 600 B.
 *
 * A window ring is 11 slots of pinned host memory, slot-major: ring[slot][env][15] with a slot pitch that is
 * a multiple of the page size. Each step the newest frames of all envs are DMA-ed - one contiguous copy -
 * into the next slot. The ring's pages are mapped twice, back to back, in virtual memory (memfd + two
 * MAP_FIXED mappings), so the ten slots of the current window are always contiguous in address space and
 * the stacked observation of all envs is a zero-copy strided array
 *     obs[n][k][f] = *(base + (first_slot + k) * slot_pitch + n * 60 + f * 4),  first_slot in 0..10
 * (row 0 oldest, row 9 newest, as the reference's). If the double mapping cannot be pinned for DMA the ring
 * falls back to 22 separately allocated slots and every write is mirrored into slot + 11.
 *
 * An env that finished (f16_done_record, f16_b200.h) gets its terminal stack gathered out of the ring
 * (info["terminal_observation"], dummy_vec_env.py:68) and its nine older slots overwritten with the reset
 * frame, as the reference's reset fills the deque with ten copies (jsbsim_gym.py:325-329). Those fix-ups run on
 * host threads while the frame DMA is still in flight.
 *
 * n_rings = 2 alternates two rings so that the array returned by step t is not written again before step
 * t + 2 (SB3 reads `_last_obs` after the next env.step, stable_baselines3/common/on_policy_algorithm.py:247).
 * The frames still cross PCIe once: they land in the ring being returned and a few host threads copy that slot
 * (streaming stores) into the other ring in the background, before the next step returns that one
 * (F16_HOSTWIN_DMA_BOTH sends them twice instead). n_rings = 1 keeps an array valid only until the next step.
 *
 * A step of a large batch is pipelined in pieces over several streams (f16_step_range): the upload of one
 * piece's actions, the kernel of the previous piece and the download of the one before run concurrently.
 *
 * Environment variables read by f16_hostwin_create (tuning; the defaults are the measured best on the B200 boxes):
 * F16_HOSTWIN_CHUNKS (1..8 pieces per step; default 4 from 524 288 envs, 2 from 131 072, else 1),
 * F16_HOSTWIN_THREADS (1..16 worker threads per pool; default up to 4), F16_HOSTWIN_NUMA (0 off, 1 place buffers
 * and worker threads on the GPU's NUMA node, 2 also leave the calling thread there; default 1), F16_HOSTWIN_ZEROCOPY
 * (1: steps of up to 32 768 envs have the kernel store its frames straight into the ring, which is mapped into the
 * GPU's address space, instead of copying them afterwards; 2: every step; 0: never; default 1).
 */
#ifndef F16_HOSTWIN_H
#define F16_HOSTWIN_H

#include <stdint.h>

#include "f16_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct f16_hostwin* f16_hostwin_handle;

enum { F16_HOSTWIN_SLOTS = 11 };
enum { F16_HOSTWIN_PIN = 1,        /* pin the rings for CUDA DMA (needs a CUDA device) */
       F16_HOSTWIN_NO_ALIAS = 2,   /* skip the double mapping, use the mirrored 22-slot ring */
       F16_HOSTWIN_DMA_BOTH = 4 }; /* two rings: DMA every frame into both rings (2 x 60 B per env-step over PCIe).
                                      Default: DMA into the returned ring only and let host threads carry the slot over
                                      to the other ring before the next step returns that one - measured on the B200 box
                                      with episodes ending all along: 2.1 ms against 3.1 ms per step of 1M envs */
enum { F16_HOSTWIN_MAX_CHUNKS = 8 };

typedef struct f16_hostwin_result {
  int32_t ring;                    /* which ring holds this step's window */
  int32_t first_slot;              /* first slot of the window in that ring's (double) mapping, 0..10 */
  int64_t n_done;                  /* envs that finished in this step */
  const float* reward;             /* N float    (valid until the step after next) */
  const uint8_t* done;             /* N uint8                                      */
  const uint8_t* truncated;        /* N uint8                                      */
  const f16_done_record* records;  /* n_done records (valid until the next step)   */
  const float* terminal_obs;       /* n_done x 10 x 15 float, in record order (valid until the step after next) */
} f16_hostwin_result;

/* Rings for n_envs environments; n_rings is 1 or 2. Without F16_HOSTWIN_PIN nothing touches CUDA (the
 * host-only entry points below still work; used by the CPU tests). */
int f16_hostwin_create(f16_hostwin_handle* out, int64_t n_envs, int n_rings, int flags);
/* Contiguous (N,10,15) copy of the window (ring, first_slot) into caller memory, on the window's worker threads. */
int f16_hostwin_gather(f16_hostwin_handle w, int ring, int first_slot, float* dst);
/* The env's done list points into this window's memory: detach the env before destroying the window (or destroy the
 * env first). */
int f16_hostwin_detach(f16_hostwin_handle w, f16_handle env);
int f16_hostwin_destroy(f16_hostwin_handle w);

/* Geometry of one ring for building the strided view: base address, slot pitch in bytes, number of slots
 * addressable from base (22), and whether the second half aliases the first (1) or is a mirror copy (0). */
int f16_hostwin_layout(f16_hostwin_handle w, int ring, float** base, int64_t* slot_pitch_bytes, int32_t* n_slots, int32_t* aliased);

/* Pinned N x 4 float staging buffers for actions (which = 0 or 1): filling one of them and passing it to
 * f16_hostwin_step makes the host->device copy a plain DMA. */
float* f16_hostwin_action_buffer(f16_hostwin_handle w, int which);

/* After f16_reset on a frame-layout env: copy every env's reset frame into all slots of all rings
 * (jsbsim_gym.py:325-329) and report the window. Synchronises `stream`. */
int f16_hostwin_reset(f16_hostwin_handle w, f16_handle env, void* stream, f16_hostwin_result* out);

/* One env-step through host buffers: actions (N x 4 float, host) -> device, f16_step, newest frames ->
 * the next slot of the ring(s), reward / done / truncated -> host, fix-ups for finished envs.
 * Synchronises `stream` before returning. */
int f16_hostwin_step(f16_hostwin_handle w, f16_handle env, const float* actions_host, int auto_reset, void* stream,
                     f16_hostwin_result* out);

/* Average wall time per f16_hostwin_step since the last reset of the counters, by phase, in seconds:
 * [0] enqueue (copies, kernel launch), [1] wait for upload + kernel, [2] fix-ups of the older slots (under the
 * frame DMA), [3] wait for the device->host copies, [4] wait for the previous carry-over, [5] fix-ups of slot
 * head-1, [6] hand-off to the copier threads, [7] duration of the carry-over on the copier threads (overlaps the
 * next step; not part of its wall time). */
enum { F16_HOSTWIN_PHASES = 8 };
int f16_hostwin_timing(f16_hostwin_handle w, double* seconds_per_step, int reset);

/* NUMA node the pinned buffers were first touched on and the worker threads are kept on: the node of the current
 * GPU's PCIe root (/sys/bus/pci/devices/<bus id>/numa_node), so that the frame DMA, the fix-ups and the carry-over of
 * one rank stay on one socket. -1: no placement (single-node host, no information, F16_HOSTWIN_NUMA=0, or none of
 * that node's CPUs are available to the process). */
int f16_hostwin_numa_node(f16_hostwin_handle w);

/* The same two operations for a producer that already has the data in host memory (tests, replay):
 * frames N x 15, reward N, done N, truncated N, records[n_done]. No CUDA calls. */
int f16_hostwin_fill(f16_hostwin_handle w, const float* frames, f16_hostwin_result* out);
int f16_hostwin_push(f16_hostwin_handle w, const float* frames, const float* reward, const uint8_t* done, const uint8_t* truncated,
                     const f16_done_record* records, int64_t n_done, f16_hostwin_result* out);

#ifdef __cplusplus
}
#endif
#endif /* F16_HOSTWIN_H */
