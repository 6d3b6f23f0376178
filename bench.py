#!/usr/bin/env python3
"""bench.py - F-16 env-steps/s on N B200s (one process per GPU) next to the CPU reference arm.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm (oracle port, host threads)

One "step" = one env-step (4 FDM frames + observation stack + reward/termination, auto-reset on) of
every env of the job. Workload (BASELINE.json configs[3]): 1 048 576 F-16 envs per GPU, FP32
throughput mode, stacked (10,15) observations and goal reward, random actions. Per-GPU work is fixed
as N grows (weak scaling, no data-path collective; only the rollout statistics are all-reduced).
The per-GPU working set (state + observations, ~0.9 GB) is far larger than the 126 MB L2, so every
step streams from HBM without an explicit flush.

`value`  : device-resident throughput - actions already in HBM, CUDA-event timing, max over ranks.
`e2e`    : the same metric through the public SB3 VecEnv call (NumPy actions in, NumPy obs/reward/done
           out, pinned host buffers, copies and sync inside the timed region).
`roofline`: achieved algorithmic bytes/s of the step kernel vs the measured HBM copy bandwidth.
`cpu_baseline`: the oracle (a C++ port of the reference path, not JSBSim itself) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = 1_048_576
METRIC = "F-16 env-steps/sec"
UNIT = "env-steps/s"
WORKLOAD = ("BASELINE configs[3]: 1M F-16 envs per GPU (aircraft/f16, 4 FDM frames per env-step), FP32 mode, "
            "stacked (10,15) observations + goal reward, random actions, auto-reset")
# algorithmic bytes per env-step, FP32 mode with the materialised stack (DESIGN.md "Roofline"):
# state read+write 2*(11*8 + 42*4 + 8*4) = 576, action 16, previous 9 frames read 540, stack write 600,
# reward 4 + done 1 + truncated 1
BYTES_PER_ENV_STEP_FP32 = 576 + 16 + 540 + 600 + 6
BYTES_PER_ENV_STEP_FP64 = 2 * (11 * 8 + 42 * 8 + 8 * 4) + 16 + 540 + 600 + 6


def ncu_traffic(kernel, envs):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kernel)
        if t and int(t["envs"]) == int(envs):
            return float(t["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(seconds=12.0, threads=None):
    """The oracle's env (CPU restatement of the reference path) on host threads, bounded sample."""
    from oracle import f16_oracle
    threads = threads or (os.cpu_count() or 1)
    f16_oracle.rollout(threads, 50, 1, threads)                      # warm-up / page-in
    t0 = time.perf_counter()
    n, _ = f16_oracle.rollout(threads * 4, 250, 2, threads)
    rate = n / (time.perf_counter() - t0)
    per_thread_steps = max(250, int(rate * seconds / (threads * 4)))
    t0 = time.perf_counter()
    n, cs = f16_oracle.rollout(threads * 4, per_thread_steps, 3, threads)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d envs x %d env-steps, random actions, auto-reset, %d host threads, %.1f s (oracle/libf16oracle.so: C++ restatement of the reference path, not JSBSim itself)"
                      % (threads * 4, per_thread_steps, threads, dt)}


def run_reference(args):
    """--impl reference: the reference path's CPU implementation (oracle port; real JSBSim is not
    installable offline) on all host threads; a 'step' is a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import f16_oracle
    threads = os.cpu_count() or 1
    envs = threads * 4
    steps_per_sample = 4000          # ~0.1-0.3 s of host work per 'step'
    for _ in range(max(1, args.warmup)):
        f16_oracle.rollout(envs, 50, 0, threads)
    t0 = time.perf_counter()
    total = 0
    for k in range(args.steps):
        n, _ = f16_oracle.rollout(envs, steps_per_sample, k + 1, threads)
        total += n
    dt = time.perf_counter() - t0
    value = total / dt
    sample = "%d samples of %d envs x %d env-steps on %d host threads (oracle port of the reference path)" % (args.steps, envs, steps_per_sample, threads)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "reference_arm": "CPU restatement of JSBSim F-16 env-step (jsbsim PyPI package unavailable offline)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_ours(args):
    import numpy as np
    import torch

    from f16_jsb_b200 import F16BatchedEnv, F16VecEnv
    from f16_jsb_b200.distributed import allreduce_stats, barrier, init_from_env, max_over_ranks, shard_range

    rank, local_rank, world = init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    n_env = args.envs
    total_envs = n_env * world
    lo, hi = shard_range(total_envs, rank, world)
    mode = args.mode
    env = F16BatchedEnv(hi - lo, device=dev, mode=mode, seed=args.seed, env_id_base=lo,
                        ground_reactions={"default": None, "on": True, "off": False}[args.ground])
    ground_main = env.ground_reactions
    env.reset()
    # actions resident in HBM before the timed region: a ring of distinct uniform action batches
    ring = 8
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    lo_a = torch.tensor([-1, -1, -1, 0], dtype=torch.float32, device=dev)
    hi_a = torch.tensor([1, 1, 1, 1], dtype=torch.float32, device=dev)
    actions = [lo_a + (hi_a - lo_a) * torch.rand((hi - lo, 4), generator=gen, device=dev) for _ in range(ring)]
    for w in range(max(3, args.warmup)):
        env.step(actions[w % ring], auto_reset=True)
    torch.cuda.synchronize(dev)
    env.stats(reset=True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    torch.cuda.synchronize(dev)
    if sampler:
        sampler.start()
    launches0 = env.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for k in range(args.steps):
        env.step(actions[k % ring], auto_reset=True)
    ev1.record()
    torch.cuda.synchronize(dev)
    barrier()
    launches = env.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    ms_total = max_over_ranks(ev0.elapsed_time(ev1), device=dev)
    ms_step = ms_total / args.steps
    value = total_envs * args.steps / (ms_total * 1e-3)
    st = env.stats()
    stats_t = torch.tensor([st[k] for k in ("episodes", "return_sum", "length_sum", "crashes", "goals", "truncations", "env_steps", "ground_redos")],
                           dtype=torch.float64, device=dev)
    allreduce_stats(stats_t)                        # the only collective: rollout statistics over NVLink
    stats = stats_t.cpu().tolist()

    # ---- the same rollout, continued with the other setting of the ground reactions (include/f16_b200.h)
    from f16_jsb_b200 import _lib as _f16lib
    _f16lib.check(env.lib.f16_set_ground_reactions(env._h, 0 if ground_main else 1), "f16_set_ground_reactions")
    for w in range(3):
        env.step(actions[w % ring], auto_reset=True)
    g_steps = max(10, args.steps // 4)
    barrier()
    torch.cuda.synchronize(dev)
    ev0.record()
    for k in range(g_steps):
        env.step(actions[k % ring], auto_reset=True)
    ev1.record()
    torch.cuda.synchronize(dev)
    ground_other_ms = max_over_ranks(ev0.elapsed_time(ev1), device=dev) / g_steps

    # ---- and with every reference detail on: ground reactions + the carry-over reset (F16_AUTO_RESET_CARRYOVER:
    # a finished env restarts as the reference's run_ic() leaves a used env object, include/f16_b200.h)
    import ctypes as _C
    _f16lib.check(env.lib.f16_set_ground_reactions(env._h, 1), "f16_set_ground_reactions")

    def step_carryover(a):
        _f16lib.check(env.lib.f16_step(env._h, _C.c_void_p(a.data_ptr()), 2, env._stream()), "f16_step")

    for w in range(3):
        step_carryover(actions[w % ring])
    barrier()
    torch.cuda.synchronize(dev)
    ev0.record()
    for k in range(g_steps):
        step_carryover(actions[k % ring])
    ev1.record()
    torch.cuda.synchronize(dev)
    carry_ms = max_over_ranks(ev0.elapsed_time(ev1), device=dev) / g_steps

    # ---- end to end through the public VecEnv API with host buffers
    env.close()
    del env
    torch.cuda.empty_cache()
    e2e_steps = max(3, min(args.steps, args.e2e_steps))

    phases, done_rate, numa = {}, {}, [-1]

    def e2e_run(host_obs, rings, dma_both=False):
        """F16VecEnv.step with actions in pinned host memory and NumPy results out, copies and sync inside."""
        venv = F16VecEnv(hi - lo, device=dev, mode=mode, seed=args.seed, env_id_base=lo, host_obs=host_obs, host_rings=rings,
                         host_dma_both=dma_both)
        venv.reset()
        rng = np.random.default_rng(99 + rank)
        bufs = [venv.action_buffer(), venv.action_buffer()]
        for b in bufs:
            b[...] = rng.uniform([-1, -1, -1, 0], [1, 1, 1, 1], size=(hi - lo, 4)).astype(np.float32)
        # warm-up through the same call, long enough for episodes to be ending all along the timed steps
        # (random actions: the first crashes come after ~300 env-steps), so the timed region pays for the
        # terminal observations and reset fix-ups of finished envs too
        for w in range(args.e2e_warmup if host_obs == "window" else min(args.e2e_warmup, 50)):
            venv.step(bufs[w % 2])
        if venv._win is not None:
            venv._win.timing(reset=True)
        barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        finished = 0
        for k in range(e2e_steps):
            obs, rew, dones, infos = venv.step(bufs[k % 2])
            finished += int(np.count_nonzero(dones))
        torch.cuda.synchronize(dev)
        barrier()
        secs = max_over_ranks(time.perf_counter() - t0, device=dev)
        assert obs.shape == (hi - lo, 10, 15) and rew.shape == (hi - lo,) and dones.shape == (hi - lo,)
        if venv._win is not None:
            numa[0] = venv._win.numa_node
            phases[(host_obs, rings, dma_both)] = {k: round(v, 4) for k, v in venv._win.timing().items() if k != "carry_over_duration" or v}
        venv.close()
        del venv
        torch.cuda.empty_cache()
        done_rate[(host_obs, rings, dma_both)] = finished / float(e2e_steps)
        return total_envs * e2e_steps / secs

    e2e_value = e2e_run("window", 2)
    e2e_other = ({"window_1ring": e2e_run("window", 1), "window_2rings_dma_both": e2e_run("window", 2, True),
                  "copy_whole_stacks": e2e_run("copy", 2)} if args.e2e_variants else {})

    if rank != 0:
        return
    peak, peak_src = measured_peaks()
    bpe = BYTES_PER_ENV_STEP_FP32 if mode == "fp32" else BYTES_PER_ENV_STEP_FP64
    achieved = (hi - lo) * bpe / (ms_step * 1e-3) / 1e9          # per GPU: one launch processes one rank's envs
    cpu = cpu_baseline(seconds=args.cpu_seconds) if (not args.no_cpu_baseline and world == 1) else None   # rank 0, N=1 only
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if mode == "fp32" else "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_gpu": n_env, "total_envs": total_envs, "mode": mode,
                   "frames_per_env_step": 4,
                   "actions": "uniform over the action Box, 8 distinct (N,4) batches resident in HBM before the timed region and used in turn (the 16 B per env-step of the roofline's action read); in-kernel Philox actions (f16_step(actions=NULL)) give the same rate", "fdm_frames_per_s": value * 4, "l2": "inputs larger than L2 (state+obs %.0f MB per GPU); no flush" % ((hi - lo) * (bpe - 22) / 2e6),
                   "parallelism": "env-sharded x%d, no data-path collective" % world,
                   "ground_reactions": {"timed": "on" if ground_main else "off",
                                        "note": "default of the mode (on in fp64, off in fp32); they only act inside the last env-step of a crash",
                                        "other_setting_ms_per_step": ground_other_ms,
                                        "other_setting_value": total_envs / (ground_other_ms * 1e-3)},
                   "reset": {"timed": "snapshot (finished envs restart from the state of a fresh reference env object)",
                             "carryover_with_ground_reactions_ms_per_step": carry_ms,
                             "carryover_with_ground_reactions_value": total_envs / (carry_ms * 1e-3)},
                   "rollout_stats": {"episodes": stats[0], "mean_return": (stats[1] / stats[0]) if stats[0] else None,
                                     "mean_length": (stats[2] / stats[0]) if stats[0] else None, "crashes": stats[3], "goals": stats[4],
                                     "truncations": stats[5], "ground_redos": stats[7],
                                     "note": "episodes that ENDED inside the timed steps: with all envs reset together shortly before, long episodes are under-represented (an unbiased 6 000-step run gives a mean length of ~855, tools/soak.py)"}},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": total_envs * 16, "d2h_bytes_per_step": total_envs * (60 + 4 + 1 + 1),
                "steps": e2e_steps,
                "api": "F16VecEnv.step(actions in pinned host memory) -> NumPy obs (N,10,15), rewards, dones, infos; host-resident "
                       "observation windows (two rings): only the newest frame of every env crosses PCIe (60 B instead of the 600 B "
                       "stack), host threads carry it over to the second ring, the step is pipelined in four pieces (upload | kernel | "
                       "download); finished envs' records (144 B each, < 1 % of the envs per step) come through mapped host memory "
                       "and are not counted",
                "variants": e2e_other, "host_ms_per_step_by_phase": phases.get(("window", 2, False)),
                "numa_node_rank0": numa[0], "warmup_steps": args.e2e_warmup, "episodes_finished_per_step": done_rate.get(("window", 2, False))},
        "gpu_launches": int(launches) * world,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic("f16_step_kernel<%s>" % ("float" if mode == "fp32" else "double"), hi - lo),
                     "traffic_unit": "DRAM bytes per launch (ncu --set full, profiles/)", "algorithmic_bytes_per_launch": (hi - lo) * bpe,
                     "peak_source": peak_src, "kernel": "f16_step_kernel<%s>" % ("float" if mode == "fp32" else "double"),
                     "algorithmic_bytes_per_env_step": bpe, "per": "GPU"},
    }
    if cpu:
        out["cpu_baseline"] = cpu
    print(json.dumps(out))


def main():
    # exactly one line on stdout: whatever libraries print there (NCCL's version banner, ...) is sent to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="envs per GPU")
    ap.add_argument("--mode", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--ground", default="default", choices=["default", "on", "off"],
                    help="ground reactions of the device-resident run (default: on in fp64, off in fp32)")
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--e2e-warmup", type=int, default=600)
    ap.add_argument("--no-e2e-variants", dest="e2e_variants", action="store_false",
                    help="skip the one-ring and whole-stack-copy variants of the end-to-end measurement")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
