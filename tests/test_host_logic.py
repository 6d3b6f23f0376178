"""Host-side logic that needs no GPU: constants and spaces mirror the reference, goal sampling,
sharding, lazy infos, and the world_size-2 statistics all-reduce over gloo."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_constants_match_reference_values():
    from f16_jsb_b200 import constants as c
    assert c.STATE_FORMAT[0] == "position/lat-gc-rad" and c.STATE_FORMAT[-1] == "attitude/psi-rad" and len(c.STATE_FORMAT) == 12
    assert c.RADIUS == 6.3781e6 and c.NUM_STACKED_FRAMES == 10 and c.MAX_EPISODE_STEPS == 1200 and c.DOWN_SAMPLE == 4
    assert c.SINGLE_OBS_LOW.shape == (15,) and c.SINGLE_OBS_LOW[3] == 0 and c.SINGLE_OBS_LOW[14] == 0
    assert np.isclose(c.SINGLE_OBS_HIGH[10], np.pi / 2 + 1e-5)
    assert np.array_equal(c.ACTION_LOW, [-1, -1, -1, 0]) and np.array_equal(c.ACTION_HIGH, [1, 1, 1, 1])


def test_spaces_shape_and_dtype():
    from f16_jsb_b200.vec_env import make_spaces
    o, a = make_spaces()
    assert o.shape == (10, 15) and o.dtype == np.float32 and a.shape == (4,) and a.dtype == np.float32
    s = a.sample()
    assert a.contains(s) and o.contains(np.zeros((10, 15), np.float32) + np.float32(0.5))


def test_angle_helpers():
    from f16_jsb_b200.constants import normalize_angle_0_2pi, normalize_angle_mpi_pi
    assert normalize_angle_mpi_pi(np.pi) == pytest.approx(-np.pi)
    assert normalize_angle_mpi_pi(-np.pi) == pytest.approx(-np.pi)
    assert normalize_angle_mpi_pi(3 * np.pi / 2) == pytest.approx(-np.pi / 2)
    assert normalize_angle_mpi_pi(float("nan")) == 0.0 and normalize_angle_mpi_pi(float("inf")) == 0.0
    assert normalize_angle_0_2pi(-0.5) == pytest.approx(2 * np.pi - 0.5)


def test_goal_sampling_matches_reference_draw_order(golden):
    from f16_jsb_b200.constants import sample_goal_numpy
    assert np.array_equal(sample_goal_numpy(0), golden["random0"]["goal"])
    g = sample_goal_numpy(12345)
    assert g.dtype == np.float32 and 1000.0 <= np.hypot(g[0], g[1]) < 10000.0 and 1000.0 <= g[2] < 4000.0


def test_shard_range_partitions_exactly():
    from f16_jsb_b200.distributed import shard_range
    for total, world in ((1_000_000, 8), (65_536, 4), (10, 3), (7, 8)):
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    assert shard_range(1_048_576, 3, 8) == (393216, 524288)


def test_lazy_infos_behave_like_a_list_of_dicts():
    from f16_jsb_b200.vec_env import _LazyInfos
    infos = _LazyInfos(5, {2: {"TimeLimit.truncated": True, "terminal_observation": np.zeros((10, 15)), "episode": {"r": 1.0, "l": 3, "t": 0.1}}})
    assert len(infos) == 5 and infos[0].get("terminal_observation") is None and not infos[4]["TimeLimit.truncated"]
    assert infos[2]["TimeLimit.truncated"] and infos[-3]["episode"]["l"] == 3
    assert sum(1 for i in infos if i.get("episode")) == 1
    with pytest.raises(IndexError):
        infos[5]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_envs, out_q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from f16_jsb_b200.distributed import allreduce_stats, barrier, init_from_env, max_over_ranks, shard_range
    r, lr, w = init_from_env(backend="gloo")
    lo, hi = shard_range(total_envs, r, w)
    # per-rank rollout statistics as the step kernel would accumulate them for its shard
    stats = torch.zeros(8, dtype=torch.float64)
    stats[0] = hi - lo            # episodes
    stats[1] = float(sum(range(lo, hi)))   # sum of "returns" = sum of global env ids
    stats[6] = 10.0 * (hi - lo)   # env steps
    allreduce_stats(stats)
    tmax = max_over_ranks(1.0 + r)
    barrier()
    out_q.put((r, lo, hi, stats.tolist(), tmax))
    dist.destroy_process_group()


def test_stats_allreduce_world_size_2_gloo():
    world, total = 2, 1001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1:3] == (0, 501) and res[1][1:3] == (501, 1001)
    for r in res:
        assert r[3][0] == total and r[3][1] == float(sum(range(total))) and r[3][6] == 10.0 * total
        assert r[4] == 2.0


def test_tensor_core_linear_shape_support_is_host_logic():
    """f16_lma_linear_supported (include/f16_lma.h) is pure host code: which Linear shapes of the reference's LMA policy
    (jsbsim_gym/LMA_features.py:221-279,315-385; SB3 MlpExtractor heads) the tensor-core kernel builds, and that the
    launcher refuses bad arguments with a message before touching a device."""
    import ctypes as C

    from f16_jsb_b200 import _lib
    L = _lib.load()
    built = [(17, 64), (128, 32), (32, 96), (32, 32), (32, 128), (160, 64), (64, 64), (128, 64),    # forward
             (96, 32), (64, 160), (64, 128), (32, 128),                                             # input gradients (W^T)
             (160, 128), (128, 160), (64, 256)]                                                     # in column groups
    for k, n in built:
        assert L.f16_lma_linear_supported(k, n) == 1, (k, n)
    for k, n in [(64, 4), (64, 1), (40, 32), (32, 48), (32, 0), (0, 32), (160, 320)]:
        assert L.f16_lma_linear_supported(k, n) == 0, (k, n)
    assert L.f16_lma_linear_forward(0, 32, 32, None, None, None, None, None) != 0
    assert b"rows must be positive" in L.f16_last_error()
    assert L.f16_lma_linear_forward(128, 32, 32, None, None, None, None, None) != 0
    assert b"NULL pointer" in L.f16_last_error()
    buf = (C.c_float * 64)()
    assert L.f16_lma_linear_forward(128, 64, 4, buf, buf, None, buf, None) != 0
    assert b"unsupported shape" in L.f16_last_error()


def test_tensor_core_wgrad_shape_support_is_host_logic():
    """f16_lma_linear_wgrad_tc_supported (include/f16_lma.h): the weight-gradient shapes of the reference's policy the
    tensor-core kernel builds (its shared-memory planner must find a chunk / ring configuration), and the ones it leaves to
    the FP32 slab kernel."""
    from f16_jsb_b200 import _lib
    L = _lib.load()
    for k, n in [(128, 32), (32, 96), (32, 32), (32, 128), (160, 64), (64, 64), (128, 64), (96, 32)]:
        assert L.f16_lma_linear_wgrad_tc_supported(k, n) == 1, (k, n)
    for k, n in [(17, 64), (64, 4), (64, 1), (160, 128), (128, 128), (32, 160), (192, 32), (0, 32), (32, 0)]:
        assert L.f16_lma_linear_wgrad_tc_supported(k, n) == 0, (k, n)
    assert L.f16_lma_linear_wgrad_tc(0, 32, 32, None, None, None, None, None) != 0
    assert b"rows must be positive" in L.f16_last_error()
    assert L.f16_lma_linear_wgrad_tc(64, 32, 32, None, None, None, None, None) != 0
    assert b"NULL pointer" in L.f16_last_error()
