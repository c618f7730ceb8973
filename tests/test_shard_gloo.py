"""Multi-rank host logic on CPU: world_size-2 gloo processes shard a frame sequence, run the per-frame front-end on
their block (the CPU oracle stands in for the GPU step here) and all-gather the descriptor records."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from orb_slam2_ros2_b200 import shard


def test_frame_ranges_partition_the_sequence():
    for F in (0, 1, 7, 8, 4541, 4544):
        for R in (1, 2, 4, 8):
            blocks = [shard.frame_range(F, r, R) for r in range(R)]
            flat = [f for b in blocks for f in b]
            assert flat == list(range(F))
            assert max((len(b) for b in blocks), default=0) == shard.frames_per_rank(F, R)
    assert len(shard.frame_range(4541, 7, 8)) == 4541 - 7 * 568  # 568 frames per rank at R=8


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, ret):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle_py as O
    from orb_slam2_ros2_b200 import synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N = 20
    frames = shard.frame_range(n_frames, rank, world)
    desc = torch.zeros((len(frames), N, 32), dtype=torch.uint8)
    cnt = torch.zeros((len(frames), 1), dtype=torch.int32)
    for i, f in enumerate(frames):
        e = O.extract(synth.synth_image(120, 160, 100 + f), N, 2, 1.2)
        desc[i, : len(e.kps)] = torch.from_numpy(e.desc)
        cnt[i, 0] = len(e.kps)
    all_desc = shard.gather_sequence(desc, n_frames, rank, world)
    all_cnt = shard.gather_sequence(cnt, n_frames, rank, world)
    ret[rank] = (all_desc.numpy().copy(), all_cnt.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [5, 6])
def test_two_rank_gather_equals_single_process(oracle, n_frames):
    from orb_slam2_ros2_b200 import synth

    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_frames, ret), nprocs=world, join=True)
    N = 20
    exp_desc = np.zeros((n_frames, N, 32), np.uint8)
    exp_cnt = np.zeros((n_frames, 1), np.int32)
    for f in range(n_frames):
        e = oracle.extract(synth.synth_image(120, 160, 100 + f), N, 2, 1.2)
        exp_desc[f, : len(e.kps)] = e.desc
        exp_cnt[f, 0] = len(e.kps)
    assert exp_cnt.min() == 20
    for r in range(world):
        d, c = ret[r]
        assert np.array_equal(d, exp_desc) and np.array_equal(c, exp_cnt)
