"""torchrun worker of tests/test_gpu_sequence.py: two processes, one GPU each, share a 21-frame sequence (blocks 11 + 10)
through orbx_sequence_stereo with the NCCL or the CUDA-IPC peer-memory transport; each rank then recomputes the whole
sequence alone and compares the gathered arrays."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from orb_slam2_ros2_b200 import api, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")  # host-side plumbing only (id / handle exchange); the data path is the library's
    H, W, NF, NL, F = 240, 320, 300, 5, 21
    cam = api.Camera(300.0, 300.0, 160.0, 120.0, 0.1)
    lefts, rights = synth.synth_stereo_pool(H, W, F, seed0=500, disparity=9)
    ctx = api.Context(W, H, NF, NL, 1.2, camera=cam, max_batch=4, device=local)
    ids = [api.Communicator.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = api.Communicator(ctx, rank, world, ids[0], F)
    if os.environ.get("ORBX_TEST_TRANSPORT", "nccl") == "peer":
        handles = [None] * world
        dist.all_gather_object(handles, comm.ipc_handle())
        comm.open_peers(handles)
        assert comm.info()["transport"] == api.TRANSPORT_PEER
    blk = api.frame_range(F, rank, world)
    for rep in range(2):  # twice: the second pass overwrites live gathered arrays
        rec, gd, gn = ctx.sequence_stereo(lefts[blk.start:blk.stop], rights[blk.start:blk.stop], F, comm)
    alone = api.Context(W, H, NF, NL, 1.2, camera=cam, max_batch=4, device=local)
    rec1, gd1, gn1 = alone.sequence_stereo(lefts, rights)
    assert rec.tobytes() == rec1[blk.start:blk.stop].tobytes(), "records differ from the single-rank run"
    assert np.array_equal(gn, gn1) and np.array_equal(gd, gd1), "gathered descriptors differ from the single-rank run"
    assert gn.min() > 0
    dist.barrier()
    comm.close()
    ctx.close()
    alone.close()
    print(f"SEQUENCE_OK rank {rank} transport {os.environ.get('ORBX_TEST_TRANSPORT')} frames {F} kps {int(gn.sum())}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
