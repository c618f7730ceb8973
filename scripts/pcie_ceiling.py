#!/usr/bin/env python
"""Platform ceiling for the end-to-end number: bare pinned-memory cudaMemcpyAsync H2D + D2H on N GPUs at the same time.

    python scripts/pcie_ceiling.py                       # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/pcie_ceiling.py

Every rank copies in the sequence path's proportions (per stereo frame 2 x 466,616 B in, one 272,016-B record out) in 8-frame
chunks on two streams, with nothing else running; rank 0 prints one JSON line with the aggregate GB/s and the frames/s that
byte rate would carry -- the upper bound of bench.py's `e2e` on this box, whatever the kernels do."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")
    frame_in, frame_out, chunk = 2 * 1241 * 376, 272016, 8
    n_frames = 568 if world > 1 else 2048
    wc = "--write-combined" in sys.argv
    h_in = torch.empty(n_frames * frame_in, dtype=torch.uint8, pin_memory=True)
    h_in.random_(0, 255)
    h_out = torch.empty(n_frames * frame_out, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(64 * frame_in, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(64 * frame_out, dtype=torch.uint8, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    results = {}

    def run(do_in, do_out, reps=6):
        def once():
            for f0 in range(0, n_frames, chunk):
                slot = (f0 // chunk) % 8
                if do_in:
                    with torch.cuda.stream(s_in):
                        d_in[slot * chunk * frame_in:(slot + 1) * chunk * frame_in].copy_(h_in[f0 * frame_in:(f0 + chunk) * frame_in], non_blocking=True)
                if do_out:
                    with torch.cuda.stream(s_out):
                        h_out[f0 * frame_out:(f0 + chunk) * frame_out].copy_(d_out[slot * chunk * frame_out:(slot + 1) * chunk * frame_out], non_blocking=True)
            torch.cuda.synchronize()

        once()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return reps * n_frames * world / float(t.item())

    for name, a, b in (("h2d_only", True, False), ("d2h_only", False, True), ("h2d_and_d2h", True, True)):
        fps = run(a, b)
        results[name] = {"frames_per_s": fps, "h2d_GBps": fps * frame_in / 1e9 if a else 0.0, "d2h_GBps": fps * frame_out / 1e9 if b else 0.0}
    if rank == 0:
        print(json.dumps({"what": "bare pinned cudaMemcpyAsync ceiling, sequence-path byte mix", "n_gpus": world, "chunk_frames": chunk,
                          "bytes_in_per_frame": frame_in, "bytes_out_per_frame": frame_out, "cpus": os.cpu_count(), "write_combined": wc, **results}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
