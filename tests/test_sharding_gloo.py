"""CPU: the multi-GPU path is a static split of the stream axis with a final host gather and no data-path collective.
World size 2 over gloo: every rank derives its shard, regenerates exactly its own synthetic streams, and rank 0
gathers per-rank results; the union must equal the unsharded job."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from arm_pose_estimation_b200 import synthetic as syn
from arm_pose_estimation_b200.estimate.batched import gather_host_results, ring_slots, shard_streams


def test_shard_streams_partitions_exactly():
    for n in (0, 1, 7, 8, 1024, 1025):
        for w in (1, 2, 3, 8):
            spans = [shard_streams(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_streams(4, 2, 2)


def test_ring_slots():
    assert ring_slots(1, 6) == 6 and ring_slots(14, 6) == 19 and ring_slots(1, 1) == 1 and ring_slots(0, 0) == 1


def _worker(rank, world, port, n_streams, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard_streams(n_streams, world, rank)
    rows = syn.synth_rows(syn.KIND_WATCH_ONLY, count, 3, config_id=4, first_stream=first)    # this rank's streams only
    result = rows.reshape(count, -1).astype(np.float64).sum(axis=1)                            # stand-in for per-stream results
    gathered = gather_host_results(result, first, n_streams)
    if rank == 0:
        out_q.put(gathered)
    dist.destroy_process_group()


def test_world_size_2_gloo_gather_matches_unsharded():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    n_streams = 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_streams, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = syn.synth_rows(syn.KIND_WATCH_ONLY, n_streams, 3, config_id=4).reshape(n_streams, -1).astype(np.float64).sum(axis=1)
    np.testing.assert_array_equal(got, want)
