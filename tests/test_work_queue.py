"""CPU tests of the multi-GPU host logic with world_size 2 over gloo (no GPU): every prompt is claimed exactly once, the
dynamic queue balances variable-length trajectories better than a static split, max-over-ranks timing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tpdm_b200.work_queue import PromptQueue, static_shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, lengths, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tpdm_b200.work_queue import gather_results, max_over_ranks, sample_prompts

    import time

    busy = [0.0]

    def fake_trajectory(i):
        # a fake step() that consumes the pre-drawn trajectory length of prompt i
        time.sleep(0.002 * lengths[i])
        busy[0] += lengths[i]
        return {"steps": lengths[i], "rank": rank}

    mine = sample_prompts(fake_trajectory, len(lengths), dynamic=True, name="q1")
    merged = gather_results(mine)
    # a second queue of the same name in the same process group starts from zero (its own store key)
    again = gather_results(sample_prompts(lambda i: {"steps": 1, "rank": rank}, 5, dynamic=True, name="q1"))
    if rank == 0:
        assert sorted(again) == [0, 1, 2, 3, 4]
    worst = max_over_ranks(busy[0])
    dist.barrier()
    if rank == 0:
        out_q.put((merged, worst))
    dist.destroy_process_group()


def test_dynamic_queue_world_size_2():
    g = torch.Generator().manual_seed(7)
    lengths = torch.randint(6, 29, (24,), generator=g).tolist()   # variable-length trajectories, 6..28 steps
    lengths[0] = 60                                               # one straggler
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, worst = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(merged) == list(range(24))                      # each prompt exactly once
    assert all(merged[i]["steps"] == lengths[i] for i in merged)
    total = sum(lengths)
    static_worst = max(sum(lengths[i] for i in static_shard(24, r, 2)) for r in range(2))
    assert worst <= static_worst                                  # never worse than round-robin
    assert worst <= total / 2 + max(lengths)                      # list-scheduling bound: makespan <= mean + longest job
    assert {merged[i]["rank"] for i in merged} == {0, 1}


def test_longest_expected_first_order():
    """The ticket order of the device queue's 'lpt' schedule: estimates from the sigma after the first step, longest first, stable."""
    from tpdm_b200.work_queue import expected_remaining_steps, longest_first_order

    s1 = torch.tensor([0.73, 0.30, 0.95, 0.0005, 0.73, 0.50])
    est = expected_remaining_steps(s1, min_sigma=1e-3, max_steps=28)
    # a constant ratio of 0.73 needs 22 steps to get below 1e-3 (0.73^22 = 9.8e-4): 21 more after the first
    assert est.tolist() == [21.0, 5.0, 27.0, -1.0, 21.0, 9.0]
    order = longest_first_order(est).tolist()
    assert order == [2, 0, 4, 5, 1, 3]                  # slowest schedule first, ties in prompt order, the finished prompt last
    # list scheduling with that order never does worse than prompt order on this instance (2 workers)
    def makespan(seq):
        load = [0.0, 0.0]
        for i in seq:
            load[load.index(min(load))] += float(est[i]) + 1
        return max(load)
    assert makespan(order[:-1]) <= makespan([0, 1, 2, 4, 5])


def test_queue_without_process_group():
    q = PromptQueue(3)
    assert list(q) == [0, 1, 2] and q.claim() is None
    q2 = PromptQueue(2)
    assert q2.key != q.key and list(q2) == [0, 1]
    assert static_shard(10, 1, 4) == [1, 5, 9]
