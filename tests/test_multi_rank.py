"""N > 1 host logic on CPU: world_size-2 gloo — rank -> sequence assignment is a disjoint cover, seeds do not depend on the
world size, and the timing reduction is the max over ranks (what bench.py does with NCCL on the GPU box)."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_seq, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import alego_pkg
    alego = alego_pkg.load()
    from alego_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.sequences_of_rank(n_seq, rank, world)
    seeds = [sharding.sequence_seed(s) for s in mine]
    # pretend device times: rank r took 10 + 5 r ms (value) and 20 - 3 r ms (e2e)
    red = sharding.reduce_max_ms([10.0 + 5 * rank, 20.0 - 3 * rank], dist)
    stats = sharding.gather_rank_stats({"rank": rank, "device": sharding.device_for_local_rank(rank, 8, world), "ms": 10.0 + 5 * rank}, dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    dist.barrier()
    q.put((rank, mine, seeds, red, gathered, stats))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_seq", [7, 16])
def test_two_rank_sharding_and_time_reduction(n_seq):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_seq, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, s0, red0, g0, st0), (r1, m1, s1, red1, g1, st1) = out
    assert sorted(m0 + m1) == list(range(n_seq)) and not set(m0) & set(m1)      # disjoint cover
    assert abs(len(m0) - len(m1)) <= 1                                          # balanced
    assert red0 == red1 == [15.0, 20.0]                                         # max over ranks, element-wise
    assert g0 == g1 == [m0, m1]
    assert len(set(s0 + s1)) == n_seq                                           # distinct worlds per sequence
    # per-rank attribution: every rank sees every rank's figures, in rank order; two ranks on an 8-GPU node take GPUs 0 and 4
    assert st0 == st1 == [{"rank": 0, "device": 0, "ms": 10.0}, {"rank": 1, "device": 4, "ms": 15.0}]


def test_single_process_helpers():
    sys.path.insert(0, ROOT)
    import alego_pkg
    alego_pkg.load()
    from alego_b200 import sharding
    assert sharding.sequences_of_rank(5, 0, 1) == [0, 1, 2, 3, 4]
    assert sharding.sequences_of_rank(5, 3, 4) == [3]
    assert sharding.sequences_of_rank(2, 3, 4) == []
    with pytest.raises(ValueError):
        sharding.sequences_of_rank(5, 4, 4)
    assert sharding.reduce_max_ms([1.5, 2.5]) == [1.5, 2.5]
    # seeds are a function of the global sequence id only (weak scaling adds sequences, it does not reshuffle them)
    assert [sharding.sequence_seed(s) for s in sharding.sequences_of_rank(8, 1, 2)] == [101, 103, 105, 107]
    assert sharding.whole_job_throughput(64, 10, 8, 500.0) == 64 * 10 * 8 / 0.5
    # rank -> GPU: a permutation of the node's GPUs that alternates between its two halves
    assert sharding.device_order(8) == [0, 4, 1, 5, 2, 6, 3, 7] and sharding.device_order(1) == [0] and sharding.device_order(2) == [0, 1]
    assert sorted(sharding.device_order(7)) == list(range(7))
    assert [sharding.device_for_local_rank(r, 8, 4) for r in range(4)] == [0, 4, 1, 5]
    assert sharding.device_for_local_rank(3, 2, 4) == 1  # more ranks than GPUs: wrap
    assert sharding.gather_rank_stats({"a": 1}) == [{"a": 1}]
