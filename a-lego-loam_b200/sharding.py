"""Multi-GPU host logic: the path shards by INDEPENDENT scan sequences — sequence s belongs to rank
s mod world, one process / handle / CUDA stream per GPU, no data-path collective.  torch.distributed (NCCL on the GPU box,
gloo in the CPU tests) is used only for the barrier and for the max-over-ranks of the device-timed region."""


def sequences_of_rank(n_sequences, rank, world):
    """Global sequence ids owned by `rank` (round-robin, like the reference would run one robot per process)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_sequences, world))


def sequence_seed(global_sequence_id, base=100):
    """Seed of the synthetic world / trajectory of a global sequence id: distinct per sequence, independent of world size."""
    return base + int(global_sequence_id)


def reduce_max_ms(values_ms, dist=None, device="cpu"):
    """Element-wise max over ranks of device-timed durations (ms). Without an initialised process group: identity."""
    import torch
    t = torch.tensor(list(values_ms), dtype=torch.float64, device=device)
    if dist is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def whole_job_throughput(scans_per_rank_per_step, steps, world, max_ms):
    """scans/s of the whole job: every rank processed scans_per_rank_per_step * steps sweeps in max_ms (max over ranks)."""
    return scans_per_rank_per_step * steps * world / (max_ms * 1e-3)


def device_order(n_visible):
    """Order in which local ranks take the node's GPUs: alternating between the two halves of the box (0, n/2, 1, n/2+1, ...), so
    that a job on fewer GPUs than the node has spreads its host->device sweep copies over both host I/O halves instead of
    crowding the first one (the end-to-end rate of this path is bounded by pinned-memory H2D bandwidth)."""
    n = int(n_visible)
    if n < 2:
        return list(range(max(n, 0)))
    half = (n + 1) // 2
    order = []
    for i in range(half):
        order.append(i)
        if i + half < n:
            order.append(i + half)
    return order


def device_for_local_rank(local_rank, n_visible, local_world):
    """CUDA device index of a local rank (see device_order); with as many ranks as GPUs every GPU is used exactly once."""
    if n_visible < 1:
        raise ValueError("no visible device")
    if local_world > n_visible:
        return int(local_rank) % n_visible
    return device_order(n_visible)[int(local_rank)]


def gather_rank_stats(stats, dist=None):
    """Every rank's own figures (a small dict) on every rank, ordered by rank — for attributing a max-over-ranks number."""
    if dist is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, stats)
        return out
    return [stats]
