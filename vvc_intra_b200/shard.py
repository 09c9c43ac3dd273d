"""Multi-GPU layout of the path: all-intra pictures are independent (IntraPeriod 1, GOPSize 1: BIN/encoder_intra.cfg:26-28;
CABAC contexts re-initialise per slice, EL/EncSlice.cpp:1640-1647), so (frame, QP) work units are dealt round-robin to the
ranks -- one process per GPU -- and nothing is exchanged on the data path.  torch.distributed carries only the start
barrier, the max-over-ranks of the device time and the final gather of per-unit statistics."""


def work_units(n_frames, qps):
    """The n_frames x len(qps) (frame, qp) units of one job in the order the single-GPU run walks them: consecutive
    units use different frames and different QPs."""
    return [(s % n_frames, qps[(s // n_frames + s % n_frames) % len(qps)]) for s in range(n_frames * len(qps))]


def shard_units(units, rank, world):
    """Units of `rank`: frame f of the sequence goes to GPU (index mod world)."""
    if not 0 <= rank < world:
        raise ValueError('rank %d outside world of %d' % (rank, world))
    return units[rank::world] if world > 1 else list(units)


def max_over_ranks(value, dist=None, device=None):
    """The slowest rank's time (the job is done when the last GPU is)."""
    if dist is None:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_stats(stats, dist=None):
    """Final gather of the per-unit results (list of picklable records per rank) on every rank, in unit order."""
    if dist is None:
        return list(stats)
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, list(stats))
    merged = []
    for i in range(max(len(o) for o in out)):
        for o in out:
            if i < len(o):
                merged.append(o[i])
    return merged
