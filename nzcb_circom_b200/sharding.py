"""Multi-GPU partitioning of the hot path: independent proofs, one process per GPU, no data-path collective
(SURVEY.md 8e).  Rank r proves passes r, r + W, r + 2W, ...; results are gathered on rank 0 only because a caller
wants them in one place -- the proving itself never communicates."""


def shard_indices(n_items, rank, world):
    """indices of the items rank `rank` of `world` proves (round robin, so ragged batches stay balanced)"""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    return list(range(rank, n_items, world))


def gather_results(local_results, n_items, rank, world, dist=None):
    """local_results: what this rank produced for shard_indices(n_items, rank, world), in that order.
    Returns the full list in item order on rank 0 (None elsewhere).  `dist` = torch.distributed (any backend)."""
    if world == 1:
        return list(local_results)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(list(local_results), gathered, dst=0)
    if rank != 0:
        return None
    out = [None] * n_items
    for r, part in enumerate(gathered):
        for i, v in zip(shard_indices(n_items, r, world), part):
            out[i] = v
    return out


def torch_allgather_bytes(dist, device=None):
    """allgather callable for Context.set_msm_split on top of torch.distributed: NCCL over NVLink when `device` is a
    CUDA device (the 128-byte partial sums of the latency mode), gloo otherwise."""
    import torch

    world = dist.get_world_size()

    def allgather(send: bytes) -> bytes:
        t = torch.frombuffer(bytearray(send), dtype=torch.uint8)
        if device is not None:
            t = t.to(device)
        out = torch.empty(world * t.numel(), dtype=torch.uint8, device=t.device)
        dist.all_gather_into_tensor(out, t)
        return out.cpu().numpy().tobytes()

    return allgather
