"""Multi-GPU sharding of independent sequences: one process per GPU, sequence s goes to
rank s mod world, every rank encodes its shard with its own batch encoder, and the streams are gathered IN ORDER
on rank 0 through torch.distributed (object gather over the host; the data path has no collective)."""
import torch.distributed as dist


def my_shard(nseq, rank, world):
    return list(range(rank, nseq, world))


def gather_in_order(local_items, nseq, rank, world):
    """local_items: list of (sequence index, payload) of this rank.  Returns the full ordered list on rank 0."""
    if world == 1:
        out = [None] * nseq
        for s, p in local_items:
            out[s] = p
        return out
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(local_items, bucket, dst=0)
    if rank != 0:
        return None
    out = [None] * nseq
    for items in bucket:
        for s, p in items:
            out[s] = p
    assert all(o is not None for o in out), "a shard is missing"
    return out
