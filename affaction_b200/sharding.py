"""Static sharding of candidate rollouts over the GPUs of one box.

Candidates are independent (the reference runs them on independent threads with
cloned worlds, src/ActionComponent.cpp:184-198), so rank r simply takes the
contiguous block [r*N/G, (r+1)*N/G) of the global candidate index space and the
only exchange is a gather of the fixed-size verdict records to rank 0 at the end
(the future.wait() join of the reference, :202-205).  bench.py shards with these
helpers; inside one process aff_predict_batch_multi does the same split in C."""
import ctypes as C


def shard_bounds(total, rank, world):
    """Contiguous block of rank `rank`."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    lo = (total * rank) // world
    hi = (total * (rank + 1)) // world
    return lo, hi


def gather_verdicts(local_bytes, world, dist=None, rank=0, out=None):
    """Gather equally sized uint8 verdict buffers (torch tensors on the backend's device) to rank 0 in
    rank order: rank 0 gets one tensor of world*len bytes (`out` if given), the others None."""
    import torch
    if world == 1:
        return local_bytes
    if rank == 0:
        if out is None:
            out = torch.empty(world * local_bytes.numel(), dtype=torch.uint8, device=local_bytes.device)
        dist.gather(local_bytes, list(out.chunk(world)), dst=0)
        return out
    dist.gather(local_bytes, None, dst=0)
    return None


def records_from_bytes(buf, record_type):
    n = len(buf) // C.sizeof(record_type)
    return list((record_type * n).from_buffer_copy(buf))
