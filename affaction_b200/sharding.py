"""Static sharding of candidate rollouts over the GPUs of one box.

Candidates are independent (the reference runs them on independent threads with
cloned worlds, src/ActionComponent.cpp:184-198), so rank r simply takes the
contiguous block [r*N/G, (r+1)*N/G) of the global candidate index space and the
only exchange is an all-gather of the fixed-size verdict records at the end
(the future.wait() join of the reference, :202-205)."""
import ctypes as C


def shard_bounds(total, rank, world):
    """Contiguous block of rank `rank`."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    lo = (total * rank) // world
    hi = (total * (rank + 1)) // world
    return lo, hi


def gather_verdicts(local_bytes, world, dist=None):
    """All-gather equally sized uint8 verdict buffers (torch tensors on the
    backend's device) in rank order; returns one tensor of world*len bytes."""
    import torch
    if world == 1:
        return local_bytes
    out = torch.empty(world * local_bytes.numel(), dtype=torch.uint8, device=local_bytes.device)
    dist.all_gather_into_tensor(out, local_bytes)
    return out


def records_from_bytes(buf, record_type):
    n = len(buf) // C.sizeof(record_type)
    return list((record_type * n).from_buffer_copy(buf))
