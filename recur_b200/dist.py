"""Host-side plumbing of a multi-GPU run (one process per GPU).

Streams shard across ranks; the only data-path exchange is the sum of the
shared delta arrays, which the library does itself over NCCL
(rnn_b200_comm_join).  What remains for the launcher is: give every rank its
stretch of the text and its device, pass the 128-byte NCCL id around, and
reduce timings.  These helpers use torch.distributed for that (NCCL on the
GPU box, gloo in the CPU tests) and nothing else.
"""
import ctypes as C
import os


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_bounds(n_items, rank, world):
    """Contiguous, disjoint, equal-length shards (the remainder is dropped so
    that every rank does identical work)."""
    per = n_items // world
    return rank * per, (rank + 1) * per


def stream_positions(text_len, n_streams, step):
    """Text positions read by the n streams of one rank at character position
    `step`: stream j reads step + j * ((len-1)/n), wrapped — the spacing rule
    of rnn_char_epoch (charmodel-predict.c:273,295-298)."""
    spacing = (text_len - 1) // n_streams
    return [(step + j * spacing) % (text_len - 1) for j in range(n_streams)]


def broadcast_bytes(dist, payload, n_bytes, src=0, device=None):
    """Rank `src` supplies `payload` (bytes); everyone returns the same bytes."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return bytes(payload)
    buf = torch.zeros(n_bytes, dtype=torch.uint8, device=device or "cpu")
    if dist.get_rank() == src:
        buf.copy_(torch.tensor(list(payload), dtype=torch.uint8))
    dist.broadcast(buf, src)
    return bytes(buf.cpu().tolist())


def max_over_ranks(dist, value, device=None):
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def join_comm(lib, dist, rank, world, device=None):
    """Create the library's NCCL communicator over all ranks."""
    if world <= 1:
        return
    idbuf = (C.c_uint8 * 128)()
    if rank == 0 and lib.rnn_b200_comm_unique_id(idbuf) != 0:
        raise RuntimeError("NCCL could not be loaded")
    raw = broadcast_bytes(dist, bytes(idbuf), 128, 0, device)
    idbuf = (C.c_uint8 * 128).from_buffer_copy(raw)
    if lib.rnn_b200_comm_join(idbuf, rank, world) != 0:
        raise RuntimeError("rnn_b200_comm_join failed")
