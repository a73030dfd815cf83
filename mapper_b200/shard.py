"""Multi-GPU host logic: reads shard by contiguous query ranges (pairs stay together), the reference / index /
duplication table are replicated per GPU, and the only exchange step is the int32 sum of the per-position count
planes (QV/DirectionalAlignments.java:20-28 makes the counts integers, so the reduction is exact
and order-free).  One process per GPU; torch.distributed carries the collective (NCCL on GPUs, gloo in CPU tests)."""
import numpy as np


def shard_bounds(n_queries, rank, world):
    """Contiguous, balanced [lo, hi) of queries for this rank."""
    base, extra = divmod(n_queries, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def take_shard(batch, lo, hi):
    """Slices a packed batch (mapper_b200.synth / xm_align_batch layout) to queries [lo, hi); sequence word offsets are rebased."""
    n_seqs = np.asarray(batch["n_seqs"])
    first = np.concatenate([[0], np.cumsum(n_seqs, dtype=np.int64)])
    s0, s1 = int(first[lo]), int(first[hi])
    off = np.asarray(batch["seq_word_off"])
    w0, w1 = int(off[s0]), int(off[s1])
    packed = np.ascontiguousarray(batch["packed"][w0:w1])
    if len(packed) == 0:
        packed = np.zeros(1, dtype=np.uint16)
    return dict(packed=packed, seq_word_off=np.ascontiguousarray(off[s0:s1 + 1] - w0), seq_len=np.ascontiguousarray(batch["seq_len"][s0:s1]),
                n_seqs=np.ascontiguousarray(n_seqs[lo:hi]), expected_inner=np.ascontiguousarray(batch["expected_inner"][lo:hi]),
                per_penalty=np.ascontiguousarray(batch["per_penalty"][lo:hi]))


def allreduce_planes_(planes):
    """In-place sum of an int32 tensor of count planes over all ranks (NCCL over NVLink on GPUs)."""
    import torch
    import torch.distributed as dist
    assert planes.dtype == torch.int32
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(planes, op=dist.ReduceOp.SUM)
    return planes


def wrap_device_planes(ptr, n_int32, device):
    """Zero-copy torch view of the library's device count planes (xm_counts_device_ptr) for the collective."""
    import torch

    class _Planes:
        __cuda_array_interface__ = dict(shape=(int(n_int32),), typestr="<i4", data=(int(ptr), False), version=2)
    return torch.as_tensor(_Planes(), device=device)
