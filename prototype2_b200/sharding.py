"""Multi-GPU plumbing of the collision path (one process per GPU, torch.distributed).

Static pass: characters are independent units -> contiguous index ranges per rank, no collective.
Dynamic pass (PRTC_DYN_JACOBI): one exchange per frame -- every rank's 64-byte character records (start-of-frame
capsule segment + stored fat box).  Two ways:
  * ipc_connect(): the ranks swap CUDA IPC handles once; from then on the kernel that computes the records stores them
    straight into every peer's array (peer stores over NVLink) and a device-side arrival table orders the frames -- no
    collective and no host call per frame (prtc_dyn_ipc_*);
  * allgather_records(): the records are all-gathered IN PLACE into the record array each context owns (prtc_dyn_buffer),
    NCCL over NVLink on GPUs, gloo on CPU tensors in the tests.
"""
import numpy as np

RECORD_FLOATS = 16          # 4 x float4, see kernels.cu k_dyn_prep


def partition(n_global, world, rank):
    """contiguous slice [first, first + count) of rank `rank`; equal counts (all-gather needs them), remainder dropped
    into the last ranks one by one is NOT allowed -> n_global must be divisible by world"""
    if n_global % world != 0:
        raise ValueError("n_global (%d) must be divisible by the world size (%d)" % (n_global, world))
    count = n_global // world
    return rank * count, count


class _CudaBuffer:
    """minimal __cuda_array_interface__ holder so torch can alias a raw device pointer owned by libprtc"""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def records_tensor(ctx):
    """torch view (n_global, 16) float32 over the context's record array"""
    import torch
    ptr, record_bytes, first, n_global = ctx.dyn_buffer()
    assert record_bytes == 4 * RECORD_FLOATS
    flat = torch.as_tensor(_CudaBuffer(ptr, n_global * RECORD_FLOATS), device="cuda")
    return flat.view(n_global, RECORD_FLOATS), first


def allgather_records(records, first, count, group=None):
    """in-place all-gather: rank r contributes rows [first, first+count) of `records` (torch tensor, any device)"""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return records
    dist.all_gather_into_tensor(records, records[first:first + count], group=group)
    return records


def gather_rows(local_rows, group=None):
    """concatenate per-rank row blocks in rank order (used to compare a sharded run with an unsharded one)"""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_rows
    out = [torch.empty_like(local_rows) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, local_rows, group=group)
    return torch.cat(out, dim=0)


def ipc_connect(ctx, rank, world, group=None):
    """One-time set-up of the collective-free record exchange: every rank exports its handles, all ranks import all of them.
    Returns False (and leaves the context on the all-gather path) when the devices cannot map each other's memory."""
    import torch.distributed as dist
    from . import capi
    mine = ctx.dyn_ipc_export(rank, world)
    blobs = [None] * world
    dist.all_gather_object(blobs, mine, group=group)
    try:
        ctx.dyn_ipc_import(blobs)
        ok = True
    except capi.PrtcError:
        ok = False
    flags = [None] * world
    dist.all_gather_object(flags, ok, group=group)
    if not all(flags):                                             # some rank cannot map its peers: everybody stays on the all-gather path
        ctx.dyn_ipc_import(None)
        return False
    return True
