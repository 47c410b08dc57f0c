"""Multi-GPU plumbing for launchers that start one process per GPU (torchrun, mpirun).

The data path is inside libgrtrans_b200.so: rays i1..i2 are sharded over the ranks in interleaved 256-ray shards
(grt_shard_rays), and grt_session_launch_gather sends every finished tile of the float32 Stokes image to rank 0
with ncclSend / ncclRecv behind the kernels of the following tiles.  The path has no other exchange step: rays are
independent (pgrtrans.f90:198-209).  What is left for the launcher -- and all this module does -- is to carry the 128-byte
NCCL unique id from rank 0 to the other ranks and to reduce a few scalars (timings); torch.distributed (nccl or gloo) is
used for that and for nothing else.  A single process drives several GPUs without any of this: grt_main() does it with
one host thread per device.
"""
import numpy as np

SHARD = 256  # rays per interleaved shard; must match GRT_SHARD (grt_engine.cu)


def shard_indices(i1, i2, rank, nranks, shard=SHARD):
    """1-based geodesic indices owned by `rank` (numpy mirror of grt_shard_rays)."""
    i = np.arange(i1, i2 + 1, dtype=np.int64)
    return i[((i - i1) // shard) % nranks == rank].astype(np.int32)


def shard_pixels(q, rank, nranks, i1=1, shard=SHARD):
    """0-based pixel of local ray q of `rank`: the scatter rule of k_gather_scatter / convert_block."""
    q = np.asarray(q, np.int64)
    return (i1 - 1) + ((q // shard) * nranks + rank) * shard + q % shard


def assemble(images, indices, npix):
    """images[r]: [ncams, nlocal_r, nvals] float32, indices[r]: 1-based pixel ids -> full [ncams, npix, nvals]."""
    ncams, _, nv = images[0].shape
    full = np.zeros((ncams, npix, nv), np.float32)
    for img, idx in zip(images, indices):
        if len(idx):
            full[:, np.asarray(idx, np.int64) - 1, :] = img
    return full


def broadcast_unique_id(make_id, rank, device=None):
    """Rank 0 calls make_id() (-> 128 bytes); every rank returns the same bytes (torch.distributed broadcast)."""
    import torch
    import torch.distributed as dist
    t = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        t = torch.frombuffer(bytearray(make_id()), dtype=torch.uint8).clone()
        if device is not None:
            t = t.to(device)
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


def reduce_max(x, device=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(x, device=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
