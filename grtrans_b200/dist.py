"""Multi-GPU plumbing: one process per GPU, pixel range sharded in interleaved tiles, one gather of the
float32 Stokes image to rank 0 over NCCL/NVLink (torch.distributed is used for the collective only).

The path has no data-path exchange step: rays are independent (pgrtrans.f90:198-209), so the only collective is the
final image gather."""
import numpy as np

SHARD = 256  # rays per interleaved tile; must match grt_session_create (grt_engine.cu)


def shard_indices(i1, i2, rank, nranks, shard=SHARD):
    """1-based geodesic indices owned by `rank` (same rule as the C++ session)."""
    i = np.arange(i1, i2 + 1, dtype=np.int64)
    return i[((i - i1) // shard) % nranks == rank].astype(np.int32)


def assemble(images, indices, npix):
    """images[r]: [ncams, nlocal_r, nvals] float32, indices[r]: 1-based pixel ids -> full [ncams, npix, nvals]."""
    ncams, _, nv = images[0].shape
    full = np.zeros((ncams, npix, nv), np.float32)
    for img, idx in zip(images, indices):
        if len(idx):
            full[:, np.asarray(idx, np.int64) - 1, :] = img
    return full


class _CudaArray(object):
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class DeviceImage(object):
    """Zero-copy torch views of the session's device-resident image and ray-index buffers."""

    def __init__(self, sess):
        import torch
        p, n, pi, nr = sess.device_image()
        self.ncams, self.nvals, self.nlocal = sess.ncams, sess.nvals, nr
        self.image = torch.as_tensor(_CudaArray(p, (max(n, 1),), "<f4"), device="cuda")[:n]
        self.index = torch.as_tensor(_CudaArray(pi, (max(nr, 1),), "<i4"), device="cuda")[:nr]


def gather_blocks(img_flat, idx, ncams, nvals, rank, world):
    """Gather per-rank flat images [ncams*nlocal*nvals] and index vectors to rank 0 (works with nccl and gloo).
    torch.distributed.gather needs equal sizes, so blocks are padded to the largest rank's count."""
    import torch
    import torch.distributed as dist
    dev = img_flat.device
    nloc = int(idx.numel())
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([nloc], dtype=torch.int64, device=dev))
    counts = [int(c.item()) for c in counts]
    cmax = max(counts)
    if nloc == cmax:
        pimg, pidx = img_flat, idx
    else:
        pimg = torch.zeros(ncams * cmax * nvals, dtype=img_flat.dtype, device=dev)
        pimg.view(ncams, cmax, nvals)[:, :nloc, :] = img_flat.view(ncams, nloc, nvals)
        pidx = torch.zeros(cmax, dtype=idx.dtype, device=dev)
        pidx[:nloc] = idx
    if rank == 0:
        imgs = [torch.empty(ncams * cmax * nvals, dtype=img_flat.dtype, device=dev) for _ in range(world)]
        idxs = [torch.empty(cmax, dtype=idx.dtype, device=dev) for _ in range(world)]
    else:
        imgs = idxs = None
    dist.gather(pimg, imgs, dst=0)
    dist.gather(pidx, idxs, dst=0)
    if rank != 0:
        return None, None, counts
    imgs = [im.view(ncams, cmax, nvals)[:, :c, :] for im, c in zip(imgs, counts)]
    idxs = [ix[:c] for ix, c in zip(idxs, counts)]
    return imgs, idxs, counts


def gather_image(dimg, npix, world, rank):
    """NCCL gather of every rank's [ncams, nlocal, nvals] block; rank 0 scatters them into the full device image."""
    import torch
    imgs, idxs, counts = gather_blocks(dimg.image, dimg.index, dimg.ncams, dimg.nvals, rank, world)
    if rank != 0:
        return None
    full = torch.zeros(dimg.ncams, npix, dimg.nvals, dtype=torch.float32, device="cuda")
    for im, ix, c in zip(imgs, idxs, counts):
        if c:
            full[:, ix.long() - 1, :] = im
    return full
