"""Multi-GPU plumbing of the hot path: one process per GPU, `torch.distributed` for the rendezvous.

The path shards with no data-path collective (mirrors PartitionMode::CollectLeft, interval_join.rs:396-399,466-480):
  * the build-side index is built on rank 0 and REPLICATED: its single device slab, which starts with its own POD header,
    travels with one broadcast (NCCL over NVLink on GPUs) after an 8-byte size broadcast;
  * probe rows are split into contiguous row ranges, rank g gets rows [g*n/N, (g+1)*n/N);
  * every rank produces the results of its own range; concatenating the per-rank results in rank order reproduces the
    single-GPU row order (pairs: probe_idx is rebased by the shard's first row).

Nothing here touches CUDA directly: the index object only has to offer device_bytes and export() -> (header bytes, slab
tensor), the context slab_alloc(nbytes) / attach_slab(ptr, nbytes, stream) (or attach_tensor(slab) for the CPU stand-in of
the world_size-2 gloo tests), so the same code runs over gloo with CPU tensors.
"""
import numpy as np


def shard_bounds(n_rows, world):
    """[(lo, hi)] * world: contiguous, ordered, sizes differ by at most one row."""
    n_rows, world = int(n_rows), int(world)
    if world < 1:
        raise ValueError("world must be >= 1")
    return [(g * n_rows // world, (g + 1) * n_rows // world) for g in range(world)]


def shard_range(n_rows, rank, world):
    if not 0 <= rank < world:
        raise ValueError("rank outside the world")
    return shard_bounds(n_rows, world)[rank]


def bind_near_gpu(local_rank):
    """One process per GPU: restrict this process to the CPU cores next to its GPU (NVML's ideal affinity), so that the pinned
    staging buffers it allocates afterwards are first-touched on the GPU's NUMA node.  With several ranks streaming probe
    tables over PCIe at once, buffers that all sit on one node cap the end-to-end rate of the whole box.  Returns the
    number of cores the process is bound to, or 0 when NVML is not usable (nothing is changed then)."""
    import os

    try:
        import pynvml
        import torch

        pynvml.nvmlInit()
        try:
            pr = torch.cuda.get_device_properties(local_rank)
            bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return len(os.sched_getaffinity(0))
    except Exception:
        return 0


def _spread(slab, dist, rank, world, src):
    """`slab` (uint8, same size on every rank) becomes rank `src`'s on every rank.

    Default: ONE broadcast (NCCL over NVSwitch: 285 MB in 0.73 ms = 389 GB/s at N = 4).  BRG_REPLICATE=spread: scatter +
    all-gather instead -- rank `src` sends each rank a different 1/world of the slab, then all ranks exchange their parts, both
    steps in place on views of the slab, the last < world * 256 bytes as a small broadcast.  On paper the bandwidth-optimal
    broadcast; through torch.distributed's send/recv scatter it is the slower one (1.05 ms at N = 4,
    profiles/r2_4gpu_replication.txt), so it stays an option."""
    import os

    n = slab.numel()
    part = (n // world) // 256 * 256
    if world <= 2 or part == 0 or os.environ.get("BRG_REPLICATE", "broadcast") != "spread" or not hasattr(dist, "all_gather_into_tensor"):
        dist.broadcast(slab, src)
        return
    body = slab[: part * world]
    mine = body[rank * part:(rank + 1) * part]
    dist.scatter(mine, [body[g * part:(g + 1) * part] for g in range(world)] if rank == src else None, src)
    dist.all_gather_into_tensor(body, mine)
    if part * world < n:
        dist.broadcast(slab[part * world:], src)


def replicate_index(ctx, index, dist, rank, world, device, src=0):
    """Every rank returns an index equal to rank `src`'s.  `index` is only read on rank `src`.

    The slab carries its own header (brg_index_attach reads it back on the receiver), so the data path moves ONE buffer (see
    _spread); the only thing a receiver needs first is the byte count (an 8-byte broadcast)."""
    import torch

    if world == 1:
        return index
    size = torch.tensor([index.device_bytes if rank == src else 0], dtype=torch.int64, device=device)
    dist.broadcast(size, src)
    nbytes = int(size.item())
    if rank == src:
        _, slab = index.export()
        _spread(slab, dist, rank, world, src)
        return index
    if hasattr(ctx, "slab_alloc"):
        ptr, slab = ctx.slab_alloc(nbytes)
        _spread(slab, dist, rank, world, src)
        st = torch.cuda.current_stream().cuda_stream if torch.cuda.is_available() and str(device).startswith("cuda") else None
        return ctx.attach_slab(ptr, nbytes, st)
    # CPU stand-in of the gloo tests: same protocol, the slab's first bytes are the header
    slab = torch.empty(nbytes, dtype=torch.uint8, device=device)
    _spread(slab, dist, rank, world, src)
    return ctx.attach_tensor(slab)


def gather_rows_in_order(local, dist, rank, world, device="cpu"):
    """Concatenation of every rank's 1-D numpy array in rank order, on every rank (variable lengths)."""
    import torch

    if world == 1:
        return local
    local = np.ascontiguousarray(local)
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    raw = np.zeros(cap * local.dtype.itemsize, np.uint8)
    raw[: local.nbytes] = local.view(np.uint8).reshape(-1)
    mine = torch.from_numpy(raw).to(device)
    bufs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(bufs, mine)
    parts = [b.cpu().numpy()[: s * local.dtype.itemsize].view(local.dtype) for b, s in zip(bufs, sizes)]
    return np.concatenate(parts)


class ShardedProbe:
    """Probe-side driver of one rank: slices the probe columns to this rank's row range and rebases row ids."""

    def __init__(self, index, rank, world):
        self.index, self.rank, self.world = index, int(rank), int(world)

    def _slice(self, key, start, end):
        lo, hi = shard_range(len(start), self.rank, self.world)
        return lo, (None if key is None else key[lo:hi]), start[lo:hi], end[lo:hi]

    def count_overlaps(self, key, start, end, strict=False):
        _, k, s, e = self._slice(key, start, end)
        return self.index.count_overlaps(k, s, e, strict=strict)

    def coverage(self, key, start, end, strict=False):
        _, k, s, e = self._slice(key, start, end)
        return self.index.coverage(k, s, e, strict=strict)

    def nearest_k1(self, key, start, end, include_overlaps=True, strict=False, compute_distance=True):
        _, k, s, e = self._slice(key, start, end)
        return self.index.nearest_k1(k, s, e, include_overlaps=include_overlaps, strict=strict,
                                     compute_distance=compute_distance)

    def overlap_pairs(self, key, start, end):
        """(build_idx, probe_idx) of this shard with probe_idx in GLOBAL row numbering."""
        lo, k, s, e = self._slice(key, start, end)
        b, p = self.index.overlap_pairs(k, s, e)
        return b, (p.astype(np.uint64) + lo).astype(np.uint32)
