"""
Multi-GPU sharding of the regridding path (one process per GPU).

Every target face owns one CSR row and one output column, so the path shards
with no exchange during the build or the apply: rank r builds and applies the
contiguous target range ``shard_range(n_target, r, world)`` against a
replicated source mesh.  The only collective is one all-gather that
reassembles the (n_extra, n_target) output (NCCL over NVLink on the GPU box;
gloo in the CPU tests).  The reference has no counterpart: its only
parallelism is numba prange over the extra axis
(/root/reference/xugrid/regrid/regridder.py:50).
"""

from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of n items for `rank` of `world`: every rank owns
    ceil(n / world) items except the last ones, which own what is left.  Equal
    strides make the gathered output contiguous (see all_gather_columns)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("invalid rank/world")
    c = max_shard(n, world)
    return min(n, rank * c), min(n, (rank + 1) * c)


def max_shard(n: int, world: int) -> int:
    return (n + world - 1) // world


def all_gather_columns(local: torch.Tensor, n_total: int, group: Optional[dist.ProcessGroup] = None,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """
    ``local`` is this rank's (n_extra, rows_local) block of columns
    ``shard_range(n_total, rank, world)``.  Returns the full (n_extra, n_total)
    array on every rank with ONE all_gather_into_tensor of equally sized slabs
    laid out [rank][n_extra][max_shard].  Because every shard starts at a
    multiple of max_shard, a single time step (n_extra == 1) comes back as a
    contiguous view of the receive buffer, without any reassembly copy.
    """
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local if out is None else out.copy_(local)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_extra = int(local.shape[0])
    pad = max_shard(n_total, world)
    b, e = shard_range(n_total, rank, world)
    if local.shape[1] != e - b:
        raise ValueError(f"rank {rank} holds {local.shape[1]} columns, expected {e - b}")
    if e - b == pad and local.is_contiguous():
        send = local
    else:
        send = local.new_zeros((n_extra, pad))
        send[:, : e - b] = local
    recv = local.new_empty((world * n_extra, pad))  # concatenation along dim 0
    dist.all_gather_into_tensor(recv, send, group=group)
    full = recv.view(world, n_extra, pad).permute(1, 0, 2).reshape(n_extra, world * pad)[:, :n_total]
    if out is not None:
        return out.copy_(full)
    return full


class GatheredOutput:
    """
    The (n_extra, n_total) output of a job whose target columns are sharded by
    ``shard_range`` over the ranks, assembled WITHOUT a collective: the buffer lives in
    symmetric memory (every rank maps every rank's copy over NVLink / NVSwitch) and the
    apply kernel of rank r stores its columns into all copies while it computes them
    (``xg_csr_apply_peers``).  Replaces ``csr_apply`` + ``all_gather_columns`` (send slab,
    padding, NCCL launch): at 8 ranks and 10 M columns 0.04 + 0.20 ms become one kernel.

    ``apply(code, percentile, source, A)`` runs the fused kernel between two device-side
    barriers of the group -- peers must be done reading the previous result before it is
    overwritten, and all stores must have landed before anybody reads -- and returns this
    rank's full (n_extra, n_total) copy.  Outside a process group it is a plain apply.
    """

    def __init__(self, n_extra: int, n_total: int, device=None, group: Optional[dist.ProcessGroup] = None):
        self.n_extra, self.n_total = int(n_extra), int(n_total)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.world == 1:
            self.buf = torch.empty((self.n_extra, self.n_total), dtype=torch.float64, device=dev)
            self.handle, self.peer_ptrs = None, [self.buf.data_ptr()]
            return
        import torch.distributed._symmetric_memory as symm

        self.buf = symm.empty((self.n_extra, self.n_total), dtype=torch.float64, device=dev)
        self.handle = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.peer_ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(self.peer_ptrs) != self.world or len(self.peer_ptrs) > 16:
            raise RuntimeError("symmetric memory rendezvous returned an unexpected number of peers")

    def apply(self, code: int, percentile: float, source: torch.Tensor, A) -> torch.Tensor:
        from . import sparse

        b, e = shard_range(self.n_total, self.rank, self.world)
        if A.n != e - b:
            raise ValueError(f"rank {self.rank} holds {A.n} rows, expected {e - b}")
        if self.handle is not None:
            self.handle.barrier(channel=0)
        sparse.csr_apply_peers(code, percentile, source, A, self.peer_ptrs, self.n_total, b)
        if self.handle is not None:
            self.handle.barrier(channel=1)
        return self.buf


def all_gather_csr(local, n_total: int, group: Optional[dist.ProcessGroup] = None):
    """
    ``local`` is this rank's MatrixCSR over the target rows
    ``shard_range(n_total, rank, world)``.  Returns the full (n_total rows) MatrixCSR on
    every rank: the row lengths travel in one all-gather of equally sized slabs, indices
    and weights in one all-gather each (padded to the largest shard's nnz), and the row
    pointer is rebuilt by one exclusive scan -- the optional second collective of SURVEY
    section 8(e), only paid when a caller asks for ``.weights`` of a sharded regridder.
    """
    from .sparse import MatrixCSR

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    b, e = shard_range(n_total, rank, world)
    rows = e - b
    if local.n != rows:
        raise ValueError(f"rank {rank} holds {local.n} rows, expected {rows}")
    dev = local.data.device
    pad_rows = max_shard(n_total, world)
    counts = torch.zeros(pad_rows, dtype=torch.int64, device=dev)
    counts[:rows] = local.indptr[1:] - local.indptr[:-1]
    all_counts = torch.empty(world * pad_rows, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_counts, counts, group=group)
    all_counts = all_counts.view(world, pad_rows)
    nnz_per_rank = all_counts.sum(dim=1)
    pad_nnz = max(int(nnz_per_rank.max().item()), 1)
    nnz_local = int(local.data.numel())
    send_i = torch.zeros(pad_nnz, dtype=local.indices.dtype, device=dev)
    send_d = torch.zeros(pad_nnz, dtype=local.data.dtype, device=dev)
    send_i[:nnz_local] = local.indices
    send_d[:nnz_local] = local.data
    recv_i = torch.empty(world * pad_nnz, dtype=send_i.dtype, device=dev)
    recv_d = torch.empty(world * pad_nnz, dtype=send_d.dtype, device=dev)
    dist.all_gather_into_tensor(recv_i, send_i, group=group)
    dist.all_gather_into_tensor(recv_d, send_d, group=group)
    sizes = [int(v) for v in nnz_per_rank.tolist()]
    indices = torch.cat([recv_i[r * pad_nnz: r * pad_nnz + sizes[r]] for r in range(world)])
    data = torch.cat([recv_d[r * pad_nnz: r * pad_nnz + sizes[r]] for r in range(world)])
    row_counts = torch.cat([all_counts[r, : shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0]]
                            for r in range(world)])
    indptr = torch.zeros(n_total + 1, dtype=torch.int64, device=dev)
    torch.cumsum(row_counts, 0, out=indptr[1:])
    return MatrixCSR(data, indices, indptr, int(n_total), int(local.m), int(data.numel()))


def sharded_upload(host: torch.Tensor, device: torch.device,
                   group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """
    Replicate a (pinned) host array that every rank holds on every GPU without
    every rank pulling all of it over its own PCIe link: rank r uploads rows
    ``shard_range(n, r, world)`` (1/world of the bytes, all links in parallel) and
    one all_gather_into_tensor over NVLink assembles the full array on each GPU.
    Falls back to a plain upload outside a process group.  Runs on the current
    stream.
    """
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return host.to(device, non_blocking=True)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = int(host.shape[0])
    pad = max_shard(n, world)
    b, e = shard_range(n, rank, world)
    recv = torch.empty((world * pad,) + tuple(host.shape[1:]), dtype=host.dtype, device=device)
    send = torch.empty((pad,) + tuple(host.shape[1:]), dtype=host.dtype, device=device)
    if e > b:
        send[: e - b].copy_(host[b:e], non_blocking=True)
    dist.all_gather_into_tensor(recv, send, group=group)
    return recv[:n]


def bind_to_gpu_cpus(device_index: int):
    """
    Restrict this process to the CPU cores that NVML reports as local to its GPU (same
    NUMA node / PCIe root complex), BEFORE it allocates pinned host memory: Linux places
    pages on the node of the core that first touches them, and a rank whose staging
    buffers sit on the other socket pulls every host<->device copy over the inter-socket
    link.  One process per GPU makes this a per-process setting.  Returns the set of cores,
    or None when nothing was changed (no NVML, no affinity information, or an empty
    intersection with the cores the process may use).
    """
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        local = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cores = local & allowed
        if not cores or cores == allowed:
            return None
        os.sched_setaffinity(0, cores)
        return cores
    except Exception:
        return None


_shared_arrays = []  # (tensor, path): registered host mappings are kept alive for the process


def shared_host_array(shape, dtype: torch.dtype = torch.float64,
                      group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """
    A host array in POSIX shared memory that EVERY rank of the node maps, page-locked
    (``cudaHostRegister``) so that each rank's GPU can DMA into it: the counterpart of
    ``sharded_upload`` for results.  Rank r copies only its own columns device -> host
    (``regridder.regrid(values, out=shared[:, b:e])``) over its own PCIe link and the full
    result is assembled in host memory that all ranks can read -- instead of every rank
    pulling the whole gathered array through the host at once (measured at 8 ranks:
    7.1 ms for 8 x 80 MB against 0.2 ms for 8 x 10 MB).  Outside a process group: a
    plain pinned tensor.  Collective: all ranks must call it with the same shape.
    """
    import math
    import os

    shape = tuple(int(v) for v in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return torch.empty(shape, dtype=dtype).pin_memory()
    rank = dist.get_rank(group)
    numel = max(1, math.prod(shape))
    name = [f"/dev/shm/xg_b200_{os.getpid()}_{len(_shared_arrays)}" if rank == 0 else None]
    dist.broadcast_object_list(name, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    path = name[0]
    if rank == 0:
        t = torch.from_file(path, shared=True, size=numel, dtype=dtype)  # creates and sizes the file
    dist.barrier(group)
    if rank != 0:
        t = torch.from_file(path, shared=True, size=numel, dtype=dtype)
    dist.barrier(group)
    if rank == 0:
        os.unlink(path)  # the mappings stay; the name does not outlive the job
    if torch.cuda.is_available():  # (the gloo tests of the host logic map it without a GPU)
        err = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), numel * t.element_size(), 0)
        if int(err) != 0:
            raise RuntimeError(f"cudaHostRegister of the shared host array failed: {err}")
    _shared_arrays.append((t, path))
    return t[: math.prod(shape)].reshape(shape) if math.prod(shape) else t[:0].reshape(shape)


def target_range_box(xy: torch.Tensor, faces: torch.Tensor, begin: int, end: int):
    """(xmin, xmax, ymin, ymax) of the target faces [begin, end) (device tensors);
    the clip box of a shard's source index."""
    import ctypes

    from . import _native

    L = _native.lib()
    box = (ctypes.c_double * 4)()
    with torch.cuda.device(xy.device):
        _native.check(
            L.xg_face_range_bounds(_native.ptr(xy), _native.ptr(faces), faces.element_size(), int(begin),
                                   int(end), int(faces.shape[1]), box, _native.stream_ptr())
        )
    return (box[0], box[1], box[2], box[3])


def curve_shard(xy: torch.Tensor, faces: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Face numbers (int32 device tensor, no particular order) of the share of ``rank`` when
    the faces of a mesh are partitioned over ``world`` ranks along the space-filling curve
    (``xg_target_shard``): a compact patch of the mesh whatever the caller's face numbering
    means, where ``shard_range`` would hand a rank faces from all over the domain.  Every
    rank computes the same cuts, so the shares are disjoint and cover the mesh."""
    import ctypes

    from . import _native

    L = _native.lib()
    n = int(faces.shape[0])
    ids = torch.empty(n, dtype=torch.int32, device=faces.device)
    count = ctypes.c_int64()
    with torch.cuda.device(faces.device):
        _native.check(
            L.xg_target_shard(_native.ptr(xy, torch.float64), int(xy.shape[0]), _native.ptr(faces),
                              faces.element_size(), n, int(faces.shape[1]), int(rank), int(world),
                              _native.ptr(ids), ctypes.byref(count), _native.stream_ptr())
        )
    return ids[: count.value]


def all_gather_ids(ids: torch.Tensor, group: Optional[dist.ProcessGroup] = None):
    """Every rank's id list: (padded (world, max_len) int32 tensor, lengths).  Done once per
    regridder; ``scatter_gathered`` then places the outputs of every later apply."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return ids.reshape(1, -1), [int(ids.numel())]
    world = dist.get_world_size(group)
    n_local = torch.tensor([ids.numel()], dtype=torch.int64, device=ids.device)
    lens = torch.empty(world, dtype=torch.int64, device=ids.device)
    dist.all_gather_into_tensor(lens, n_local, group=group)
    lens = [int(v) for v in lens.tolist()]
    pad = max(max(lens), 1)
    send = torch.full((pad,), -1, dtype=torch.int32, device=ids.device)
    send[: ids.numel()] = ids
    recv = torch.empty((world, pad), dtype=torch.int32, device=ids.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    return recv, lens


def all_gather_by_ids(local: torch.Tensor, all_ids: torch.Tensor, lens, n_total: int,
                      group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """``local`` is this rank's (n_extra, len(ids)) block of output columns for the faces in
    its id list.  Returns the full (n_extra, n_total) array in the caller's face order on
    every rank: one all_gather_into_tensor of padded slabs + one scatter through the
    gathered id lists (``all_gather_ids``)."""
    n_extra = int(local.shape[0])
    world = int(all_ids.shape[0])
    pad = int(all_ids.shape[1])
    out = torch.full((n_extra, n_total), float("nan"), dtype=local.dtype, device=local.device)
    if world == 1:
        out[:, all_ids[0, : lens[0]].long()] = local
        return out
    send = local.new_zeros((n_extra, pad))
    send[:, : local.shape[1]] = local
    recv = local.new_empty((world * n_extra, pad))
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view(world, n_extra, pad)
    for r in range(world):
        out[:, all_ids[r, : lens[r]].long()] = recv[r, :, : lens[r]]
    return out
