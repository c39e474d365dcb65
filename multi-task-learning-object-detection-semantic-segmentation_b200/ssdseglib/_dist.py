"""The only collective on this path: an all-gather of the per-rank evaluation counts.

Images are independent, so encode / decode / NMS / TP flagging shard over ranks with no communication; the
per-class (TP, FP, n_gt) int64 table (C x 3) is gathered so every rank can form the global counts.

On GPUs the gather is `ssdx_allgather_counts` (include/ssdx.h): one ncclAllGather on the caller's stream over a
communicator the library owns.  torch.distributed is only the out-of-band channel that carries the 128-byte NCCL
unique id from rank 0 to the other ranks when the communicator is created.  CPU tensors (the gloo tests) go
through torch.distributed.all_gather_into_tensor."""
from __future__ import annotations

import ctypes as C
import threading

_comms = {}
_comms_lock = threading.Lock()


def shard_range(n_items: int, rank: int, world_size: int):
    """Contiguous batch split: items [lo, hi) of rank `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class CountsCommunicator:
    """NCCL communicator of libssdx over the ranks of a torch.distributed process group (one per device).
    Creation is collective: every rank of the group must construct it."""

    def __init__(self, process_group=None, device=None):
        import torch
        import torch.distributed as dist

        from . import _ssdx

        torch = _ssdx.require_cuda()
        self._sx = _ssdx
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        uid = (C.c_uint8 * 128)()
        if self.rank == 0:
            _ssdx.check(_ssdx.lib.ssdx_comm_unique_id(uid))
        if self.world > 1:
            backend = dist.get_backend(process_group)
            dev = f"cuda:{self.device}" if backend == "nccl" else "cpu"
            t = torch.tensor(list(uid), dtype=torch.uint8, device=dev)
            src = dist.get_global_rank(process_group, 0) if process_group is not None else 0
            dist.broadcast(t, src=src, group=process_group)
            uid = (C.c_uint8 * 128)(*t.cpu().tolist())
        handle = C.c_void_p()
        _ssdx.check(_ssdx.lib.ssdx_comm_create(uid, self.world, self.rank, self.device, C.byref(handle)))
        self.handle = handle

    def all_gather_counts(self, counts):
        """counts (C, 3) int64 CUDA tensor -> (world, C, 3), stream-ordered on the current stream."""
        torch = self._sx._torch()
        assert counts.is_cuda and counts.dtype == torch.int64 and counts.dim() == 2 and counts.shape[1] == 3
        counts = counts.contiguous()
        out = torch.empty((self.world,) + tuple(counts.shape), dtype=torch.int64, device=counts.device)
        self._sx.check(self._sx.lib.ssdx_allgather_counts(self.handle, counts.data_ptr(), out.data_ptr(),
                                                           counts.shape[0], self._sx.stream_ptr(self.device)))
        return out

    def all_gather_eval(self, counts, tp, confidence, label):
        """Everything the global AP tail needs in one NCCL launch (`ssdx_allgather_eval`): counts (C, 3) int64 ->
        (world, C, 3); tp uint8 / confidence float32 / label int32, each with the same number of elements on every
        rank -> flat (world * n) arrays, rank-major."""
        torch = self._sx._torch()
        counts, tp = counts.contiguous(), tp.contiguous().view(-1)
        confidence, label = confidence.contiguous().view(-1), label.contiguous().view(-1)
        assert tp.dtype == torch.uint8 and confidence.dtype == torch.float32 and label.dtype == torch.int32
        n = tp.numel()
        assert confidence.numel() == n and label.numel() == n
        dev = counts.device
        g_counts = torch.empty((self.world,) + tuple(counts.shape), dtype=torch.int64, device=dev)
        g_tp = torch.empty(self.world * n, dtype=torch.uint8, device=dev)
        g_conf = torch.empty(self.world * n, dtype=torch.float32, device=dev)
        g_label = torch.empty(self.world * n, dtype=torch.int32, device=dev)
        self._sx.check(self._sx.lib.ssdx_allgather_eval(
            self.handle, counts.data_ptr(), g_counts.data_ptr(), counts.shape[0], tp.data_ptr(), confidence.data_ptr(),
            label.data_ptr(), n, g_tp.data_ptr(), g_conf.data_ptr(), g_label.data_ptr(), self._sx.stream_ptr(self.device)))
        return g_counts, g_tp, g_conf, g_label

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self._sx.lib.ssdx_comm_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def all_gather_counts(counts, process_group=None):
    """counts: (C, 3) int64 tensor on this rank -> (world_size, C, 3) tensor with every rank's table.

    CUDA tensors: `ssdx_allgather_counts` over a CountsCommunicator cached per (process group, device) - the
    first call on a group is collective (it creates the communicator).  CPU tensors: torch.distributed (gloo)."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return counts.unsqueeze(0)
    if counts.is_cuda:
        key = (id(process_group) if process_group is not None else 0, counts.device.index)
        with _comms_lock:
            comm = _comms.get(key)
            if comm is None:
                comm = _comms[key] = CountsCommunicator(process_group, counts.device.index)
        return comm.all_gather_counts(counts)
    world = dist.get_world_size(process_group)
    flat = counts.contiguous().view(-1)
    out = torch.empty(world * flat.numel(), dtype=counts.dtype, device=counts.device)  # concatenated form
    dist.all_gather_into_tensor(out, flat, group=process_group)
    return out.view((world,) + tuple(counts.shape))


def close_communicators():
    """Destroy the cached communicators (call before torch.distributed.destroy_process_group)."""
    with _comms_lock:
        for comm in _comms.values():
            comm.close()
        _comms.clear()
