"""Multi-GPU plumbing: one process per GPU, utterances sharded, statistics all-reduced.

The reference is single-process; its embarrassingly parallel loops are the per-utterance
E-step (custom_hmm.py:422-439) and per-utterance x per-model decode (decoder.py:42-47,58-60).  The only
cross-utterance reduction is the accumulator block custom_hmm.py:417-419,434-439 (+ the sums in update_B
:372-386): that is the ONE all-reduce per Baum-Welch iteration.  Viterbi needs no exchange at all.
"""
from __future__ import annotations

import os

import numpy as np


def shard_bounds(offsets_host: np.ndarray, world: int):
    """Contiguous utterance ranges per rank, balanced by FRAMES (ragged batches), returned as a list of
    (u_begin, u_end).  Every utterance lands in exactly one shard; shards are contiguous and ordered."""
    offs = np.asarray(offsets_host, dtype=np.int64)
    B = len(offs) - 1
    total = int(offs[-1])
    bounds, start = [], 0
    for r in range(world):
        if r == world - 1:
            end = B
        else:
            target = total * (r + 1) / world
            end = int(np.searchsorted(offs, target, side="left"))
            end = min(max(end, start), B)
        bounds.append((start, end))
        start = end
    return bounds


def _parse_cpulist(txt: str):
    cpus = []
    for part in txt.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.extend(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa_node(local_rank: int):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs: the PCI device's numa_node and the node's
    cpulist), so pinned host buffers allocated afterwards are node-local: with one process per GPU, eight concurrent
    host-to-device streams otherwise contend for one socket's memory.  Returns the node, or None when the topology is
    not exposed (then nothing is changed)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        dev = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{dev}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


class Dist:
    """Thin wrapper over torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, backend=None, init=True):
        import torch
        import torch.distributed as td

        self.td = td
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if init and self.world > 1 and not td.is_initialized():
            backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
            td.init_process_group(backend=backend, rank=self.rank, world_size=self.world)

        self._comm = None       # sapr_comm handle (C ABI, NCCL), created on the first device all-reduce
        self._comm_ctx = None

    def _native_comm(self, ctx):
        """The library's own communicator (sapr_comm_init_rank): rank 0 draws the NCCL unique id, torch.distributed is
        only the side channel that hands the 128 bytes to the other ranks."""
        if self._comm is None:
            import ctypes as C
            import torch
            uid = torch.zeros(128, dtype=torch.uint8)
            if self.rank == 0:
                buf = (C.c_ubyte * 128)()
                rc = ctx.lib.sapr_comm_unique_id(C.cast(buf, C.c_void_p))
                if rc != 0:
                    raise RuntimeError("sapr_comm_unique_id failed: libnccl.so.2 not loadable")
                uid = torch.tensor(list(buf), dtype=torch.uint8)
            dev_uid = uid.to(ctx.device) if self.td.get_backend() == "nccl" else uid
            self.td.broadcast(dev_uid, src=0)
            host = bytes(dev_uid.cpu().tolist())
            h = C.c_void_p()
            ctx.check(ctx.lib.sapr_comm_init_rank(ctx.h, C.c_char_p(host), self.rank, self.world, C.byref(h)))
            self._comm, self._comm_ctx = h, ctx
        return self._comm

    def allreduce_(self, t, ctx=None):
        """In-place sum over ranks of one packed float64 buffer (57 KB at cfg 3: latency bound).  Device float64 buffers go
        through the C ABI (sapr_stats_allreduce: ncclAllReduce enqueued on the context's stream, no host sync); host
        tensors (the gloo CPU tests) and other dtypes through torch.distributed."""
        if self.world > 1:
            import torch
            if t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and os.environ.get("SAPR_NATIVE_COMM", "1") != "0":
                if ctx is None:
                    from . import _lib
                    ctx = _lib.default_context(t.device.index)
                comm = self._native_comm(ctx)
                ctx.check(ctx.lib.sapr_stats_allreduce(ctx.h, comm, t.data_ptr(), t.numel()))
            else:
                self.td.all_reduce(t, op=self.td.ReduceOp.SUM)
        return t

    def max_(self, t):
        if self.world > 1:
            self.td.all_reduce(t, op=self.td.ReduceOp.MAX)
        return t

    def barrier(self):
        if self.world > 1:
            self.td.barrier()

    def shutdown(self):
        if self._comm is not None:
            import torch
            torch.cuda.synchronize()
            self._comm_ctx.lib.sapr_comm_destroy(self._comm)
            self._comm = None
        if self.world > 1 and self.td.is_initialized():
            self.td.destroy_process_group()
