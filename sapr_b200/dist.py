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

    def allreduce_(self, t):
        """In-place sum over ranks of one packed float64 buffer (57 KB at cfg 3: latency bound)."""
        if self.world > 1:
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
        if self.world > 1 and self.td.is_initialized():
            self.td.destroy_process_group()
