"""Multi-GPU check (run under torchrun, one rank per GPU): the C-ABI collective sapr_stats_allreduce (NCCL behind the
library) against torch.distributed's all_reduce on the same buffers, and a sharded Baum-Welch iteration through it against the
unsharded one."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from sapr_b200 import _lib, engine, synth
from sapr_b200.dist import Dist, shard_bounds

d = Dist()
torch.cuda.set_device(d.local_rank)
dev = torch.device("cuda", d.local_rank)
ctx = _lib.default_context()
g = torch.Generator(device=dev); g.manual_seed(100 + d.rank)
a = torch.randn(7139 * 11 + 11, dtype=torch.float64, device=dev, generator=g)
b = a.clone()
d.allreduce_(a, ctx)                                   # native: sapr_stats_allreduce
d.td.all_reduce(b, op=d.td.ReduceOp.SUM)               # torch.distributed
torch.cuda.synchronize()
assert d._comm is not None, "native communicator was not used"
err = float((a - b).abs().max().item())
assert err <= 1e-12 * float(b.abs().max().item()), err
# sharded E-step + native all-reduce == unsharded E-step
feats, labels, mu, sd = synth.make_corpus(512, 11, 8, 39, 60, 60, seed=5)
A, means, var = synth.truth_models(mu, sd, 0.9)
m = engine.WordModels(11, 8, 39); m.set(means, var, A)
full = engine.PackedBatch.from_features(feats)
lab = torch.tensor(labels, dtype=torch.int32, device=dev)
st_full, _, _ = m.estep(full, lab, None, engine.FP64)
lo, hi = shard_bounds(full.offsets_host, d.world)[d.rank]
part = engine.PackedBatch.from_features(feats[lo:hi])
st, _, _ = m.estep(part, lab[lo:hi], None, engine.FP64)
d.allreduce_(st, ctx)
torch.cuda.synchronize()
rel = float(((st - st_full).abs() / (1e-9 + st_full.abs().clamp(min=1.0))).max().item())
assert rel < 1e-10, rel
if d.rank == 0:
    print("comm_check ok: world", d.world, "allreduce max err", err, "sharded-vs-full stats rel", rel)
d.shutdown()
