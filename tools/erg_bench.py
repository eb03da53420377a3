"""BASELINE cfg 4 timing: N = 256 fully connected states, D = 39, T = 1000, one 128-utterance tile per SM.
Prints per-kernel CUDA-event times of sapr_ergodic_score and the float64 one-CTA-per-utterance kernel on a sample."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from sapr_b200._lib import ptr, EMIT_DIAG, TOPO_DENSE
from sapr_b200.engine import WordModels

S, D, T = 256, 39, int(sys.argv[2]) if len(sys.argv) > 2 else 1000
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128
rng = np.random.default_rng(0)
means = 2.0 * rng.standard_normal((S, D)); var = rng.uniform(0.5, 1.5, (S, D)) ** 2
tm = rng.dirichlet(np.ones(S), size=S); sp = rng.dirichlet(np.ones(S))
m = WordModels(1, S, D, EMIT_DIAG, TOPO_DENSE)
m.set(means, var, tm, sp)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
mt = torch.tensor(means, dtype=torch.float32, device=dev); sd = torch.tensor(np.sqrt(var), dtype=torch.float32, device=dev)
st = torch.randint(0, S, (B * T,), device=dev, generator=g)
X = mt[st] + sd[st] * torch.randn(B * T, D, device=dev, generator=g)
offsets = torch.arange(0, B + 1, device=dev, dtype=torch.int64) * T
lp = torch.zeros(B, dtype=torch.float64, device=dev)
def run():
    m.ctx.check(m.lib.sapr_ergodic_score(m.ctx.h, m.h, 0, ptr(X), D, ptr(offsets), B, T, ptr(lp)))
for _ in range(2): run()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
n = 3
ev[0].record()
for _ in range(n): run()
ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / n
m.ctx.profile(True)
run()
e_ms, e_n = m.ctx.profile_read(4)
f_ms, f_n = m.ctx.profile_read(5)
m.ctx.profile(False)
lf_gb = B * T * S * 4 / 1e9
print(f"  emission kernel {e_ms:.3f} ms ({e_n} launches, {lf_gb / e_ms * 1e3:.0f} GB/s written)   "
      f"forward kernel {f_ms:.3f} ms ({f_n} launches, {lf_gb / f_ms * 1e3:.0f} GB/s read, "
      f"{4 * B * T * S * S / f_ms * 1e3 / 1e12:.0f} TFLOP/s)")
upd = B * T * S * S
print(f"ergodic tc: B={B} T={T} S={S}: {ms:.3f} ms/call  {upd / ms * 1e3:.3e} state-pair updates/s  "
      f"{4 * upd / ms * 1e3 / 1e12:.1f} fp16 TFLOP/s (hi+lo)")
# float64 reference kernel on a sample
Bs = min(B, 296)
lp64 = torch.zeros(Bs, dtype=torch.float64, device=dev)
def run64():
    m.ctx.check(m.lib.sapr_hl_score(m.ctx.h, m.h, 0, ptr(X), D, ptr(offsets), Bs, Bs * T, ptr(lp64)))
run64(); torch.cuda.synchronize()
ev[0].record(); run64(); ev[1].record(); torch.cuda.synchronize()
ms64 = ev[0].elapsed_time(ev[1])
print(f"float64 cta kernel: B={Bs}: {ms64:.3f} ms  {Bs * T * S * S / ms64 * 1e3:.3e} updates/s")
err = (lp[:Bs] - lp64).abs()
print("max |dlogP|", err.max().item(), "rel", (err / lp64.abs()).max().item(), "mean logP", lp64.mean().item())
