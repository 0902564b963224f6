"""Phase timeline of the fused attention kernels (clock64 stamps written by the first CTAs):
python tools/flash_trace.py [fwd|bwd] [L] [S] -> prints per-CTA phase durations in cycles."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pointcloudmatters_b200 import kernels as K  # noqa: E402
from pointcloudmatters_b200._lib import lib, ptr  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
L = int(sys.argv[2]) if len(sys.argv) > 2 else 515
S = int(sys.argv[3]) if len(sys.argv) > 3 else 515
B, nh = 64, 8
Z, E = B * nh, nh * 64
q = torch.randn(Z * L, 64, device="cuda").bfloat16()
k = torch.randn(Z * S, 64, device="cuda").bfloat16()
v = torch.randn(Z * S, 64, device="cuda").bfloat16()
do = torch.randn(Z * L, 64, device="cuda").bfloat16()
sb = torch.tensor([1234567], dtype=torch.int64, device="cuda")
buf = torch.empty(L * B, E, dtype=torch.bfloat16, device="cuda")
kvb = torch.empty(S * B, 2 * E, dtype=torch.bfloat16, device="cuda")
NCTA = 2048
trace = torch.zeros(NCTA, 64, dtype=torch.int64, device="cuda")
for _ in range(2):
    O, lse = K.flash_attn_fwd(q, k, v, B, nh, L, S, None, 0.125, 0.1, sb, 77)
    K.flash_attn_bwd(q, k, v, O, do, lse, B, nh, L, S, None, 0.125, 0.1, sb, 77, buf, kvb[:, :E], kvb[:, E:])
torch.cuda.synchronize()
lib.pcm_flash_attn_debug_trace(ptr(trace), NCTA)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
if which == "fwd":
    K.flash_attn_fwd(q, k, v, B, nh, L, S, None, 0.125, 0.1, sb, 77)
else:
    lib.pcm_flash_attn_debug_trace(None, 0)
    O, lse = K.flash_attn_fwd(q, k, v, B, nh, L, S, None, 0.125, 0.1, sb, 77)
    torch.cuda.synchronize()
    lib.pcm_flash_attn_debug_trace(ptr(trace), NCTA)
    e0.record()
    K.flash_attn_bwd(q, k, v, O, do, lse, B, nh, L, S, None, 0.125, 0.1, sb, 77, buf, kvb[:, :E], kvb[:, E:])
e1.record()
torch.cuda.synchronize()
lib.pcm_flash_attn_debug_trace(None, 0)
print(f"{which} L={L} S={S}: {e0.elapsed_time(e1) * 1e3:.1f} us (traced launch incl. helpers)")
t = trace.cpu()
names = {0: "start", 1: "setup", 62: "sm_done", 63: "end"}
for j in range(8):
    names[8 + 2 * j] = f"mma_A{j}"
    names[9 + 2 * j] = f"mma_B{j}"
for j in range(7):
    for q_, nm in enumerate(("got_in", "p1", "acc", "p2", "arr")):
        names[24 + 5 * j + q_] = f"sm{j}_{nm}"
names[60], names[61] = "fin_wait", "fin_dq"
for cta in (0, 1, 2, 3, 4, 700, 701, 1500):
    if cta >= NCTA:
        continue
    row = t[cta]
    t0 = int(row[0])
    ev = sorted((int(row[s]) - t0, names.get(s, str(s))) for s in range(64) if int(row[s]) != 0)
    print(f"CTA {cta}: " + " ".join(f"{n}@{c}" for c, n in ev))
life = (t[:, 63] - t[:, 0]).float()
ok = t[:, 63] != 0
print("CTA lifetime cycles: mean %.0f min %.0f max %.0f (n=%d)" % (life[ok].mean(), life[ok].min(), life[ok].max(), int(ok.sum())))
setup = (t[:, 1] - t[:, 0]).float()[ok]
print("setup cycles: mean %.0f max %.0f" % (setup.mean(), setup.max()))
