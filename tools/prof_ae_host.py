"""Where the time of a host-to-host AE encode goes (bring-up aid): raw copy rates of the box, then the call."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latent_rag_b200 as lrb

m = 262_144
gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
ae = lrb.load_autoencoder("cae", os.path.join(gold, "ae_weights_cae.npz"))
x = torch.randn((m, 384))
xp = x.pin_memory()
xd = torch.empty((m, 384), device="cuda")
zd = torch.empty((m, 64), device="cuda")
zp = torch.empty((m, 64)).pin_memory()

def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return ts

print("H2D 402 MB pinned   ms", [f"{v:.1f}" for v in t(lambda: xd.copy_(xp, non_blocking=True))])
print("H2D 402 MB pageable ms", [f"{v:.1f}" for v in t(lambda: xd.copy_(x, non_blocking=True))])
print("D2H  67 MB pinned   ms", [f"{v:.1f}" for v in t(lambda: zp.copy_(zd, non_blocking=True))])
print("pinned alloc 64 MiB ms", [f"{v:.1f}" for v in t(lambda: torch.empty((m, 64), pin_memory=True))])
for prec in ("fp32", "bf16"):
    ae.set_precision(prec)
    print(prec, "encode device->device ms", [f"{v:.2f}" for v in t(lambda: ae.encode(xd))])
    print(prec, "encode pinned->host   ms", [f"{v:.1f}" for v in t(lambda: ae.encode(xp), 5)])
    print(prec, "encode pageable->host ms", [f"{v:.1f}" for v in t(lambda: ae.encode(x), 3)])
