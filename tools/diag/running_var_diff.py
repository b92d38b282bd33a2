"""Diag: the buffers of tests/test_training_parity_gpu.py after 5 steps -- worst relative running_var deviation per buffer."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_training_parity_gpu as T
import msfwsi_b200 as M
B, steps, lr = 32, int(os.environ.get("STEPS", "5")), 1e-4
mine, ref = T._pair(3)
opt_m = M.FusedAdam(T._groups(mine), lr=lr); opt_r = torch.optim.Adam(T._groups(ref), lr=lr)
cf, tf, rev = T._features(B)
for _ in range(steps):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss_m = M.ssl_loss(mine.heads(cf[0], cf[1], tf[0], tf[1], rev), T.W, mode="cosine")
        loss_r = T.R.ref_ssl_loss(ref.heads(cf[0], cf[1], tf[0], tf[1], rev), T.W)
    for opt, loss in ((opt_m, loss_m), (opt_r, loss_r)):
        opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
br = dict(ref.named_buffers())
rows = []
for n, b in mine.named_buffers():
    if n.endswith("running_var"):
        rel = ((b - br[n]).abs() / br[n].abs().clamp_min(1e-4))
        rows.append((float(rel.max()), float(rel.mean()), n, int(rel.argmax()), float(b.flatten()[rel.argmax()]), float(br[n].flatten()[rel.argmax()])))
rows.sort(reverse=True)
for r in rows[:12]:
    print("max rel %.4f mean rel %.5f  %s  idx %d mine %.5f ref %.5f" % r)
