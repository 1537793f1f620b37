"""Run the same training step twice on fresh models (one GPU) and report the run-to-run spread per parameter gradient."""
import os, sys
sys.path[:0] = [os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"), os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")]
import torch
from oracle import fsnet_oracle as O
from helpers import build_model

H, W, B = int(os.environ.get("H", 64)), int(os.environ.get("W", 128)), int(os.environ.get("B", 2))
topo = O.Topology(height=H, width=W)
data = O.synthetic_batch(B, H, W, 78, topo.frame_ids)
noise = O.tie_break_noise(B, H, W, topo.scales, 0)
cuda = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in data.items()}


def run():
    m = build_model(topo).cuda()
    m.head.tie_break_noise = {s: n.cuda() for s, n in noise.items()}
    outs = {}
    ret = m(dict(cuda), dict(is_training=True, epoch_num=0, global_step=0))
    ret["loss"].mean().backward()
    torch.cuda.synchronize()
    return {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}, float(ret["loss"])


a, la = run()
b, lb = run()
gmax = max(float(v.norm()) for v in a.values())
rows = []
for k in a:
    if float(a[k].norm()) > 1e-7 * gmax:
        rows.append((float((a[k] - b[k]).norm() / a[k].norm()), k))
rows.sort(reverse=True)
print(f"loss {la!r} vs {lb!r}; {len(rows)} tensors; worst run-to-run spreads:")
for e, k in rows[:8]:
    print(f"  {e:.3e}  {k}")
print("median", rows[len(rows) // 2])
bitwise = sum(1 for k in a if torch.equal(a[k], b[k]))
print(f"bitwise identical tensors: {bitwise} of {len(a)}")
