"""Where does the run-to-run spread of the parameter gradients come from?  Two runs of the same step on fresh models; compares the
forward maps, the gradients the loss hands to the network (d loss / d depth, d loss / d logits) and the parameter gradients."""
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
    rec = {}
    feats = m.depth_backbone(cuda[("image", 0)])
    outs = m.head.forward_depth(feats, cuda["P2"])
    for s in topo.scales:
        rec[f"depth/{s}"] = outs[("depth", s, s)].detach().clone()
        rec[f"logits/{s}"] = outs[("logits", s)].detach().clone()
        outs[("depth", s, s)].register_hook(lambda g, s=s: rec.__setitem__(f"g_depth/{s}", g.detach().clone()))
        outs[("logits", s)].register_hook(lambda g, s=s: rec.__setitem__(f"g_logits/{s}", g.detach().clone()))
    for f in topo.frame_ids[1:]:
        outs[("cam_T_cam", f)] = cuda[("relative_pose", f)]
    ret = m.head.loss(outs, cuda)
    ret["loss"].mean().backward()
    torch.cuda.synchronize()
    for k, p in m.named_parameters():
        if p.grad is not None:
            rec["param/" + k] = p.grad.detach().clone()
    return rec


a, b = run(), run()
def rel(x, y):
    return float((x.double() - y.double()).norm() / (x.double().norm() + 1e-300))
for k in sorted(a):
    if not k.startswith("param/"):
        print(f"{k:14s} rel {rel(a[k], b[k]):.3e}  maxabs {float((a[k]-b[k]).abs().max()):.3e}  nonzero {float((a[k]!=0).float().mean()):.4f}")
rows = sorted(((rel(a[k], b[k]), k) for k in a if k.startswith("param/") and float(a[k].norm()) > 0), reverse=True)
print("params worst", rows[:3], "median", rows[len(rows)//2])
for k in ("param/head.depth_decoder.decoder.13.weight", "param/head.depth_decoder.decoder.9.sequence.0.weight", "param/head.depth_decoder.decoder.0.sequence.0.weight", "param/depth_backbone.layer4.1.conv2.weight", "param/depth_backbone.conv1.weight"):
    if k in a:
        print(k, f"{rel(a[k], b[k]):.3e}")
