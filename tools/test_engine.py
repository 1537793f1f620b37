"""Hardware check of the tcgen05 whole-network path against the torch (cuDNN fp32) back-end of the same modules."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

from fsnet_b200.networks import ops
from oracle import fsnet_oracle as O
from helpers import build_model

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def run(topo, B, backend):
    ops.set_backend(backend)
    data = {k: v.cuda() for k, v in O.synthetic_batch(B, topo.height, topo.width, 1234, topo.frame_ids).items()}
    model = build_model(topo).cuda()
    model.head.tie_break_noise = O.tie_break_noise(B, topo.height, topo.width, topo.scales, 0)
    out = model(data, dict(is_training=True, epoch_num=0, global_step=0))
    out["loss"].mean().backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    bufs = {k: v.detach().clone() for k, v in model.state_dict().items() if "running" in k}
    model.eval()
    with torch.no_grad():
        pred = model(data, dict(is_training=False))["depth"]
    return out, grads, bufs, pred


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    topo = {"tiny": O.Topology(height=64, width=128), "pose": O.Topology(height=64, width=128, posenet=True, overlapped_mask=False),
            "r50": O.Topology(height=64, width=96, depth=50, base_fx=40.0), "cfg1": O.Topology(height=128, width=416, scales=(0,)),
            "sig": O.Topology(height=64, width=96, multi_channel=False, n_bins=1, min_depth=0.1)}[name]
    ref, gref, bref, pref = run(topo, 2, "torch")
    got, ggot, bgot, pgot = run(topo, 2, "tc")
    ok = True
    print("loss", float(ref["loss"]), float(got["loss"]))
    for k in ref["loss_dict"]:
        e = abs(float(got["loss_dict"][k]) / float(ref["loss_dict"][k]) - 1)
        ok &= e < 1e-3
        print(f"  {k}: rel {e:.2e}")
    e = rel(pgot, pref)
    ok &= e < 2e-3
    print(f"eval-mode depth rel {e:.2e}")
    worst = sorted(((rel(bgot[k], bref[k]), k) for k in bref), reverse=True)[:3]
    print("running stats worst:", worst)
    ok &= worst[0][0] < 1e-3
    missing = [k for k in gref if k not in ggot]
    print("missing grads:", missing[:5], len(missing))
    ok &= not missing
    errs = sorted(((rel(ggot[k], gref[k]), k) for k in gref if k in ggot and float(gref[k].norm()) > 1e-10), reverse=True)
    print("param grad rel err: worst", errs[:6])
    print("param grad rel err: median", errs[len(errs) // 2])
    ok &= errs[0][0] < 0.2 and errs[len(errs) // 2][0] < 3e-2
    print("ENGINE OK" if ok else "ENGINE FAIL")
    sys.exit(0 if ok else 1)
