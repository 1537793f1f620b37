"""Debug: per-pixel comparison of the fisheye loss gradient with the oracle (GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import fsnet_oracle as O
from test_oracle_golden import LOSS_CASES, build_loss_case, rel
from test_loss_gpu import run_gpu_loss

for name in ("loss_fe", "loss_fe_nomask"):
    case = LOSS_CASES[name]
    topo = case["topo"]
    data, outputs, noise = build_loss_case(**case)
    for v in outputs.values():
        v.requires_grad_(True)
    cam_T = {f: data[("relative_pose", f)].clone().requires_grad_(True) for f in topo.frame_ids[1:]}
    ref = O.loss_chain(outputs, data, cam_T, topo, noise, keep=True)
    ref["loss"].backward()
    total, stats, depths, disps, T = run_gpu_loss(topo, data, outputs, noise)
    lut = O.mei_lut_batch(data["P2"], data["calib_meta"], topo.height, topo.width)
    for i, s in enumerate(topo.scales):
        a, b = depths[i].grad.cpu(), outputs[("depth", s, s)].grad
        d = (a - b).abs()
        print(name, "scale", s, "rel", rel(a, b), "max abs diff", float(d.max()), "ref max", float(b.abs().max()))
        if s == 0:
            flat = d.flatten().topk(8)
            for v, idx in zip(flat.values, flat.indices):
                bb, rem = divmod(int(idx), topo.height * topo.width)
                y, x = divmod(rem, topo.width)
                print("   b", bb, "y", y, "x", x, "gpu", float(a[bb, 0, y, x]), "ref", float(b[bb, 0, y, x]), "lutmask", float(lut[bb, 3, y, x]),
                      "idx", int(ref["aux"][("idxs", 0)][bb, y, x]), "depth", float(outputs[("depth", 0, 0)][bb, 0, y, x]))
    for fi, f in enumerate(topo.frame_ids[1:]):
        print("  grad_T", f, rel(T[fi].grad.cpu()[:, :3], cam_T[f].grad[:, :3]))
