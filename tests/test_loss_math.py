"""The closed-form loss arithmetic the CUDA kernels implement (tests/loss_math_ref.py) vs autograd
through the oracle (which is itself pinned to the reference)."""
import torch
import torch.nn.functional as F
import pytest

from oracle import fsnet_oracle as O
import loss_math_ref as M
from test_oracle_golden import LOSS_CASES, build_loss_case, rel


@pytest.mark.parametrize("name", ["loss_a", "loss_b", "loss_c", "loss_mm"])
def test_closed_form_matches_autograd(name):
    case = LOSS_CASES[name]
    topo = case["topo"]
    data, outputs, noise = build_loss_case(**case)
    B, H, W = data[("original_image", 0)].shape[0], topo.height, topo.width
    for v in outputs.values():
        v.requires_grad_(True)
    cam_T = {f: data[("relative_pose", f)].clone().requires_grad_(True) for f in topo.frame_ids[1:]}
    ret = O.loss_chain(outputs, data, cam_T, topo, noise, keep=True)
    ret["loss"].backward()

    target = data[("original_image", 0)]
    srcs = [data[("original_image", f)] for f in topo.frame_ids[1:]]
    mask = data.get("patched_mask")
    S = len(topo.scales)
    any_flip = False
    for s in topo.scales:
        cams, Ks = [], []
        for f in topo.frame_ids[1:]:
            invK, P, K4 = M.camera(data["P2"], data[("relative_pose", f)])
            cams.append((invK, P))
            Ks.append(K4)
        if "motion_mask" in data:
            ident = None
        else:
            ident = torch.stack([M.photometric(src, target)[0] for src in srcs], 1) + noise[s] * 0.00001
        loss, gdepth, gP, idx = M.scale_forward_backward(
            outputs[("depth", s, s)].detach(), target, srcs, cams, mask, ident, topo.overlapped_mask, 1.0 / S,
            motion_mask=data.get("motion_mask"))
        color = target if s == 0 else F.adaptive_avg_pool2d(target, outputs[("disp", s)].shape[-2:])
        sm, gdisp = M.smooth_forward_backward(outputs[("disp", s)].detach(), color, 1e-5 / (2 ** s))
        assert abs(float(sm) - float(ret["loss_dict"][f"smooth_loss/{s}"])) < 1e-5 * float(sm)
        tot = float(loss) + float(sm)
        assert abs(tot - float(ret["loss_dict"][f"loss/{s}"])) < 1e-5 * tot
        assert rel(gdisp / S, outputs[("disp", s)].grad) < 1e-4
        if ident is not None:
            flips = float((idx != ret["aux"][("idxs", s)]).float().mean())
            assert flips < 2e-3, flips      # near-ties may flip with rounding
            any_flip = any_flip or flips > 0
        # a flipped arg-min moves a whole 3x3 SSIM footprint of gradient: widen on flips
        tol = 2e-3 if (ident is None or flips == 0) else 3e-2
        err = rel(gdepth, outputs[("depth", s, s)].grad)
        assert err < tol, (s, err, flips)
    # pose gradient: dL/dT = K4[:3,:]^T dL/dP accumulated over scales -- check the last scale contribution sums
    # (autograd accumulated all scales; redo all scales explicitly)
    for fi, f in enumerate(topo.frame_ids[1:]):
        acc = torch.zeros(B, 4, 4)
        for s in topo.scales:
            cams = [M.camera(data["P2"], data[("relative_pose", ff)])[:2] for ff in topo.frame_ids[1:]]
            ident = None if "motion_mask" in data else torch.stack([M.photometric(src, target)[0] for src in srcs], 1) + noise[s] * 0.00001
            _, _, gP, _ = M.scale_forward_backward(outputs[("depth", s, s)].detach(), target, srcs, cams, mask, ident,
                                                   topo.overlapped_mask, 1.0 / S, motion_mask=data.get("motion_mask"))
            K4 = M.camera(data["P2"], data[("relative_pose", f)])[2]
            acc += torch.matmul(K4[:, :3, :].transpose(1, 2), gP[fi])
        # pose gradients are a heavily cancelling sum over pixels: one flipped arg-min shows up at the % level
        assert rel(acc, cam_T[f].grad) < (0.15 if any_flip else 5e-3), rel(acc, cam_T[f].grad)
