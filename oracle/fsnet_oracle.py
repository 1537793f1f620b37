"""Torch-fp32 CPU restatement of the FSNet training step (encoder, decoder, PoseNet, loss chain).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Written from the reference's behaviour,
function by function, as plain functional torch on a flat ``state_dict`` (reference key layout,
SURVEY.md section 8(b)).  All ``file:line`` citations are relative to the reference checkout.

Pinned by ``tests/golden`` (outputs of the reference itself, see ``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

NUM_CH_DEC = (16, 32, 64, 128, 256)          # monodepth/networks/models/heads/depth_encoder.py:28
RESNET_LAYERS = {18: (2, 2, 2, 2), 34: (3, 4, 6, 3), 50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}
RESNET_PLANES = (64, 128, 256, 512)          # vision_base/networks/models/backbone/resnet.py:94


# --------------------------------------------------------------------------------------------
# topology description
# --------------------------------------------------------------------------------------------
@dataclass
class Topology:
    """What the in-scope configs vary (configs/kitti_wpose_example:174-215)."""
    depth: int = 18                      # ResNet depth of the depth encoder
    n_bins: int = 16                     # num_output_channels of the decoder
    scales: Tuple[int, ...] = (0, 1, 2, 3)
    multi_channel: bool = True           # MultiChannelDepthDecoder vs DepthDecoder
    min_depth: float = 0.5
    max_depth: float = 100.0
    base_fx: Optional[float] = None
    use_skips: bool = True
    posenet: bool = False                # MonoDepthMeta (+ R18 6-ch PoseNet + PoseDecoder)
    pose_depth: int = 18
    height: int = 192
    width: int = 640
    overlapped_mask: bool = True
    frame_ids: Tuple[int, ...] = (0, 1, -1)
    fisheye: bool = False                # FishEyeDecoder: MEI camera, the network output is the ray norm
    norm_eval: bool = False              # ResNet(norm_eval=True): backbone BatchNorms stay in eval mode while training (resnet.py:169-197)
    frozen_stages: int = -1              # ResNet(frozen_stages=k): stem + layer1..k in eval mode, no gradients (resnet.py:176-190)
    distill: bool = False                # DistillWPoseMeta: frozen eval-mode teacher + MultiChannelDepthDecoderUncertain student
    distill_weight: float = 0.3          # distillation_loss_weight (configs/distill_kitti_example:218)
    uncertain_distill: bool = True       # is_uncertain_distill (configs/distill_kitti_example:219)
    teacher_depth: int = 18
    teacher_bins: int = 16

    @property
    def bottleneck(self) -> bool:
        return self.depth >= 50

    @property
    def num_ch_enc(self) -> Tuple[int, ...]:
        e = 4 if self.bottleneck else 1
        return (64, 64 * e, 128 * e, 256 * e, 512 * e)


def _bn_entries(prefix: str, c: int) -> List[Tuple[str, Tuple[int, ...], str]]:
    return [(prefix + ".weight", (c,), "bn_w"), (prefix + ".bias", (c,), "bn_b"),
            (prefix + ".running_mean", (c,), "zeros"), (prefix + ".running_var", (c,), "ones"),
            (prefix + ".num_batches_tracked", (), "count")]


def resnet_param_specs(prefix: str, depth: int, num_input_images: int = 1):
    """Key/shape list of vision_base/networks/models/backbone/resnet.py:96-167."""
    specs = [(prefix + "conv1.weight", (64, 3 * num_input_images, 7, 7), "conv")]
    specs += _bn_entries(prefix + "bn1", 64)
    bottleneck = depth >= 50
    exp = 4 if bottleneck else 1
    inplanes = 64
    for li, (planes, nblocks) in enumerate(zip(RESNET_PLANES, RESNET_LAYERS[depth])):
        stride = 1 if li == 0 else 2
        for bi in range(nblocks):
            p = f"{prefix}layer{li + 1}.{bi}."
            s = stride if bi == 0 else 1
            if bottleneck:
                specs.append((p + "conv1.weight", (planes, inplanes, 1, 1), "conv"))
                specs += _bn_entries(p + "bn1", planes)
                specs.append((p + "conv2.weight", (planes, planes, 3, 3), "conv"))
                specs += _bn_entries(p + "bn2", planes)
                specs.append((p + "conv3.weight", (planes * 4, planes, 1, 1), "conv"))
                specs += _bn_entries(p + "bn3", planes * 4)
            else:
                specs.append((p + "conv1.weight", (planes, inplanes, 3, 3), "conv"))
                specs += _bn_entries(p + "bn1", planes)
                specs.append((p + "conv2.weight", (planes, planes, 3, 3), "conv"))
                specs += _bn_entries(p + "bn2", planes)
            if bi == 0 and (s != 1 or inplanes != planes * exp):
                specs.append((p + "downsample.0.weight", (planes * exp, inplanes, 1, 1), "conv"))
                specs += _bn_entries(p + "downsample.1", planes * exp)
            inplanes = planes * exp
    return specs


def decoder_param_specs(prefix: str, topo: Topology, uncertain: bool = False):
    """monodepth/networks/models/heads/depth_encoder.py:45-66 (creation order = ModuleList index)."""
    specs = []
    enc = topo.num_ch_enc
    k = 0
    for i in range(4, -1, -1):
        cin = enc[-1] if i == 4 else NUM_CH_DEC[i + 1]
        cout = NUM_CH_DEC[i]
        for j in range(2):
            if j == 1:
                cin = NUM_CH_DEC[i] + (enc[i - 1] if (topo.use_skips and i > 0) else 0)
            p = f"{prefix}decoder.{k}.sequence."
            specs.append((p + "0.weight", (cout, cin, 3, 3), "conv"))
            specs.append((p + "0.bias", (cout,), "bias"))
            specs += _bn_entries(p + "1", cout)
            k += 1
    for s in topo.scales:
        specs.append((f"{prefix}decoder.{k}.weight", (topo.n_bins, NUM_CH_DEC[s], 3, 3), "conv"))
        specs.append((f"{prefix}decoder.{k}.bias", (topo.n_bins,), "bias"))
        k += 1
    if uncertain:       # MultiChannelDepthDecoderUncertain._init_layers, depth_encoder.py:164-168: one 1-channel head per scale
        for s in topo.scales:
            specs.append((f"{prefix}decoder.{k}.weight", (1, NUM_CH_DEC[s], 3, 3), "conv"))
            specs.append((f"{prefix}decoder.{k}.bias", (1,), "bias"))
            k += 1
    return specs


def pose_decoder_param_specs(prefix: str, c_last: int, num_input_features: int = 1, n_pred: int = 2):
    """monodepth/networks/models/heads/pose_decoder.py:17-24."""
    shapes = [(256, c_last, 1, 1), (256, num_input_features * 256, 3, 3), (256, 256, 3, 3), (6 * n_pred, 256, 1, 1)]
    specs = []
    for i, sh in enumerate(shapes):
        specs.append((f"{prefix}net.{i}.weight", sh, "conv"))
        specs.append((f"{prefix}net.{i}.bias", (sh[0],), "bias"))
    return specs


def depth_bins(topo: Topology) -> Tensor:
    """depth_encoder.py:68-74 -- log-spaced bins (float64 arange, exp, stored as default dtype)."""
    lo, hi = np.log(topo.min_depth), np.log(topo.max_depth)
    return torch.exp(torch.arange(lo, hi, (hi - lo) / topo.n_bins)).float()


def param_specs(topo: Topology):
    specs = []
    if topo.distill:    # MonoDepthInference (teacher_model.py:5-13): depth_backbone + depth_decoder; DistillWPoseMeta builds it first
        t = teacher_topology(topo)
        specs += resnet_param_specs("teacher_net.depth_backbone.", t.depth)
        specs.append(("teacher_net.depth_decoder.depth_bins", (t.n_bins,), "teacher_bins"))
        specs += decoder_param_specs("teacher_net.depth_decoder.", t)
    specs += resnet_param_specs("depth_backbone.", topo.depth)
    if topo.posenet:
        specs += resnet_param_specs("pose_backbone.", topo.pose_depth, num_input_images=2)
    specs.append(("head.depth_decoder.depth_bins", (topo.n_bins,), "bins"))
    specs += decoder_param_specs("head.depth_decoder.", topo, uncertain=topo.distill)
    if topo.posenet:
        c_last = 512 * (4 if topo.pose_depth >= 50 else 1)
        specs += pose_decoder_param_specs("head.pose_decoder.", c_last, 1, 2)
    return specs


def teacher_topology(topo: Topology) -> Topology:
    """The teacher of the distillation configs: the stage-1 network (configs/distill_kitti_example:176-197), called
    without P2 (teacher_model.py:15-18), so never fx-scaled."""
    return Topology(depth=topo.teacher_depth, n_bins=topo.teacher_bins, scales=topo.scales, multi_channel=True,
                    min_depth=topo.min_depth, max_depth=topo.max_depth, base_fx=None, use_skips=True,
                    height=topo.height, width=topo.width)


def make_state_dict(topo: Topology, seed: int = 123) -> "OrderedDict[str, Tensor]":
    """Deterministic synthetic weights in the reference's key layout.

    Both the reference (via ``load_state_dict``) and the B200 model are loaded from this, so no
    57 MB checkpoint has to travel.  Scales follow the reference initialisers (fan-out kaiming for
    the ResNet, resnet.py:126-132; fan-in uniform for ``nn.Conv2d`` defaults elsewhere); BN affine
    parameters are perturbed away from (1, 0) so that parity tests see them.
    """
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, Tensor]" = OrderedDict()

    def used_stats(name):       # running statistics that a TRAINING step reads get the values of a trained network, not (0, 1)
        return name.startswith("teacher_net.") or ((topo.norm_eval or topo.frozen_stages >= 0) and name.startswith("depth_backbone."))

    for name, shape, kind in param_specs(topo):
        if kind == "conv":
            cout, cin, kh, kw = shape
            if name.startswith(("depth_backbone", "pose_backbone", "teacher_net.depth_backbone")):
                std = math.sqrt(2.0 / (kh * kw * cout))
            else:
                std = math.sqrt(1.0 / (3.0 * cin * kh * kw)) * math.sqrt(3.0)  # ~ kaiming_uniform(a=sqrt(5)) variance
            sd[name] = torch.randn(shape, generator=g) * std
        elif kind == "bias":
            sd[name] = (torch.rand(shape, generator=g) - 0.5) * 0.1
        elif kind == "bn_w":
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "bn_b":
            sd[name] = 0.1 * torch.randn(shape, generator=g)
        elif kind == "zeros":
            sd[name] = 0.05 * torch.randn(shape, generator=g) if used_stats(name) else torch.zeros(shape)
        elif kind == "ones":
            sd[name] = 1.0 + 0.2 * torch.rand(shape, generator=g) if used_stats(name) else torch.ones(shape)
        elif kind == "count":
            sd[name] = torch.zeros((), dtype=torch.long)
        elif kind == "bins":
            sd[name] = depth_bins(topo)
        elif kind == "teacher_bins":
            sd[name] = depth_bins(teacher_topology(topo))
        else:
            raise ValueError(kind)
    return sd


# --------------------------------------------------------------------------------------------
# networks (functional)
# --------------------------------------------------------------------------------------------
def _bn(sd, prefix: str, x: Tensor, training: bool = True) -> Tensor:
    """nn.BatchNorm2d, train mode => batch statistics (norm_eval=False, configs/kitti_wpose_example:184)."""
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], training=training, momentum=0.1, eps=1e-5)


def resnet_forward(sd, prefix: str, x: Tensor, depth: int, training: bool = True, norm_eval: bool = False,
                   frozen_stages: int = -1) -> List[Tensor]:
    """ResNet.forward, resnet.py:199-213 with out_indices=(-1,0,1,2,3); BasicBlock :34-50, Bottleneck :71-89.
    ``norm_eval`` / ``frozen_stages`` (ResNet.train, resnet.py:169-197): which BatchNorms use their running statistics although
    the model trains (all of them / those of the stem and of layer1..k)."""
    outs = []
    stem_training = training and not norm_eval and frozen_stages < 0
    x = F.conv2d(x, sd[prefix + "conv1.weight"], None, stride=2, padding=3)
    x = F.relu(_bn(sd, prefix + "bn1", x, stem_training))
    outs.append(x)
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    bottleneck = depth >= 50
    model_training = training
    for li, nblocks in enumerate(RESNET_LAYERS[depth]):
        training = model_training and not norm_eval and li + 1 > frozen_stages
        for bi in range(nblocks):
            p = f"{prefix}layer{li + 1}.{bi}."
            stride = 2 if (li > 0 and bi == 0) else 1
            res = x
            if bottleneck:
                o = F.relu(_bn(sd, p + "bn1", F.conv2d(x, sd[p + "conv1.weight"]), training))
                o = F.relu(_bn(sd, p + "bn2", F.conv2d(o, sd[p + "conv2.weight"], stride=stride, padding=1), training))
                o = _bn(sd, p + "bn3", F.conv2d(o, sd[p + "conv3.weight"]), training)
            else:
                o = F.relu(_bn(sd, p + "bn1", F.conv2d(x, sd[p + "conv1.weight"], stride=stride, padding=1), training))
                o = _bn(sd, p + "bn2", F.conv2d(o, sd[p + "conv2.weight"], padding=1), training)
            if (p + "downsample.0.weight") in sd:
                res = _bn(sd, p + "downsample.1", F.conv2d(x, sd[p + "downsample.0.weight"], stride=stride), training)
            x = F.relu(o + res)
        outs.append(x)
    return outs


def _conv3x3(x: Tensor, w: Tensor, b: Optional[Tensor], replicate: bool) -> Tensor:
    if replicate:   # padding_mode='replicate' (depth_encoder.py:59,62)
        return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="replicate"), w, b)
    return F.conv2d(x, w, b, padding=1)


def gather_depth(logits: Tensor, bins: Tensor, topo: Topology, depth_scale) -> Tuple[Tensor, Tensor]:
    """MultiChannelDepthDecoder.gather_output + _gather_activation (depth_encoder.py:76-88,115-121),
    depth_to_disp (monodepth_utils.py:19-24)."""
    p = torch.softmax(torch.clamp(logits, -10.0, 10.0), dim=1)
    depth = (p * bins.reshape(1, -1, 1, 1)).sum(1, keepdim=True)
    if topo.base_fx is not None:
        depth = depth * depth_scale
    mn, mx = topo.min_depth * depth_scale, topo.max_depth * depth_scale
    disp = (1 / depth - 1 / mx) / (1 / mn - 1 / mx)
    return depth, disp


def decoder_forward(sd, prefix: str, feats: Sequence[Tensor], topo: Topology, P2: Optional[Tensor] = None,
                    training: bool = True, uncertain: bool = False) -> Dict:
    """DepthDecoder / MultiChannelDepthDecoder.forward (depth_encoder.py:90-111,123-139)."""
    out: Dict = {}
    if topo.base_fx is None or P2 is None:       # _get_scale, depth_encoder.py:36-43
        depth_scale = 1
    else:
        depth_scale = (P2[:, 0, 0] / topo.base_fx).reshape(-1, 1, 1, 1)
    disp_index = {s: 10 + k for k, s in enumerate(topo.scales)}
    x = feats[-1]
    k = 0
    for i in range(4, -1, -1):
        for j in range(2):
            p = f"{prefix}decoder.{k}.sequence."
            if j == 1:
                x = F.interpolate(x, scale_factor=2, mode="nearest")
                if topo.use_skips and i > 0:
                    x = torch.cat([x, feats[i - 1]], 1)
            x = _conv3x3(x, sd[p + "0.weight"], sd[p + "0.bias"], replicate=(j == 1))
            x = F.relu(_bn(sd, p + "1", x, training))          # ConvBnReLU always applies ReLU, blocks.py:47-54
            k += 1
        if i in topo.scales:
            q = f"{prefix}decoder.{disp_index[i]}."
            logits = _conv3x3(x, sd[q + "weight"], sd[q + "bias"], replicate=True)
            out[("logits", i)] = logits
            if topo.multi_channel:
                out[("depth", i, i)], out[("disp", i)] = gather_depth(logits, sd[prefix + "depth_bins"], topo, depth_scale)
            else:
                disp = torch.sigmoid(logits)
                out[("disp", i)] = disp
                depth = 1 / (1 / topo.max_depth + (1 / topo.min_depth - 1 / topo.max_depth) * disp)  # monodepth_utils.py:8-17
                out[("depth", i, i)] = depth * depth_scale
            if uncertain:   # MultiChannelDepthDecoderUncertain.forward, depth_encoder.py:190
                q = f"{prefix}decoder.{disp_index[i] + len(topo.scales)}."
                out[("uncertain_z", i)] = torch.sigmoid(_conv3x3(x, sd[q + "weight"], sd[q + "bias"], replicate=True))
    return out


def pose_decoder_forward(sd, prefix: str, last_feature: Tensor, n_pred: int = 2) -> Tuple[Tensor, Tensor]:
    """PoseDecoder.forward with num_input_features=1 (pose_decoder.py:26-45)."""
    x = F.relu(F.conv2d(last_feature, sd[prefix + "net.0.weight"], sd[prefix + "net.0.bias"]))
    x = F.relu(F.conv2d(x, sd[prefix + "net.1.weight"], sd[prefix + "net.1.bias"], padding=1))
    x = F.relu(F.conv2d(x, sd[prefix + "net.2.weight"], sd[prefix + "net.2.bias"], padding=1))
    x = F.conv2d(x, sd[prefix + "net.3.weight"], sd[prefix + "net.3.bias"])
    x = 0.01 * x.mean(3).mean(2).view(-1, n_pred, 1, 6)
    return x[..., :3], x[..., 3:]


def transformation_from_parameters(axisangle: Tensor, translation: Tensor, invert: bool) -> Tensor:
    """monodepth_utils.py:46-63 + rot_from_axisangle :298-337 (Rodrigues with the +1e-7 in the axis)."""
    B = axisangle.shape[0]
    angle = torch.norm(axisangle, 2, 2, True)
    axis = axisangle / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x, y, z = axis[..., 0:1], axis[..., 1:2], axis[..., 2:3]
    rows = [x * x * C + ca, x * y * C - z * sa, z * x * C + y * sa,
            x * y * C + z * sa, y * y * C + ca, y * z * C - x * sa,
            z * x * C - y * sa, y * z * C + x * sa, z * z * C + ca]
    R3 = torch.cat(rows, dim=-1).reshape(B, 3, 3)
    R = torch.eye(4, dtype=axisangle.dtype).repeat(B, 1, 1)
    R = torch.cat([torch.cat([R3, torch.zeros(B, 3, 1)], 2), R[:, 3:4, :]], 1)
    t = translation.reshape(B, 3, 1)
    if invert:
        R = R.transpose(1, 2)
        t = -t
    T = torch.eye(4, dtype=axisangle.dtype).repeat(B, 1, 1)
    T = torch.cat([torch.cat([T[:, :3, :3], t], 2), T[:, 3:4, :]], 1)
    return torch.matmul(R, T) if invert else torch.matmul(T, R)


# --------------------------------------------------------------------------------------------
# loss chain
# --------------------------------------------------------------------------------------------
def intrinsics_4x4(P2: Tensor) -> Tuple[Tensor, Tensor]:
    """K embedded in a 4x4 identity and its pseudo-inverse in fp64 on the host, cast to fp32
    (monodepth2_decoder.py:82-90)."""
    B = P2.shape[0]
    K = np.zeros([B, 4, 4])
    K[:, 0:3, 0:3] = P2[:, 0:3, 0:3].detach().cpu().numpy()
    K[:, 3, 3] = 1
    inv_K = np.linalg.pinv(K)
    return torch.from_numpy(K).float(), torch.from_numpy(inv_K).float()


def pixel_grid(h: int, w: int) -> Tensor:
    """BackprojectDepth.get_grid (monodepth_utils.py:105-117): rows (x, y, 1), x = column index."""
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    return torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(h * w)], 0)


def reproject_grid(depth: Tensor, K: Tensor, inv_K: Tensor, T: Tensor) -> Tensor:
    """BackprojectDepth.forward + Project3D.forward (monodepth_utils.py:132-143,154-165) -> [B,H,W,2]."""
    B, _, H, W = depth.shape
    cam = torch.matmul(inv_K[:, :3, :3], pixel_grid(H, W).unsqueeze(0).expand(B, -1, -1))
    cam = depth.view(B, 1, -1) * cam
    cam = torch.cat([cam, torch.ones(B, 1, H * W)], 1)
    P = torch.matmul(K, T)[:, :3, :]
    p = torch.matmul(P, cam)
    pix = p[:, :2, :] / (p[:, 2, :].unsqueeze(1) + 1e-7)
    pix = pix.view(B, 2, H, W).permute(0, 2, 3, 1)
    pix = torch.stack([pix[..., 0] / (W - 1), pix[..., 1] / (H - 1)], -1)
    return (pix - 0.5) * 2


def ssim(x: Tensor, y: Tensor) -> Tensor:
    """SSIM.forward (monodepth_utils.py:201-215): reflect pad 1, 3x3 mean, biased variances."""
    x = F.pad(x, (1, 1, 1, 1), mode="reflect")
    y = F.pad(y, (1, 1, 1, 1), mode="reflect")
    mu_x, mu_y = F.avg_pool2d(x, 3, 1), F.avg_pool2d(y, 3, 1)
    sigma_x = F.avg_pool2d(x * x, 3, 1) - mu_x ** 2
    sigma_y = F.avg_pool2d(y * y, 3, 1) - mu_y ** 2
    sigma_xy = F.avg_pool2d(x * y, 3, 1) - mu_x * mu_y
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    n = (2 * mu_x * mu_y + C1) * (2 * sigma_xy + C2)
    d = (mu_x ** 2 + mu_y ** 2 + C1) * (sigma_x + sigma_y + C2)
    return torch.clamp((1 - n / d) / 2, 0, 1)


def photometric(pred: Tensor, target: Tensor) -> Tensor:
    """compute_reprojection_loss (monodepth2_decoder.py:118-128)."""
    l1 = (target - pred).abs().mean(1, True)
    return 0.85 * ssim(pred, target).mean(1, True) + 0.15 * l1


def smooth_loss(disp: Tensor, img: Tensor) -> Tensor:
    """get_smooth_loss (monodepth_utils.py:168-181)."""
    gx = (disp[:, :, :, :-1] - disp[:, :, :, 1:]).abs()
    gy = (disp[:, :, :-1, :] - disp[:, :, 1:, :]).abs()
    ix = (img[:, :, :, :-1] - img[:, :, :, 1:]).abs().mean(1, True)
    iy = (img[:, :, :-1, :] - img[:, :, 1:, :]).abs().mean(1, True)
    return (gx * torch.exp(-ix)).mean() + (gy * torch.exp(-iy)).mean()


def warp_sources(depth_full: Tensor, data: Dict, cam_T: Dict[int, Tensor], topo: Topology):
    """Inner loop of _generate_images_pred (monodepth2_decoder.py:75-116) for one scale."""
    K, inv_K = intrinsics_4x4(data["P2"])
    B, _, H, W = depth_full.shape
    warped, overlap = {}, {}
    for f in topo.frame_ids[1:]:
        grid = reproject_grid(depth_full, K, inv_K, cam_T[f])
        warped[f] = F.grid_sample(data[("original_image", f)], grid, padding_mode="border", align_corners=True)
        if topo.overlapped_mask:
            pm = data["patched_mask"] if "patched_mask" in data else torch.ones(B, H, W)
            rp = F.grid_sample(pm.unsqueeze(1).float(), grid, align_corners=True, mode="nearest")
            overlap[f] = (rp == 1).squeeze(1)
    return warped, overlap


# --------------------------------------------------------------------------------------------
# MEI (unified omnidirectional) camera -- FishEyeDecoder (monodepth2_decoder.py:350-420)
# --------------------------------------------------------------------------------------------
def _mei_radial(k1: float, k2: float, r1: np.ndarray, r0: np.ndarray) -> np.ndarray:
    """radial_distort_func (mei_fisheye_utils.py:66-68), fp64."""
    r2 = r0 * r0
    return r0 - r1 / (1 + k1 * r2 + k2 * (r2 * r2))


def _mei_mirror(r0: np.ndarray, xi: float, Z) -> np.ndarray:
    """mirror_backtrack_func (mei_fisheye_utils.py:81-83), fp64."""
    return r0 * r0 - (1 - Z * Z) / ((xi + Z) * (xi + Z))


def mei_lut(H: int, W: int, gamma1: float, gamma2: float, u0: float, v0: float, k1: float, k2: float, xi: float,
            tol: float = 1e-6, max_iter: int = 100):
    """The per-calibration look-up table of MeiCameraProjection.image2cam (mei_fisheye_utils.py:139-170):
    returns fp32 ``X, Y, Z, mask`` of shape [H, W].  The per-pixel Newton solve (:70-79, finite-difference
    derivative with step ``tol``) and bisection (:85-101) are restated vectorised with per-element
    early exit, in fp64 like numba evaluates them (fp32 ``r1`` promoted by the fp64 calibration scalars)."""
    xs, ys = np.meshgrid(np.arange(W), np.arange(H), indexing="xy")
    X = ((xs.astype(np.float32) - u0) / gamma1).astype(np.float32)
    Y = ((ys.astype(np.float32) - v0) / gamma2).astype(np.float32)
    r1 = np.sqrt(X ** 2 + Y ** 2).astype(np.float32).astype(np.float64)
    # Newton (:70-79)
    x = r1.copy()
    done = np.zeros(r1.shape, bool)
    with np.errstate(all="ignore"):
        for _ in range(max_iter):
            f = _mei_radial(k1, k2, r1, x)
            done |= np.abs(f) < tol
            if done.all():
                break
            df = (_mei_radial(k1, k2, r1, x + tol) - f) / tol
            x = np.where(done, x, x - f / df)
        r0 = x
        # bisection on Z in [0, 1] (:85-101)
        y0 = _mei_mirror(r0, xi, 0.0)
        y1 = _mei_mirror(r0, xi, 1.0)
        flag = ~(y0 * y1 > 0)
        lo = np.zeros(r1.shape)
        hi = np.ones(r1.shape)
        z = np.where(flag, 0.0, -1.0)
        done = ~flag
        for _ in range(max_iter):
            mid = (lo + hi) / 2
            f = _mei_mirror(r0, xi, mid)
            z = np.where(done, z, mid)
            done |= np.abs(f) < tol
            if done.all():
                break
            left = f * _mei_mirror(r0, xi, lo) < 0
            hi = np.where(~done & left, mid, hi)
            lo = np.where(~done & ~left, mid, lo)
    Z = z.astype(np.float32)
    mask = flag.astype(np.float32)
    mask[Z < 0.05] = 0                                   # :161
    dead = mask == 0
    Z[dead] = -1
    X[dead] = -1
    Y[dead] = -1
    X = (X * (Z + np.float32(xi))).astype(np.float32)    # :167-168 (no r0/r1 un-distortion factor: follow the code)
    Y = (Y * (Z + np.float32(xi))).astype(np.float32)
    return X, Y, Z, mask


def mei_lut_batch(P2: Tensor, calib: Sequence[dict], H: int, W: int) -> Tensor:
    """[B, 4, H, W] (X, Y, Z, mask) -- one table per sample, cached per distinct calibration."""
    cache, out = {}, []
    for b in range(P2.shape[0]):
        key = (P2[b, 0, 0].item(), P2[b, 1, 1].item(), P2[b, 0, 2].item(), P2[b, 1, 2].item(),
               calib[b]["distortion_parameters"]["k1"], calib[b]["distortion_parameters"]["k2"], calib[b]["mirror_parameters"]["xi"])
        if key not in cache:
            cache[key] = torch.from_numpy(np.stack(mei_lut(H, W, *key), 0))
        out.append(cache[key])
    return torch.stack(out, 0)


def mei_cam2image(points: Tensor, P: Tensor, calib: dict) -> Tuple[Tensor, Tensor]:
    """_cam2image (mei_fisheye_utils.py:23-51) for points [..., 3] of one sample -> pixel (u, v)."""
    eps = 1e-6
    norm = torch.norm(points, dim=-1)
    x = points[..., 0] / (norm + eps)
    y = points[..., 1] / (norm + eps)
    z = points[..., 2] / (norm + eps)
    xi = calib["mirror_parameters"]["xi"]
    x = x / (z + xi + eps)
    y = y / (z + xi + eps)
    k1, k2 = calib["distortion_parameters"]["k1"], calib["distortion_parameters"]["k2"]
    ro2 = x * x + y * y
    x = x * (1 + k1 * ro2 + k2 * ro2 * ro2)
    y = y * (1 + k1 * ro2 + k2 * ro2 * ro2)
    return P[0, 0] * x + P[0, 2], P[1, 1] * y + P[1, 2]


def warp_sources_fisheye(norm_full: Tensor, data: Dict, cam_T: Dict[int, Tensor], topo: Topology):
    """FishEyeDecoder._generate_images_pred inner loop (monodepth2_decoder.py:367-413) for one scale."""
    B, _, H, W = norm_full.shape
    lut = mei_lut_batch(data["P2"], data["calib_meta"], H, W)
    points = torch.stack([lut[:, 0:1] * norm_full, lut[:, 1:2] * norm_full, lut[:, 2:3] * norm_full], -1)   # [B,1,H,W,3]
    homo = torch.cat([points, torch.ones_like(points[..., :1])], -1).squeeze(1)[..., None]
    warped, overlap = {}, {}
    for f in topo.frame_ids[1:]:
        tp = torch.matmul(cam_T[f][:, None, None], homo)[..., 0]
        grids = []
        for b in range(B):
            u, v = mei_cam2image(tp[b, ..., 0:3], data["P2"][b], data["calib_meta"][b])
            grids.append(torch.stack([u / max(W - 1, 1) * 2 - 1, v / max(H - 1, 1) * 2 - 1], -1))
        grid = torch.stack(grids, 0)
        warped[f] = F.grid_sample(data[("original_image", f)], grid, padding_mode="border", align_corners=True)
        if topo.overlapped_mask:
            pm = (data["patched_mask"] if "patched_mask" in data else torch.ones(B, H, W)) * lut[:, 3]
            rp = F.grid_sample(pm.unsqueeze(1).float(), grid, align_corners=True, mode="nearest")
            overlap[f] = (rp == 1).squeeze(1)
    return warped, overlap


def fisheye_prediction(norm: Tensor, data: Dict) -> Dict:
    """FishEyeDecoder.get_prediction (monodepth2_decoder.py:415-420): z of the back-projected ray + the norm."""
    H, W = norm.shape[-2:]
    lut = mei_lut_batch(data["P2"], data["calib_meta"], H, W)
    return {"depth": lut[:, 2:3] * norm, "norm": norm}


def loss_chain(outputs: Dict, data: Dict, cam_T: Dict[int, Tensor], topo: Topology,
               noise: Optional[Dict[int, Tensor]] = None, keep: bool = False) -> Dict:
    """compute_total_reprojection_loss + loss (monodepth2_decoder.py:205-347), pinhole, default weights.

    ``noise[s]`` is the ``randn`` draw of line :258 for scale ``s`` (shape [B,2,H,W]); None => zeros.
    """
    target = data[("original_image", 0)]
    B, _, H, W = target.shape
    losses: Dict[str, Tensor] = {}
    aux: Dict = {}
    total = 0
    for s in topo.scales:
        depth_full = F.interpolate(outputs[("depth", s, s)], [H, W], mode="bilinear", align_corners=True)
        warped, overlap = (warp_sources_fisheye if topo.fisheye else warp_sources)(depth_full, data, cam_T, topo)
        disp = outputs[("disp", s)]
        color = target if s == 0 else F.adaptive_avg_pool2d(target, disp.shape[-2:])
        reproj = []
        for f in topo.frame_ids[1:]:
            pl = photometric(warped[f], target)
            if topo.overlapped_mask:
                pl = torch.where(overlap[f].unsqueeze(1), pl, torch.full_like(pl, 100.0))
            reproj.append(pl)
        reproj = torch.cat(reproj, 1)
        if "motion_mask" in data:
            mm = data["motion_mask"]
            to_opt, idxs = torch.min(reproj, dim=1)
            to_opt = to_opt.detach() * mm + to_opt * (1 - mm)
        else:
            ident = torch.cat([photometric(data[("original_image", f)], target) for f in topo.frame_ids[1:]], 1)
            if noise is not None:
                ident = ident + noise[s] * 0.00001
            to_opt, idxs = torch.min(torch.cat((ident, reproj), dim=1), dim=1)
        pm = data["patched_mask"] if "patched_mask" in data else torch.ones(B, H, W)
        to_opt = to_opt * pm
        loss = to_opt.sum() / (pm.sum() + 1e-6)
        norm_disp = disp / (disp.mean(2, True).mean(3, True) + 1e-7)
        sm = smooth_loss(norm_disp, color) * 1e-5 / (2 ** s)
        losses[f"smooth_loss/{s}"] = sm.detach()
        loss = loss + sm
        total = total + loss
        losses[f"loss/{s}"] = loss.detach()
        if keep:
            aux[("warped", s)] = warped
            aux[("overlap", s)] = overlap
            aux[("idxs", s)] = idxs
            aux[("depth", 0, s)] = depth_full
    total = total / len(topo.scales)
    losses["total_loss"] = total.detach()
    return {"loss": total, "loss_dict": losses, "aux": aux}


# --------------------------------------------------------------------------------------------
# whole step
# --------------------------------------------------------------------------------------------
def forward_train(sd, data: Dict, topo: Topology, noise: Optional[Dict[int, Tensor]] = None, keep: bool = False) -> Dict:
    """MonoDepthWPose.forward_train / MonoDepthMeta.forward_train (monodepth2_model.py:24-46,85-130)."""
    feats = resnet_forward(sd, "depth_backbone.", data[("image", 0)], topo.depth, norm_eval=topo.norm_eval,
                           frozen_stages=topo.frozen_stages)
    P2 = None if topo.posenet else data["P2"]          # MonoDepthMeta calls forward_depth(features) without P2 (:27)
    outputs = decoder_forward(sd, "head.depth_decoder.", feats, topo, P2, uncertain=topo.distill)
    if topo.distill:
        outputs.update(teacher_depths(sd, data[("image", 0)], topo))
    cam_T: Dict[int, Tensor] = {}
    for f in topo.frame_ids[1:]:
        if topo.posenet:
            pair = [data[("image", f)], data[("image", 0)]] if f < 0 else [data[("image", 0)], data[("image", f)]]
            pf = resnet_forward(sd, "pose_backbone.", torch.cat(pair, 1), topo.pose_depth)
            aa, tr = pose_decoder_forward(sd, "head.pose_decoder.", pf[-1], 2)
            cam_T[f] = transformation_from_parameters(aa[:, 0], tr[:, 0], invert=(f < 0))
        else:
            cam_T[f] = data[("relative_pose", f)]
    out = loss_chain(outputs, data, cam_T, topo, noise, keep)
    if topo.distill:        # MonoDepth2Decoder.loss, monodepth2_decoder.py:328-334: added AFTER the division by num_scales
        total = out["loss"]
        for s in topo.scales:
            d = distill_loss(outputs, s, topo)
            out["loss_dict"][f"distilation/{s}"] = d.detach()
            total = total + d * topo.distill_weight
        out["loss"] = total
        out["loss_dict"]["total_loss"] = total.detach()
    out["outputs"] = outputs
    out["cam_T"] = cam_T
    return out


def teacher_depths(sd, image: Tensor, topo: Topology) -> Dict:
    """MonoDepthInference.compute_teacher_depth (teacher_model.py:20-32) of the frozen teacher; DistillWPoseMeta.train keeps
    it in eval mode (monodepth2_model.py:172-174): running statistics, no gradient."""
    t = teacher_topology(topo)
    with torch.no_grad():
        feats = resnet_forward(sd, "teacher_net.depth_backbone.", image, t.depth, training=False)
        out = decoder_forward(sd, "teacher_net.depth_decoder.", feats, t, None, training=False)
    return {("teacher_depth", k[1], k[2]): v for k, v in out.items() if k[0] == "depth"}


def distill_loss(outputs: Dict, s: int, topo: Topology) -> Tensor:
    """compute_distill_loss (monodepth2_decoder.py:185-203), the scaled branch every shipped config uses."""
    error = (outputs[("teacher_depth", s, s)].detach() - outputs[("depth", s, s)]).abs()
    if topo.uncertain_distill:
        u = outputs[("uncertain_z", s)]
        return (error / u + torch.log(u + 1e-5)).mean()
    return error.mean()


def forward_test(sd, data: Dict, topo: Topology) -> Dict:
    """forward_test (monodepth2_model.py:132-136); BN uses running statistics in eval mode."""
    feats = resnet_forward(sd, "depth_backbone.", data[("image", 0)], topo.depth, training=False)
    outputs = decoder_forward(sd, "head.depth_decoder.", feats, topo, None if topo.posenet else data["P2"], training=False,
                              uncertain=topo.distill)
    if topo.fisheye:
        return fisheye_prediction(outputs[("depth", 0, 0)], data)
    return {"depth": outputs[("depth", 0, 0)]}


def trainable(sd, topo: Optional[Topology] = None) -> List[str]:
    frozen: Tuple[str, ...] = ()
    if topo is not None and topo.frozen_stages >= 0:       # depth encoder only (the PoseNet is built with its own arguments)
        frozen = ("depth_backbone.conv1.", "depth_backbone.bn1.") + tuple(f"depth_backbone.layer{i}." for i in range(1, topo.frozen_stages + 1))
    return [k for k, v in sd.items() if v.is_floating_point() and not k.endswith(("running_mean", "running_var", "depth_bins"))
            and not k.startswith("teacher_net.") and not k.startswith(frozen or ("\0",))]


class OracleTrainer:
    """BaseTrainingHook.__call__ on the CPU (base_training_hooks.py:30-49): zero_grad, forward,
    ``loss.mean().backward()``, ``clip_grad_norm_``, Adam step.  Used as the CPU baseline."""

    def __init__(self, topo: Topology, seed: int = 123, lr: float = 1e-4, clip: Optional[float] = 35.0):
        self.topo = topo
        self.sd = make_state_dict(topo, seed)
        self.params = [self.sd[k].requires_grad_(True) for k in trainable(self.sd, topo)]
        self.opt = torch.optim.Adam(self.params, lr=lr, weight_decay=0)
        self.clip = clip

    def step(self, data: Dict, noise=None) -> Dict:
        self.opt.zero_grad()
        out = forward_train(self.sd, data, self.topo, noise)
        out["loss"].mean().backward()
        if self.clip is not None:
            torch.nn.utils.clip_grad_norm_(self.params, self.clip)
        self.opt.step()
        return out


# --------------------------------------------------------------------------------------------
# synthetic inputs -- shared by golden generation, tests and bench
# --------------------------------------------------------------------------------------------
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def synthetic_batch(B: int, H: int, W: int, seed: int = 1234, frame_ids=(0, 1, -1), mask_dtype=torch.float64,
                    fx_scale: float = 0.58, fy_scale: float = 1.92) -> Dict:
    """Seeded smooth-plus-noise triplets, KITTI-like P2, forward-motion poses, fp64 patched mask with
    zero border strips on half of the samples."""
    g = torch.Generator().manual_seed(seed)
    data: Dict = {}
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    base = torch.rand(B, 3, max(H // 8, 2), max(W // 8, 2), generator=g)
    for f in frame_ids:
        # the three frames share a base field (so that reprojection can win the min) plus a per-frame shift
        lo = base + 0.15 * torch.rand(base.shape, generator=g)
        img = F.interpolate(lo, size=(H, W), mode="bilinear", align_corners=False)
        img = (img / 1.15 + 0.05 * torch.rand(B, 3, H, W, generator=g)).clamp(0, 1)
        data[("original_image", f)] = img.contiguous()
        data[("image", f)] = ((img - mean) / std).contiguous()
    P2 = torch.zeros(B, 3, 4)
    P2[:, 0, 0] = fx_scale * W
    P2[:, 0, 2] = 0.5 * W
    P2[:, 1, 1] = fy_scale * H
    P2[:, 1, 2] = 0.5 * H
    P2[:, 2, 2] = 1.0
    data["P2"] = P2
    for f in frame_ids[1:]:
        ang = (torch.rand(B, generator=g) * 2 - 1) * math.radians(1.0)
        T = torch.eye(4).repeat(B, 1, 1)
        T[:, 0, 0] = torch.cos(ang)
        T[:, 0, 2] = torch.sin(ang)
        T[:, 2, 0] = -torch.sin(ang)
        T[:, 2, 2] = torch.cos(ang)
        tz = (0.8 + 0.2 * (torch.rand(B, generator=g) * 2 - 1)) * (-1.0 if f > 0 else 1.0)
        T[:, 2, 3] = tz
        T[:, 0, 3] = 0.05 * (torch.rand(B, generator=g) * 2 - 1)
        data[("relative_pose", f)] = T
    mask = torch.ones(B, H, W, dtype=mask_dtype)
    for b in range(0, B, 2):
        wstrip = int(torch.randint(1, max(W // 10, 2), (1,), generator=g))
        if (b // 2) % 2 == 0:
            mask[b, :, :wstrip] = 0
        else:
            mask[b, :, W - wstrip:] = 0
    data["patched_mask"] = mask
    return data


MEI_KITTI360 = dict(xi=2.2134, k1=0.016798, k2=1.6548, gamma=1336.3, u0=716.94, v0=705.76, size=1400.0)


def synthetic_fisheye_batch(B: int, H: int, W: int, seed: int = 1234, frame_ids=(0, 1, -1), mask_dtype=torch.float64,
                            two_calibrations: bool = False) -> Dict:
    """``synthetic_batch`` with a KITTI-360-like MEI calibration scaled to the crop, lateral
    motion, ``calib_meta`` as the dataset delivers it (fisheye_dataset.py:45-58,254)."""
    data = synthetic_batch(B, H, W, seed, frame_ids, mask_dtype)
    g = torch.Generator().manual_seed(seed + 77)
    m = MEI_KITTI360
    P2 = torch.zeros(B, 3, 4)
    calib = []
    for b in range(B):
        k = 1.0 if not (two_calibrations and b % 2) else 0.97
        P2[b, 0, 0] = m["gamma"] * W / m["size"] * k
        P2[b, 1, 1] = m["gamma"] * H / m["size"] * k
        P2[b, 0, 2] = m["u0"] * W / m["size"]
        P2[b, 1, 2] = m["v0"] * H / m["size"]
        P2[b, 2, 2] = 1.0
        calib.append(dict(mirror_parameters=dict(xi=m["xi"]), distortion_parameters=dict(k1=m["k1"], k2=m["k2"] * k)))
    data["P2"] = P2
    data["calib_meta"] = calib
    for f in frame_ids[1:]:
        T = data[("relative_pose", f)]
        tx = (0.8 + 0.2 * (torch.rand(B, generator=g) * 2 - 1)) * (-1.0 if f > 0 else 1.0)
        T[:, 0, 3] = tx                      # side-looking fisheye: the vehicle's forward motion is lateral
        T[:, 2, 3] = 0.1 * (torch.rand(B, generator=g) * 2 - 1)
    return data


def tie_break_noise(B: int, H: int, W: int, scales: Sequence[int], seed: int = 0) -> Dict[int, Tensor]:
    """The reference draws ``torch.randn([B,2,H,W])`` once per scale, in scale order, from the CPU
    default generator (monodepth2_decoder.py:258-259).  ``torch.manual_seed(seed)`` right before
    ``head.loss`` makes its draws equal to these."""
    g = torch.Generator().manual_seed(seed)
    return {s: torch.randn(B, 2, H, W, generator=g) for s in scales}


def synthetic_depth_outputs(B: int, H: int, W: int, scales: Sequence[int], seed: int, lo: float = 2.0, hi: float = 40.0,
                            min_depth: float = 0.5, max_depth: float = 100.0) -> Dict:
    """Smooth random depth / disparity pyramids standing in for the decoder's output (loss-only cases)."""
    g = torch.Generator().manual_seed(seed)
    out: Dict = {}
    for s in scales:
        h, w = H // (2 ** s), W // (2 ** s)
        field = torch.rand(B, 1, max(h // 8, 2), max(w // 8, 2), generator=g)
        field = F.interpolate(field, size=(h, w), mode="bilinear", align_corners=True)
        depth = lo * torch.exp(field * math.log(hi / lo)) * (1 + 0.02 * torch.rand(B, 1, h, w, generator=g))
        out[("depth", s, s)] = depth.contiguous()
        out[("disp", s)] = ((1 / depth - 1 / max_depth) / (1 / min_depth - 1 / max_depth)
                            * (1 + 0.05 * torch.rand(B, 1, h, w, generator=g))).contiguous()
    return out


def synthetic_motion_mask(B: int, H: int, W: int, seed: int) -> Tensor:
    g = torch.Generator().manual_seed(seed)
    m = F.interpolate(torch.rand(B, 1, H // 8, W // 8, generator=g), size=(H, W), mode="nearest")[:, 0]
    return (m > 0.7).float()
