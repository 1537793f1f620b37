"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE ITSELF.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports the reference read-only with the two shims SURVEY.md section 8(c) lists (an ``easydict``
stand-in and ``Tensor.cuda -> identity``), loads it with the deterministic synthetic weights of
``oracle.fsnet_oracle.make_state_dict`` and feeds it ``oracle.fsnet_oracle.synthetic_batch`` inputs.
Only OUTPUTS are stored (inputs and weights are regenerated from their seeds by the tests; an input
checksum is stored to catch RNG drift).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
# the reference's ``vision_base`` / ``monodepth`` must win over this repo's packages of the same name
sys.path = [REF] + [p for p in sys.path if os.path.abspath(p or ".") not in (REPO, HERE)] + [REPO]

import numpy as np  # noqa: E402
import torch  # noqa: E402

torch.Tensor.cuda = lambda self, *a, **k: self          # the reference hard-codes .cuda()
torch.set_num_threads(8)

from easydict import EasyDict as edict  # noqa: E402
from vision_base.utils.builder import build  # noqa: E402  (reference)
import vision_base  # noqa: E402

assert vision_base.__path__[0].startswith(REF), vision_base.__path__
from oracle import fsnet_oracle as O  # noqa: E402


def ref_meta_arch(topo: O.Topology):
    head = edict(
        name="monodepth.networks.models.heads.monodepth2_decoder." + ("FishEyeDecoder" if topo.fisheye else "MonoDepth2Decoder"),
        scales=list(topo.scales), height=topo.height, width=topo.width,
        min_depth=topo.min_depth, max_depth=topo.max_depth, overlapped_mask=topo.overlapped_mask, is_log_image=False,
        depth_decoder_cfg=edict(
            name="monodepth.networks.models.heads.depth_encoder." + ("MultiChannelDepthDecoder" if topo.multi_channel else "DepthDecoder"),
            num_ch_enc=np.array(topo.num_ch_enc), num_output_channels=topo.n_bins, use_skips=topo.use_skips,
            scales=list(topo.scales), min_depth=topo.min_depth, max_depth=topo.max_depth, base_fx=topo.base_fx))
    backbone = edict(name="vision_base.networks.models.backbone.resnet.resnet", depth=topo.depth, pretrained=False,
                     frozen_stages=topo.frozen_stages, num_stages=4, out_indices=(-1, 0, 1, 2, 3), norm_eval=topo.norm_eval,
                     dilations=(1, 1, 1, 1))
    cfg = edict(depth_backbone_cfg=backbone, head_cfg=head, train_cfg=edict(frame_ids=list(topo.frame_ids)), test_cfg=edict())
    sd = O.make_state_dict(topo)
    if topo.distill:
        import tempfile
        t = O.teacher_topology(topo)
        head.distillation_loss_weight = topo.distill_weight
        head.is_uncertain_distill = topo.uncertain_distill
        head.depth_decoder_cfg.name = "monodepth.networks.models.heads.depth_encoder.MultiChannelDepthDecoderUncertain"
        cfg.name = "monodepth.networks.models.meta_archs.monodepth2_model.DistillWPoseMeta"
        cfg.teacher_net_cfg = edict(
            name="monodepth.networks.models.meta_archs.teacher_model.MonoDepthInference", backbone_cfg=edict(backbone, depth=t.depth),
            depth_head_cfg=edict(name="monodepth.networks.models.heads.depth_encoder.MultiChannelDepthDecoder",
                                 num_ch_enc=np.array(t.num_ch_enc), num_output_channels=t.n_bins, use_skips=True, scales=list(t.scales),
                                 min_depth=t.min_depth, max_depth=t.max_depth))
        # the constructor wants a teacher checkpoint on disk (monodepth2_model.py:160-163): the teacher part of the synthetic weights
        with tempfile.NamedTemporaryFile(suffix=".pth", delete=False) as f:
            torch.save({k[len("teacher_net."):]: v for k, v in sd.items() if k.startswith("teacher_net.")}, f.name)
            cfg.teacher_net_path = f.name
    elif topo.posenet:
        cfg.name = "monodepth.networks.models.meta_archs.monodepth2_model.MonoDepthMeta"
        cfg.pose_backbone_cfg = edict(backbone, depth=topo.pose_depth, num_input_images=2)
        head.pose_decoder_cfg = edict(name="monodepth.networks.models.heads.pose_decoder.PoseDecoder",
                                      num_ch_enc=np.array([64, 64, 128, 256, 512]), num_input_features=1, num_frames_to_predict_for=2)
    else:
        cfg.name = "monodepth.networks.models.meta_archs.monodepth2_model.MonoDepthWPose"
    model = build(**cfg)
    if topo.distill:
        os.unlink(cfg.teacher_net_path)
    missing = model.load_state_dict(sd, strict=True)      # key layout must be identical
    model.train()
    return model, sd


def input_checksum(data):
    return np.array([float(v.double().sum()) for k, v in sorted(data.items(), key=lambda kv: str(kv[0])) if torch.is_tensor(v)])


def run_full(name, topo: O.Topology, B, seed=1234, noise_seed=0, store_disp=True, grads_of=()):
    model, sd = ref_meta_arch(topo)
    data = (O.synthetic_fisheye_batch if topo.fisheye else O.synthetic_batch)(B, topo.height, topo.width, seed, topo.frame_ids)
    out = {"input_checksum": input_checksum(data)}
    # full call through the reference's public entry, seeded so that its randn tie-break draws are known
    torch.manual_seed(noise_seed)
    ret = model(dict(data), dict(is_training=True, epoch_num=0, global_step=0))
    for k, v in ret["loss_dict"].items():
        out["loss_dict/" + k] = np.asarray(v.detach().double())
    out["loss"] = np.asarray(ret["loss"].detach().double())
    out["loss_is_fp64"] = np.asarray(ret["loss"].dtype == torch.float64)
    model.zero_grad()
    ret["loss"].mean().backward()
    named = dict(model.named_parameters())
    gn = {k: float(p.grad.double().norm()) for k, p in named.items() if p.grad is not None}
    out["grad_names"] = np.array(sorted(gn))
    out["grad_norms"] = np.array([gn[k] for k in sorted(gn)])
    for k in grads_of:
        out["grad/" + k] = named[k].grad.detach().numpy().copy()
    # second pass, piecewise, to expose the intermediate maps (same orchestration as forward_train)
    model2, _ = ref_meta_arch(topo)
    feats = model2.depth_backbone(data[("image", 0)])
    outputs = model2.head.forward_depth(feats) if topo.posenet else model2.head.forward_depth(feats, data["P2"])
    for i, f in enumerate(feats):
        out[f"feat_absmean/{i}"] = np.asarray(f.detach().abs().mean())
    if store_disp:
        for s in topo.scales:
            out[f"disp/{s}"] = outputs[("disp", s)].detach().numpy().copy()
            out[f"depth/{s}"] = outputs[("depth", s, s)].detach().numpy().copy()
            if topo.distill:
                out[f"uncertain_z/{s}"] = outputs[("uncertain_z", s)].detach().numpy().copy()
    if topo.distill:
        for k, v in model2.teacher_net.compute_teacher_depth(data[("image", 0)]).items():
            out[f"teacher_depth/{k[1]}"] = v.detach().numpy().copy()
    if topo.posenet:
        from monodepth.networks.utils.monodepth_utils import transformation_from_parameters
        for f_i in topo.frame_ids[1:]:
            pair = [data[("image", f_i)], data[("image", 0)]] if f_i < 0 else [data[("image", 0)], data[("image", f_i)]]
            aa, tr = model2.head.forward_pose([model2.pose_backbone(torch.cat(pair, 1))])
            out[f"axisangle/{f_i}"] = aa.detach().numpy().copy()
            out[f"translation/{f_i}"] = tr.detach().numpy().copy()
            out[f"cam_T_cam/{f_i}"] = transformation_from_parameters(aa[:, 0], tr[:, 0], invert=(f_i < 0)).detach().numpy().copy()
    # eval-mode prediction (forward_test)
    model2.eval()
    with torch.no_grad():
        pred = model2(dict(data), dict(is_training=False))
    out["test_depth"] = pred["depth"].numpy().copy()
    if "norm" in pred:
        out["test_norm"] = pred["norm"].numpy().copy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "loss", float(out["loss"]), "fp64" if out["loss_is_fp64"] else "fp32", os.path.getsize(path) // 1024, "KiB")


def run_loss_only(name, topo: O.Topology, B, seed, noise_seed=0, mask_dtype=torch.float64, with_mask=True,
                  motion_mask=False, depth_lo=2.0, depth_hi=40.0):
    """The reference's MonoDepth2Decoder.loss on given depth / disparity maps (no network): pins the
    fused loss kernels, including d loss / d depth and d loss / d disp."""
    from monodepth.networks.models.heads.monodepth2_decoder import MonoDepth2Decoder, FishEyeDecoder
    head = (FishEyeDecoder if topo.fisheye else MonoDepth2Decoder)(
        scales=list(topo.scales), height=topo.height, width=topo.width, frame_ids=list(topo.frame_ids),
        depth_decoder_cfg=edict(name="monodepth.networks.models.heads.depth_encoder.MultiChannelDepthDecoder",
                                num_ch_enc=np.array([64, 64, 128, 256, 512]), num_output_channels=4, scales=list(topo.scales)),
        overlapped_mask=topo.overlapped_mask, is_log_image=True)
    if topo.fisheye:
        data = O.synthetic_fisheye_batch(B, topo.height, topo.width, seed, topo.frame_ids, mask_dtype=mask_dtype, two_calibrations=True)
    else:
        data = O.synthetic_batch(B, topo.height, topo.width, seed, topo.frame_ids, mask_dtype=mask_dtype)
    if not with_mask:
        del data["patched_mask"]
    outputs = O.synthetic_depth_outputs(B, topo.height, topo.width, topo.scales, seed + 1, depth_lo, depth_hi, topo.min_depth, topo.max_depth)
    if motion_mask:
        data["motion_mask"] = O.synthetic_motion_mask(B, topo.height, topo.width, seed + 2)
    leaves = {k: v.requires_grad_(True) for k, v in outputs.items()}   # ('depth',0,0) is overwritten by the head (:73)
    for f in topo.frame_ids[1:]:
        outputs[("cam_T_cam", f)] = data[("relative_pose", f)].clone().requires_grad_(True)
    torch.manual_seed(noise_seed)
    ret = head.loss(outputs, data)
    ret["loss"].mean().backward()
    out = {"input_checksum": input_checksum(data), "loss": np.asarray(ret["loss"].detach().double())}
    for k, v in ret["loss_dict"].items():
        out["loss_dict/" + k] = np.asarray(v.detach().double())
    for s in topo.scales:
        out[f"grad_depth/{s}"] = leaves[("depth", s, s)].grad.float().numpy().copy()
        out[f"grad_disp/{s}"] = leaves[("disp", s)].grad.float().numpy().copy()
    for f in topo.frame_ids[1:]:
        out[f"grad_T/{f}"] = outputs[("cam_T_cam", f)].grad.float().numpy().copy()
    if "loss_mask_0" in ret["hm"]:
        out["loss_mask_0"] = np.packbits(ret["hm"]["loss_mask_0"]["data"].numpy().astype(np.uint8))
        out["predicted_image_1_absmean"] = np.asarray(ret["hm"]["predicted_image_1"].detach().abs().mean())
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "loss", float(out["loss"]), os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    only = sys.argv[1:]

    def want(n):
        return not only or n in only

    # loss chain alone
    if want("loss_a"):
        run_loss_only("loss_a", O.Topology(height=96, width=160, overlapped_mask=True), B=3, seed=11)
    if want("loss_b"):
        run_loss_only("loss_b", O.Topology(height=64, width=96, overlapped_mask=False, scales=(0, 2)), B=2, seed=12, with_mask=False)
    if want("loss_c"):
        run_loss_only("loss_c", O.Topology(height=64, width=128, overlapped_mask=True), B=2, seed=13, mask_dtype=torch.float32,
                      depth_lo=0.6, depth_hi=6.0)     # near depths: large flow, many out-of-view pixels
    if want("loss_mm"):
        run_loss_only("loss_mm", O.Topology(height=64, width=96, overlapped_mask=True, scales=(0, 1)), B=2, seed=14, motion_mask=True)
    if want("loss_fe"):      # MEI fisheye camera, two distinct calibrations in the batch, overlap mask x LUT mask
        run_loss_only("loss_fe", O.Topology(height=96, width=128, overlapped_mask=True, fisheye=True, max_depth=150.0), B=3, seed=15)
    if want("loss_fe_nomask"):
        run_loss_only("loss_fe_nomask", O.Topology(height=64, width=64, overlapped_mask=True, fisheye=True, scales=(0, 1)), B=2, seed=16,
                      with_mask=False, depth_lo=1.0, depth_hi=10.0)
    # full step
    if want("tiny4"):
        run_full("tiny4", O.Topology(height=64, width=128), B=2,
                 grads_of=("depth_backbone.conv1.weight", "head.depth_decoder.decoder.13.weight", "head.depth_decoder.decoder.9.sequence.1.weight"))
    if want("cfg1"):
        run_full("cfg1", O.Topology(height=128, width=416, scales=(0,)), B=2)
    if want("tiny_pose"):
        run_full("tiny_pose", O.Topology(height=64, width=128, posenet=True, overlapped_mask=False), B=2,
                 grads_of=("head.pose_decoder.net.3.weight", "pose_backbone.conv1.weight"))
    if want("tiny_sigmoid"):
        run_full("tiny_sigmoid", O.Topology(height=64, width=96, multi_channel=False, n_bins=1, min_depth=0.1, scales=(0, 1, 2, 3)), B=2)
    if want("tiny_fe"):
        run_full("tiny_fe", O.Topology(height=64, width=64, fisheye=True, n_bins=64, max_depth=150.0), B=2)
    if want("tiny_distill"):
        run_full("tiny_distill", O.Topology(height=64, width=128, distill=True), B=2,
                 grads_of=("head.depth_decoder.decoder.14.weight", "head.depth_decoder.decoder.17.bias", "depth_backbone.conv1.weight"))
    if want("tiny_normeval"):       # ResNet constructor options: BatchNorms of the encoder in eval mode / first stages frozen
        run_full("tiny_normeval", O.Topology(height=64, width=128, norm_eval=True), B=2,
                 grads_of=("depth_backbone.layer1.0.bn1.weight", "depth_backbone.conv1.weight"))
    if want("tiny_normeval_frozen"):
        run_full("tiny_normeval_frozen", O.Topology(height=64, width=128, norm_eval=True, frozen_stages=1), B=2,
                 grads_of=("depth_backbone.layer2.0.bn1.weight", "depth_backbone.layer2.0.conv1.weight"))
    if want("tiny_frozen"):
        run_full("tiny_frozen", O.Topology(height=64, width=128, frozen_stages=2), B=2,
                 grads_of=("depth_backbone.layer3.0.bn1.weight", "depth_backbone.layer3.0.downsample.0.weight"))
    # 32x64 twins of two cases above: what the CPU suite runs through the emulated executor by default (3x cheaper)
    if want("micro_distill"):
        run_full("micro_distill", O.Topology(height=32, width=64, distill=True), B=2, store_disp=False,
                 grads_of=("head.depth_decoder.decoder.14.weight", "head.depth_decoder.decoder.17.bias", "depth_backbone.conv1.weight"))
    if want("micro_normeval_frozen"):
        run_full("micro_normeval_frozen", O.Topology(height=32, width=64, norm_eval=True, frozen_stages=1), B=2, store_disp=False,
                 grads_of=("depth_backbone.layer2.0.bn1.weight", "depth_backbone.layer2.0.conv1.weight"))
    if want("tiny_r50"):
        run_full("tiny_r50", O.Topology(height=64, width=96, depth=50, base_fx=40.0), B=2, store_disp=True)
