"""Pin the oracle (oracle/fsnet_oracle.py) against outputs of the reference itself.

The fixtures under tests/golden were produced by tests/golden/make_golden.py, which imports and runs
/root/reference.  Tolerances: the oracle uses the same torch CPU kernels as the reference, so the
agreement is to rounding (1e-5 relative); the B200 path is held to 1e-3 elsewhere.
"""
import os

import numpy as np
import pytest
import torch

from oracle import fsnet_oracle as O

torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def checksum(data):
    return np.array([float(v.double().sum()) for k, v in sorted(data.items(), key=lambda kv: str(kv[0])) if torch.is_tensor(v)])


def rel(a, b):
    a = (a.detach() if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a))).double()
    b = (b.detach() if isinstance(b, torch.Tensor) else torch.as_tensor(np.asarray(b))).double()
    return float((a - b).norm() / (b.norm() + 1e-30))


LOSS_CASES = {
    "loss_a": dict(topo=O.Topology(height=96, width=160, overlapped_mask=True), B=3, seed=11),
    "loss_b": dict(topo=O.Topology(height=64, width=96, overlapped_mask=False, scales=(0, 2)), B=2, seed=12, with_mask=False),
    "loss_c": dict(topo=O.Topology(height=64, width=128, overlapped_mask=True), B=2, seed=13, mask_dtype=torch.float32,
                   depth_lo=0.6, depth_hi=6.0),
    "loss_mm": dict(topo=O.Topology(height=64, width=96, overlapped_mask=True, scales=(0, 1)), B=2, seed=14, motion_mask=True),
    # FishEyeDecoder (MEI camera, LUT back-projection), two distinct calibrations in one batch
    "loss_fe": dict(topo=O.Topology(height=96, width=128, overlapped_mask=True, fisheye=True, max_depth=150.0), B=3, seed=15),
    "loss_fe_nomask": dict(topo=O.Topology(height=64, width=64, overlapped_mask=True, fisheye=True, scales=(0, 1)), B=2, seed=16,
                           with_mask=False, depth_lo=1.0, depth_hi=10.0),
}


def build_loss_case(topo, B, seed, mask_dtype=torch.float64, with_mask=True, motion_mask=False, depth_lo=2.0, depth_hi=40.0):
    if topo.fisheye:
        data = O.synthetic_fisheye_batch(B, topo.height, topo.width, seed, topo.frame_ids, mask_dtype=mask_dtype, two_calibrations=True)
    else:
        data = O.synthetic_batch(B, topo.height, topo.width, seed, topo.frame_ids, mask_dtype=mask_dtype)
    if not with_mask:
        del data["patched_mask"]
    outputs = O.synthetic_depth_outputs(B, topo.height, topo.width, topo.scales, seed + 1, depth_lo, depth_hi,
                                        topo.min_depth, topo.max_depth)
    if motion_mask:
        data["motion_mask"] = O.synthetic_motion_mask(B, topo.height, topo.width, seed + 2)
    noise = O.tie_break_noise(B, topo.height, topo.width, topo.scales, 0)
    return data, outputs, noise


@pytest.mark.parametrize("name", sorted(LOSS_CASES))
def test_loss_chain_matches_reference(golden_dir, name):
    g = load(golden_dir, name)
    case = LOSS_CASES[name]
    topo = case["topo"]
    data, outputs, noise = build_loss_case(**case)
    np.testing.assert_allclose(checksum(data), g["input_checksum"], rtol=1e-12)
    for v in outputs.values():
        v.requires_grad_(True)
    cam_T = {f: data[("relative_pose", f)].clone().requires_grad_(True) for f in topo.frame_ids[1:]}
    ret = O.loss_chain(outputs, data, cam_T, topo, noise, keep=True)
    ret["loss"].backward()
    assert abs(float(ret["loss"].detach()) - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    for k, v in ret["loss_dict"].items():
        assert abs(float(v) - float(g["loss_dict/" + k])) <= 1e-5 * abs(float(g["loss_dict/" + k])) + 1e-12, k
    for s in topo.scales:
        assert rel(outputs[("depth", s, s)].grad, g[f"grad_depth/{s}"]) < 1e-4, s
        assert rel(outputs[("disp", s)].grad, g[f"grad_disp/{s}"]) < 1e-4, s
    for f in topo.frame_ids[1:]:
        assert rel(cam_T[f].grad, g[f"grad_T/{f}"]) < 1e-4, f
    if "loss_mask_0" in g.files and "motion_mask" not in data:
        mine = np.packbits((ret["aux"][("idxs", 0)] >= 2)[0:1].unsqueeze(1).numpy().astype(np.uint8))
        assert np.array_equal(mine, g["loss_mask_0"])


FULL_CASES = {
    "tiny4": dict(topo=O.Topology(height=64, width=128), B=2),
    "cfg1": dict(topo=O.Topology(height=128, width=416, scales=(0,)), B=2),
    "tiny_pose": dict(topo=O.Topology(height=64, width=128, posenet=True, overlapped_mask=False), B=2),
    "tiny_sigmoid": dict(topo=O.Topology(height=64, width=96, multi_channel=False, n_bins=1, min_depth=0.1, scales=(0, 1, 2, 3)), B=2),
    "tiny_r50": dict(topo=O.Topology(height=64, width=96, depth=50, base_fx=40.0), B=2),
    "tiny_fe": dict(topo=O.Topology(height=64, width=64, fisheye=True, n_bins=64, max_depth=150.0), B=2),
}
# oracle pinned here on CPU; the CUDA side of these runs from tests/test_callers_gpu.py
PENDING_FULL_CASES = {
    # second training stage: frozen eval-mode teacher, uncertainty heads, distillation loss (DistillWPoseMeta)
    "tiny_distill": dict(topo=O.Topology(height=64, width=128, distill=True), B=2),
    # ResNet constructor options of the reference: encoder BatchNorms in eval mode while training / frozen first stages
    "tiny_normeval": dict(topo=O.Topology(height=64, width=128, norm_eval=True), B=2),
    "tiny_frozen": dict(topo=O.Topology(height=64, width=128, frozen_stages=2), B=2),
    "tiny_normeval_frozen": dict(topo=O.Topology(height=64, width=128, norm_eval=True, frozen_stages=1), B=2),
}
# 32x64 twins for the emulated-executor tests of the CPU suite (tests/test_emulated_kernels_cpu.py)
MICRO_FULL_CASES = {
    "micro_distill": dict(topo=O.Topology(height=32, width=64, distill=True), B=2),
    "micro_normeval_frozen": dict(topo=O.Topology(height=32, width=64, norm_eval=True, frozen_stages=1), B=2),
}
ALL_FULL_CASES = dict(FULL_CASES, **PENDING_FULL_CASES, **MICRO_FULL_CASES)


@pytest.mark.parametrize("name", sorted(ALL_FULL_CASES))
def test_full_step_matches_reference(golden_dir, name):
    g = load(golden_dir, name)
    topo, B = ALL_FULL_CASES[name]["topo"], ALL_FULL_CASES[name]["B"]
    data = (O.synthetic_fisheye_batch if topo.fisheye else O.synthetic_batch)(B, topo.height, topo.width, 1234, topo.frame_ids)
    np.testing.assert_allclose(checksum(data), g["input_checksum"], rtol=1e-12)
    sd = O.make_state_dict(topo)
    names = O.trainable(sd, topo)
    for k in names:
        sd[k].requires_grad_(True)
    noise = O.tie_break_noise(B, topo.height, topo.width, topo.scales, 0)
    ret = O.forward_train(sd, data, topo, noise)
    assert (ret["loss"].dtype == torch.float64) == bool(g["loss_is_fp64"])
    assert abs(float(ret["loss"].detach()) - float(g["loss"])) <= 2e-6 * abs(float(g["loss"]))
    for k, v in ret["loss_dict"].items():
        assert abs(float(v) - float(g["loss_dict/" + k])) <= 1e-4 * abs(float(g["loss_dict/" + k])) + 1e-12, k
    for s in topo.scales:
        if f"disp/{s}" not in g.files:
            continue
        assert rel(ret["outputs"][("disp", s)], g[f"disp/{s}"]) < 1e-5
        assert rel(ret["outputs"][("depth", s, s)], g[f"depth/{s}"]) < 1e-5
        if topo.distill and f"uncertain_z/{s}" in g.files:
            assert rel(ret["outputs"][("uncertain_z", s)], g[f"uncertain_z/{s}"]) < 1e-5
        if topo.distill and f"teacher_depth/{s}" in g.files:
            assert rel(ret["outputs"][("teacher_depth", s, s)], g[f"teacher_depth/{s}"]) < 1e-5
    if topo.distill:
        assert not any(k.startswith("teacher_net.") for k in names) and all(not n.startswith("teacher_net.") for n in g["grad_names"].tolist())
    ret["loss"].backward()
    gn = dict(zip(g["grad_names"].tolist(), g["grad_norms"].tolist()))
    assert set(names) == set(gn), set(names) ^ set(gn)          # exactly the parameters the reference trains
    floor = 1e-9 + 1e-8 * max(gn.values())        # conv biases in front of a BatchNorm have an exactly-zero gradient: rounding noise only
    for k in names:
        if k in gn:
            mine = float(sd[k].grad.double().norm())
            assert abs(mine - gn[k]) <= 2e-3 * gn[k] + floor, (k, mine, gn[k])
    for key in g.files:
        if key.startswith("grad/"):
            assert rel(sd[key[5:]].grad, g[key]) < 2e-3, key
    if topo.posenet:
        for f in topo.frame_ids[1:]:
            assert rel(ret["cam_T"][f], g[f"cam_T_cam/{f}"]) < 1e-5
    # eval-mode prediction; the train-mode forward above updated the running statistics once, as in the reference run
    sd2 = O.make_state_dict(topo)
    with torch.no_grad():
        O.forward_train(sd2, data, topo, noise)
        pred = O.forward_test(sd2, data, topo)
    # the reference ran its train-mode forward on model2's backbone + decoder only (no PoseNet influence)
    assert rel(pred["depth"], g["test_depth"]) < 1e-4
    if topo.fisheye:
        assert rel(pred["norm"], g["test_norm"]) < 1e-4
