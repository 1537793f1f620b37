"""Walk a full training step of a shipped configuration through conv_tc.cu's host-side planner on the CPU (kernels skipped):

    python tests/host_emulation/plan_walk.py [cfg2a cfg3 nusc ...]            # all cases without arguments
    FSNET_CONV_FOLD=1 FSNET_WGRAD_OCC=2 python tests/host_emulation/plan_walk.py cfg2a     # under the A/B switches

The planner reads its environment switches once per process, hence a script rather than a parametrised test.  Same mechanism as
tests/test_emulated_kernels_cpu.py::test_conv_planner_accepts_every_shipped_configuration."""
import ctypes
import os
import sys
import time

os.environ["FSNET_EMULATE_PLAN_ONLY"] = "1"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), os.path.dirname(os.path.dirname(HERE))]


def main(names):
    from _pytest.monkeypatch import MonkeyPatch
    from host_emulation import fixture
    patch = MonkeyPatch()
    lib = fixture.install(patch)
    try:
        from fsnet_b200.networks import ops
        from helpers import build_model
        from oracle import fsnet_oracle as O
        from test_emulated_kernels_cpu import PLAN_CASES
        ops.set_backend("tc")
        counter = ctypes.c_longlong.in_dll(lib, "fsnet_emulated_plans")
        failed = 0
        for name in sorted(PLAN_CASES):
            if names and not any(n in name for n in names):
                continue
            topo, B = PLAN_CASES[name]
            t, before = time.time(), counter.value
            data = (O.synthetic_fisheye_batch if topo.fisheye else O.synthetic_batch)(B, topo.height, topo.width, 1, topo.frame_ids)
            try:
                build_model(topo)(dict(data), dict(is_training=True, epoch_num=0, global_step=0))["loss"].mean().backward()
                print(f"{name}: OK, {counter.value - before} tensor-core launches planned, {time.time() - t:.1f} s", flush=True)
            except Exception as e:  # noqa: BLE001
                failed += 1
                print(f"{name}: FAILED: {str(e)[:300]}", flush=True)
        return failed
    finally:
        patch.undo()


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
