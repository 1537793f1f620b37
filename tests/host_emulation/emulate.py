"""Runs the DEVICE code of a simple (shuffle-free, shared-memory-free) CUDA kernel on the CPU: the text between the anonymous
namespace braces of the .cu file is compiled by g++ behind a small shim (CUDA qualifiers -> nothing, round-to-nearest intrinsics
-> plain IEEE operations with FMA contraction switched off, blockIdx / threadIdx -> globals) and driven by a loop over the grid.

Test infrastructure: lets the index arithmetic and branch logic of a kernel be checked against its oracle in the CPU suite, before
(and in addition to) the -m gpu parity test of the real launch."""
import ctypes
import hashlib
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))

SHIM = r'''
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>
#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(x)
struct Idx { int x, y, z; };
static Idx blockIdx, blockDim, threadIdx, gridDim;
using std::max;
using std::min;
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline long long __double2ll_rn(double a) { return llrint(a); }
'''


def device_source(cu_path):
    text = open(cu_path).read()
    start = text.index("namespace {") + len("namespace {")
    end = text.index("}  // namespace", start)
    return text[start:end]


def build(cu_name, driver_cpp, tag):
    """-> ctypes.CDLL of shim + device code of fsnet_b200/csrc/<cu_name> + driver (cached by content hash in the temp dir)."""
    src = SHIM + "\nnamespace {\n" + device_source(os.path.join(REPO, "fsnet_b200", "csrc", cu_name)) + "\n}\n" + driver_cpp
    key = hashlib.sha1(src.encode()).hexdigest()[:16]
    out = os.path.join(tempfile.gettempdir(), f"fsnet_emul_{tag}_{key}.so")
    if not os.path.exists(out):
        cpp = out[:-3] + ".cpp"
        with open(cpp, "w") as f:
            f.write(src)
        subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-fno-fast-math", "-w", "-shared", "-fPIC", cpp, "-o", out])
    return ctypes.CDLL(out)


AUGMENT_DRIVER = r'''
extern "C" void run_augment(const uint8_t* frames, const uint8_t* mask, const double* plan, int B, int F, int H0, int W0, int H, int W,
                            const float* mean_std, float* image, float* original, double* mask_out) {
  blockDim = {32, 8, 1};
  gridDim = {(W + 31) / 32, (H + 7) / 8, B};
  for (blockIdx.z = 0; blockIdx.z < gridDim.z; ++blockIdx.z)
    for (blockIdx.y = 0; blockIdx.y < gridDim.y; ++blockIdx.y)
      for (blockIdx.x = 0; blockIdx.x < gridDim.x; ++blockIdx.x)
        for (threadIdx.y = 0; threadIdx.y < blockDim.y; ++threadIdx.y)
          for (threadIdx.x = 0; threadIdx.x < blockDim.x; ++threadIdx.x)
            augment_frames_kernel(frames, mask, plan, B, F, H0, W0, H, W, mean_std, image, original, mask_out);
}
'''


def run_augment(frames, mask, plan, H, W, mean_std):
    """numpy in / out twin of DeviceAugmentStage's launch: frames [B,F,H0,W0,3] uint8, mask [B,H0,W0] uint8, plan [B,16] float64."""
    import numpy as np
    lib = build("augment.cu", AUGMENT_DRIVER, "augment")
    B, F, H0, W0, _ = frames.shape
    frames, mask, plan = np.ascontiguousarray(frames), np.ascontiguousarray(mask), np.ascontiguousarray(plan, dtype=np.float64)
    mean_std = np.ascontiguousarray(mean_std, dtype=np.float32)
    image = np.empty((F, B, 3, H, W), dtype=np.float32)
    original = np.empty((F, B, 3, H, W), dtype=np.float32)
    mask_out = np.empty((B, H, W), dtype=np.float64)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)          # noqa: E731
    lib.run_augment(ptr(frames), ptr(mask), ptr(plan), B, F, H0, W0, H, W, ptr(mean_std), ptr(image), ptr(original), ptr(mask_out))
    return image, original, mask_out


# The block-level reduction of distill_loss_kernel needs warp shuffles; the per-pixel arithmetic does not: the kernel is emulated as
# "one thread, one block" (the shims below make the reduction scaffold the identity), which exercises exactly the device code that
# is new in csrc/distill.cu -- distill_term and the grid-stride loop.
DISTILL_SHIM = r'''
#define __shared__ static
static inline void __syncthreads() {}
static inline double warp_sum(double v) { return v; }
static inline float ldg(const float* p) { return *p; }
static inline void atomicAdd(double* p, double v) { *p += v; }
'''
DISTILL_DRIVER = r'''
extern "C" void run_distill(const float* pred, const float* teacher, const float* ulogit, long long n, double* out, float* gp, float* gl,
                            float* u) {
  blockIdx = {0, 0, 0}; threadIdx = {0, 0, 0}; gridDim = {1, 1, 1};
  distill_loss_kernel(pred, teacher, ulogit, n, 1.f / (float)n, out, gp, gl, u);
}
'''


def run_distill(pred, teacher, ulogit):
    import numpy as np
    text = device_source(os.path.join(REPO, "fsnet_b200", "csrc", "distill.cu"))
    # one emulated thread walks the whole array: the grid-stride step becomes 1 (kThreads only appears in the launch geometry)
    text = text.replace("constexpr int kThreads = 256;", "constexpr int kThreads = 1;").replace("kThreads / 32", "1")
    src = SHIM + DISTILL_SHIM + "\nnamespace {\n" + text + "\n}\n" + DISTILL_DRIVER
    key = hashlib.sha1(src.encode()).hexdigest()[:16]
    out_so = os.path.join(tempfile.gettempdir(), f"fsnet_emul_distill_{key}.so")
    if not os.path.exists(out_so):
        cpp = out_so[:-3] + ".cpp"
        with open(cpp, "w") as f:
            f.write(src)
        subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-fno-fast-math", "-w", "-shared", "-fPIC", cpp, "-o", out_so])
    lib = ctypes.CDLL(out_so)
    lib.run_distill.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_longlong] + [ctypes.c_void_p] * 4
    pred, teacher = np.ascontiguousarray(pred, dtype=np.float32), np.ascontiguousarray(teacher, dtype=np.float32)
    n = pred.size
    out = np.zeros(1, dtype=np.float64)
    gp, gl, u = np.empty(n, np.float32), np.empty(n, np.float32), np.empty(n, np.float32)
    ptr = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)          # noqa: E731
    ul = None if ulogit is None else np.ascontiguousarray(ulogit, dtype=np.float32)
    lib.run_distill(ptr(pred), ptr(teacher), ptr(ul), n, ptr(out), ptr(gp), ptr(gl) if ul is not None else None,
                    ptr(u) if ul is not None else None)
    return float(out[0]), gp, (gl if ul is not None else None), (u if ul is not None else None)


# ----------------------------------------------------------------------------------------------------------------------
# Whole-file emulation: the loss-side .cu files (host entry points included) compiled for the CPU on top of simt.h, exporting the
# same C ABI as libfsnet_b200.so but over HOST pointers.
# ----------------------------------------------------------------------------------------------------------------------
SIMT_FILES = ["abi.cu", "warp_ssim.cu", "smooth_head.cu", "optim.cu", "distill.cu", "augment.cu", "act_tc.cu"]
# conv_tc.cu (tcgen05 / TMA) can not be emulated: conv_ref.cpp implements its two entry points from the ABI contract


def _split_top_level(text):
    parts, depth, cur = [], 0, ""
    for ch in text:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def translate(cu_text, csrc_dir):
    """CUDA C++ -> C++ over simt.h: inline the local .cuh includes, swap the CUDA headers for simt.h, rewrite kernel launches."""
    import re

    def include(m):
        name = m.group(1)
        if name.endswith(".cuh"):
            return translate(open(os.path.join(csrc_dir, name)).read().replace("#pragma once", ""), csrc_dir)
        return m.group(0)

    text = re.sub(r'#include "([^"]+)"', include, cu_text)
    text = re.sub(r"#include <cuda_runtime\.h>", '#include "simt.h"', text)
    text = re.sub(r"\b__(exp|log|pow)f\(", r"\1f(", text)          # fast-math approximations -> the accurate libm functions
    # dynamic shared memory: `extern __shared__ T name[];` -> a pointer into the launch's buffer
    text = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w:]+)\s+(\w+)\[\];", r"\1* \2 = (\1*)simt::dyn_smem();", text)
    if re.search(r"#include <(cuda(?!_bf16|\.h)|mma|cooperative)", text):
        raise NotImplementedError("this file needs CUDA headers the SIMT shim does not provide")
    out, pos = "", 0
    for m in re.finditer(r"([A-Za-z_]\w*(?:<[^<>;()]*>)?)\s*<<<(.*?)>>>\s*\(", text, re.S):
        cfg = _split_top_level(m.group(2))
        rest = text[m.end():].lstrip()
        sep = "" if rest.startswith(")") else ", "
        smem = cfg[2] if len(cfg) > 2 else "0"
        out += text[pos:m.start()] + f"simt::launch({m.group(1)}, dim3({cfg[0]}), dim3({cfg[1]}), (size_t)({smem}){sep}"
        pos = m.end()
    return out + text[pos:]


def build_simt_library(files=SIMT_FILES):
    csrc = os.path.join(REPO, "fsnet_b200", "csrc")
    sources = {f: translate(open(os.path.join(csrc, f)).read(), csrc) for f in files}
    sources["conv_tc_plan.cu"] = translate_plan(open(os.path.join(csrc, "conv_tc.cu")).read(), csrc)
    extra = "".join(open(os.path.join(HERE, f)).read() for f in ("simt.h", "cuda_bf16.h", "cuda.h", "conv_ref.cpp", "simt_switch.cpp"))
    key = hashlib.sha1(("".join(sources.values()) + extra).encode()).hexdigest()[:16]
    out = os.path.join(tempfile.gettempdir(), f"fsnet_simt_{key}.so")
    if not os.path.exists(out):
        # private work directory + atomic rename: several processes (the world-2 tests) may ask for the library at the same time
        work = tempfile.mkdtemp(prefix=f"fsnet_simt_{key}_")
        objs = []
        for f, text in sources.items():
            cpp = os.path.join(work, f[:-3] + ".cpp")
            with open(cpp, "w") as fh:
                fh.write(text)
            obj = cpp[:-4] + ".o"
            subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-w", "-fPIC", "-I", HERE, "-I", os.path.join(REPO, "include"),
                                   "-I", csrc, "-c", cpp, "-o", obj])
            objs.append(obj)
        ref = os.path.join(work, "conv_ref.o")
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-w", "-fPIC", "-I", os.path.join(REPO, "include"), "-c", os.path.join(HERE, "conv_ref.cpp"),
                               "-o", ref])
        switch = os.path.join(work, "simt_switch.o")
        subprocess.check_call(["g++", "-fPIC", "-c", os.path.join(HERE, "simt_switch.cpp"), "-o", switch])
        staged = os.path.join(work, "lib.so")
        subprocess.check_call(["g++", "-shared", "-o", staged] + objs + [ref, switch])
        os.replace(staged, out)
    return out


# ----------------------------------------------------------------------------------------------------------------------
# Plan check of the tensor-core entry points: the HOST code of conv_tc.cu (argument checks, tile / pipeline / K-split planning,
# tensor-map encoding against a validating cuTensorMapEncodeTiled, tests/host_emulation/cuda.h) with every device function
# removed and the launches dropped, exported as fsnet_conv_plan / fsnet_conv_wgrad_plan; conv_ref.cpp calls them first.
# ----------------------------------------------------------------------------------------------------------------------
def _strip_device_functions(text):
    """Remove every `__device__` function; keep `__global__` kernels as empty shells (the host code takes their address)."""
    import re
    out, pos = "", 0
    pat = re.compile(r"(template\s*<[^>]*>\s*)?(__device__|__global__)")
    while True:
        m = pat.search(text, pos)
        if not m:
            return out + text[pos:]
        brace = text.index("{", m.end())
        depth, i = 0, brace
        while True:
            depth += text[i] == "{"
            depth -= text[i] == "}"
            i += 1
            if depth == 0:
                break
        out += text[pos:m.start()]
        if m.group(2) == "__global__":
            out += text[m.start():brace] + "{}"
        pos = i


def translate_plan(cu_text, csrc_dir):
    import re
    text = _strip_device_functions(cu_text)
    text = translate(text, csrc_dir)
    text = text.replace("simt::launch(", "plan_record(")
    recorder = r"""
// plan check: what would have been launched (FSNET_PLAN_PRINT=1 prints one line per planned launch)
template <class K> static void plan_record(K, dim3 grid, dim3 block, size_t smem, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap,
                                           const fsnet::ConvParams& p) {
  const char* e = getenv("FSNET_PLAN_PRINT");
  if (!e || e[0] != '1') return;
  printf("PLAN conv  N=%d %dx%d Cin=%d Cout=%d k=%d s=%d | tile %dx%d BN=%d fold=%d KC=%d kiters=%d stages=%d | tiles=%d cluster=%dx%d grid=%u smem=%zu\n",
         p.N, p.H, p.W, p.Cin, p.Cout, p.KH, p.stride, p.TH, p.TW, p.BN, p.fold, p.KC, p.kiters, p.stages, p.total_tiles, p.cm, p.cn, grid.x, smem);
}
"""
    text = text.replace('extern "C" int fsnet_conv(', recorder + 'extern "C" int fsnet_conv(', 1)
    recorder_w = r"""
template <class K> static void plan_record(K, dim3 grid, dim3 block, size_t smem, CUtensorMap, CUtensorMap, const fsnet::WgradParams& p) {
  const char* e = getenv("FSNET_PLAN_PRINT");
  if (!e || e[0] != '1') return;
  printf("PLAN wgrad N=%d %dx%d Cin=%d Cout=%d k=%d s=%d | BM=%d BN=%d fold=%d pix=%d ksplit=%d stages=%d | items=%d grid=%u smem=%zu\n",
         p.N, p.Ho, p.Wo, p.Cin, p.Cout, p.KH, p.stride, p.BM_real, p.BN, p.fold, p.pix, p.ksplit, p.stages,
         p.taps * p.co_tiles * p.ci_tiles, grid.x, smem);
}
"""
    text = text.replace('extern "C" int fsnet_conv_wgrad(', recorder_w + 'extern "C" int fsnet_conv_wgrad(', 1)
    text = text.replace("#include <stdlib.h>", "#include <stdlib.h>\n#include \"cuda.h\"").replace("#include <cuda.h>", "")
    text = re.sub(r"cudaGetDriverEntryPoint\([^;]*\)\s*!=\s*cudaSuccess", "((ptr = (void*)&cuTensorMapEncodeTiled), false)", text)
    return text.replace('extern "C" int fsnet_conv(', 'extern "C" int fsnet_conv_plan(').replace(
        'extern "C" int fsnet_conv_wgrad(', 'extern "C" int fsnet_conv_wgrad_plan(')
