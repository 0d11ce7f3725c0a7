"""In-tree build of the native pieces (explicit nvcc / gcc command lines; outputs under lib/).

  lib/libcpb200.so         csrc/world.cu  -> the CUDA step engine + C ABI (sm_100a only)
  lib/libchipmunk_b200.so  host/*.c       -> the Chipmunk2D public C API on top of the C ABI
  lib/libscene_b200.so     scenes/scene_io.c linked against libchipmunk_b200.so
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "lib")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]
CC_FLAGS = ["-O2", "-std=gnu99", "-ffp-contract=off", "-fopenmp", "-fPIC", "-fvisibility=hidden", "-Wall", "-Wextra",
            "-Wno-unused-parameter", "-DNDEBUG"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build_engine(force=False):
    os.makedirs(LIB, exist_ok=True)
    src = os.path.join(HERE, "csrc", "world.cu")
    deps = glob.glob(os.path.join(HERE, "csrc", "*")) + [os.path.join(ROOT, "include", "cpb200.h")]
    out = os.path.join(LIB, "libcpb200.so")
    if force or _newer(out, deps):
        _run(["nvcc"] + NVCC_FLAGS + ["-o", out, src])
    return out


def build_host(force=False):
    os.makedirs(LIB, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(HERE, "host", "*.c")))
    if not srcs:
        return None
    deps = srcs + glob.glob(os.path.join(HERE, "host", "*.h")) + glob.glob(os.path.join(ROOT, "include", "chipmunk", "*.h")) + \
        [os.path.join(ROOT, "include", "cpb200.h")]
    out = os.path.join(LIB, "libchipmunk_b200.so")
    if force or _newer(out, deps):
        _run(["gcc"] + CC_FLAGS + ["-shared", "-o", out, "-I", os.path.join(ROOT, "include"), "-I", os.path.join(HERE, "host")] + srcs +
             ["-L", LIB, "-lcpb200", "-Wl,-rpath,$ORIGIN", "-lm", "-lpthread"])
    scene_src = os.path.join(HERE, "scenes", "scene_io.c")
    out2 = os.path.join(LIB, "libscene_b200.so")
    if force or _newer(out2, [scene_src, out, os.path.join(HERE, "scenes", "cpb_scene.h")]):
        _run(["gcc", "-O2", "-std=gnu99", "-fPIC", "-w", "-shared", "-o", out2, "-I", os.path.join(ROOT, "include"),
              "-I", os.path.join(HERE, "scenes"), scene_src, "-L", LIB, "-lchipmunk_b200", "-Wl,-rpath,$ORIGIN", "-lm"])
    return out


def build_oracle():
    """Test infrastructure: the C restatement and (where /root/reference exists) oracle/_ref."""
    _run(["make", "-C", os.path.join(ROOT, "oracle"), "all"])


def build_all(force=False):
    build_engine(force)
    build_host(force)
    build_oracle()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
