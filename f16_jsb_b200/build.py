"""Builds libf16b200.so (the CUDA library behind include/f16_b200.h) in-tree with nvcc for sm_100a."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libf16b200.so")
SOURCES = [os.path.join(_HERE, "csrc", "f16_b200.cu"), os.path.join(_HERE, "csrc", "f16_rollout.cu"),
           os.path.join(_HERE, "csrc", "f16_features.cu"), os.path.join(_HERE, "csrc", "f16_hostwin.cu"),
           os.path.join(_HERE, "csrc", "f16_lma_attention.cu"), os.path.join(_HERE, "csrc", "f16_lma_norm.cu"),
           os.path.join(_HERE, "csrc", "f16_lma_wgrad.cu"), os.path.join(_HERE, "csrc", "f16_lma_linear.cu"), os.path.join(_HERE, "csrc", "f16_lma_wgrad_tc.cu"), os.path.join(_HERE, "csrc", "f16_lma_elementwise.cu"),
           os.path.join(_HERE, "csrc", "f16_lma_policy.cu")]
DEPS = [os.path.join(_HERE, "csrc", f) for f in
        ("f16_b200.cu", "f16_rollout.cu", "f16_features.cu", "f16_hostwin.cu", "f16_lma_attention.cu", "f16_lma_norm.cu", "f16_lma_wgrad.cu", "f16_lma_linear.cu", "f16_lma_wgrad_tc.cu", "f16_lma_elementwise.cu", "f16_lma_policy.cu", "f16_tc_common.cuh", "f16_model.cuh", "f16_ground.cuh", "f16_env.cuh", "f16_host_setup.h", "f16_model_data.h")] + [
    os.path.join(os.path.dirname(_HERE), "include", f) for f in ("f16_b200.h", "f16_rollout.h", "f16_features.h", "f16_hostwin.h", "f16_lma.h", "f16_state_fields.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the F-16 env has no CPU fallback and needs the CUDA toolkit to build")
    return p


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """One process per GPU means several ranks may get here at once (torchrun): the build runs under a file lock, into a
    temporary file that is renamed over the library in one step, so no rank ever dlopens a half-written file and only
    the first rank compiles."""
    import fcntl
    import tempfile
    if not (force or needs_build()):
        return LIB_PATH
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or needs_build():          # another rank may have built it while this one waited
                fd, tmp = tempfile.mkstemp(prefix=".libf16b200.", suffix=".so.tmp", dir=_HERE)
                os.close(fd)
                try:
                    cmd = [nvcc_path()] + NVCC_FLAGS + ["--threads", "0"] + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + SOURCES
                    subprocess.check_call(cmd)
                    os.chmod(tmp, 0o755)
                    os.replace(tmp, LIB_PATH)
                finally:
                    if os.path.exists(tmp):
                        os.unlink(tmp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
