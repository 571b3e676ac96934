"""Build libldiff_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m ldiffusion_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libldiff_sm100.so")
SOURCES = ["capi.cu", "sampler.cu", "decode_tail.cu", "bilinear.cu", "bilinear_bwd.cu", "head.cu", "lift_argmax_env.cu", "lift_argmax_row.cu", "head_tc.cu", "confusion.cu", "sliding.cu", "infonce.cu"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "ldiff.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu for sm_100a (one nvcc per file, in parallel) and link the shared library."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    obj_dir = os.path.join(HERE, "_build")
    os.makedirs(obj_dir, exist_ok=True)
    common = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if verbose:
        common += ["-Xptxas", "-v"]

    def compile_one(src):
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        res = subprocess.run(common + ["-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        return src, obj, res

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    log = "".join(r.stdout + r.stderr for _, _, r in results)
    if any(r.returncode for _, _, r in results):
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libldiff_sm100.so")
    link = subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "--shared",
                           *[o for _, o, _ in results], "-o", LIB], capture_output=True, text=True)
    if link.returncode != 0:
        sys.stderr.write(link.stdout + link.stderr)
        raise RuntimeError("linking libldiff_sm100.so failed")
    if verbose:
        sys.stderr.write(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
