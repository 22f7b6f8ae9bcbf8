"""Builds posepipeline_b200/libposeengine.so for sm_100a with nvcc (in-tree; the .so travels to the GPU box).

    python -m posepipeline_b200.csrc.build [--force]
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
SOURCES = ["conv_tc_inst_a.cu", "conv_tc_inst_b.cu", "conv_tc_inst_c.cu", "conv_tc_inst_d.cu", "conv_tc_inst_e.cu", "conv_tc_inst_f.cu",
           "engine.cu", "kernels_simt.cu", "decode.cu", "lifter.cu", "conv_tc.cu", "bytetrack.cu", "detector.cu", "vit.cu"]
HEADERS = ["kernels.h", "pe_common.cuh", "conv_tc_kernel.cuh", "engine_internal.h", os.path.join("..", "..", "include", "poseengine.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# PE_PRECISION=tf32 builds the wide-range TF32x3 variant (8 B/element); default fp16x2 (4 B/element, half the MMAs)
PRECISION = os.environ.get("PE_PRECISION", "fp16")
OUT = os.path.join(PKG, "libposeengine.so" if PRECISION == "fp16" else "libposeengine_tf32.so")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fno-fast-math", "--fmad=true", f"-DPE_FP16={1 if PRECISION == 'fp16' else 0}"]
FLAGS += os.environ.get("PE_EXTRA_NVCC_FLAGS", "").split()      # e.g. -DPE_TC_PROFILE=1 (cycle counters in conv_tc)


def _digest():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        h.update(open(os.path.join(HERE, f), "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    stamp = OUT + ".stamp"
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == dig:
        return OUT
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(HERE, src.replace(".cu", f".{PRECISION}.o"))
        cmd = [NVCC, *FLAGS, "-c", os.path.join(HERE, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(f"--- nvcc {src}\n{out}", file=sys.stderr)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [NVCC, "-shared", "-o", OUT, *objs]
    subprocess.run(cmd, check=True)
    open(stamp, "w").write(dig)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
