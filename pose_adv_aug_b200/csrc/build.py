"""Build libhgk.so (sm_100a only) in-tree with nvcc.  Invoked by __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libhgk.so")
SOURCES = ["conv_simt.cu", "conv_skinny.cu", "conv_tc.cu", "conv_tc2.cu", "conv_tc3.cu", "wgrad_tc2.cu", "stem.cu", "bn.cu", "heads.cu", "pointwise.cu", "loss_optim.cu", "eval.cu", "warp.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default", "--use_fast_math=false"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.sep not in c or os.path.exists(c)):
            return c
    return "nvcc"


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(PKG), "include", "hgk.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in FLAGS if not f.startswith("--use_fast_math")]
    flags += os.environ.get("HGK_NVCC_EXTRA", "").split()      # developer experiments (-DKNOB=value)
    if verbose:
        flags = flags + ["-Xptxas", "-v"]
    procs = []
    objs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        hdrs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))]
        hdrs.append(os.path.join(os.path.dirname(PKG), "include", "hgk.h"))
        if (not force) and os.path.exists(o) and os.path.getmtime(o) > max([os.path.getmtime(s)] +
                                                                           [os.path.getmtime(h) for h in hdrs]):
            continue
        cmd = [_nvcc()] + flags + ["-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stdout.write(out.decode())
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s" % " ".join(cmd))
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT] + objs + ["-lcuda"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
