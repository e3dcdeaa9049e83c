"""TEST INFRASTRUCTURE ONLY: compiles the product's kernel sources with g++ against
tests/emu/cuda_emu.h (host-thread SIMT emulator) into tests/emu/_build/libvqvae3d_emu.so so
that kernel logic can be checked against the oracle without a GPU.  Never used by the
product package."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "3d-vq-vae-2_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libvqvae3d_emu.so")
# sources that contain tcgen05 / TMA inline PTX cannot be emulated
SKIP = set()   # tc_kernels.cu compiles to an UNSUPPORTED stub under VQ3D_EMU


def build(force=False):
    srcs = [s for s in sorted(glob.glob(os.path.join(CSRC, "*.cu"))) if os.path.basename(s) not in SKIP]
    deps = srcs + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "cuda_emu.h"), os.path.join(ROOT, "include", "vqvae3d_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) < os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    objs, procs = [], []
    for s in srcs:
        o = os.path.join(HERE, "_build", os.path.basename(s)[:-3] + ".o")
        cmd = ["g++", "-std=c++20", "-O1", "-g", "-x", "c++", "-DVQ3D_EMU", "-I", HERE, "-I", CSRC, "-fPIC", "-pthread",
               "-ffp-contract=off", "-Wno-unknown-pragmas", "-c", s, "-o", o]
        procs.append((s, subprocess.Popen(cmd)))
        objs.append(o)
    for s, p in procs:
        if p.wait():
            raise RuntimeError("emulator build failed on " + s)
    subprocess.run(["g++", "-shared", "-pthread", "-o", OUT] + objs, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
