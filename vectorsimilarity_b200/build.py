"""Builds the product libraries in-tree (they travel to the GPU box with the snapshot):

  libvsgpu.so         hand-written sm_100a CUDA kernels + the thin C-ABI of include/vsgpu.h (nvcc)
  libvecsim_b200.so   host C++ index code exporting the VecSimIndex_* C API of include/vecsim_b200.h (g++)

nvcc cross-compiles for sm_100a without a GPU. Nothing under oracle/ is compiled or linked here.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
CUDA_SOURCES = ["vsgpu_api.cu", "vsgpu_exact.cu", "vsgpu_select.cu", "vsgpu_tensor.cu", "vsgpu_hnsw.cu", "vsgpu_scan.cu", "vsgpu_tensor_i8.cu", "vsgpu_shard.cu"]
HOST_SOURCES = ["host/vecsim_flat.cpp", "host/vecsim_flat_multi.cpp", "host/vecsim_flat_sharded.cpp", "host/vecsim_hnsw.cpp", "host/vecsim_tiered.cpp", "host/vecsim_hnsw_file.cpp", "host/vecsim_api.cpp"]
NVCC_FLAGS = ["-std=c++20", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
GXX_FLAGS = ["-std=c++20", "-O2", "-ffp-contract=off", "-fPIC", "-Wall", "-Wextra"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _stamp(paths, extra):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(repr(extra).encode())
    return h.hexdigest()


def _deps():
    out = []
    for base, _, files in os.walk(CSRC):
        out += [os.path.join(base, f) for f in files if f.endswith((".cu", ".cuh", ".cpp", ".h"))]
    out += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    return out


def build(force=False, verbose=False):
    """Compile both libraries if any source changed. Returns the two paths."""
    lib_gpu = os.path.join(HERE, "libvsgpu.so")
    lib_host = os.path.join(HERE, "libvecsim_b200.so")
    stamp_file = os.path.join(HERE, ".build_stamp")
    cuda_sources = [s for s in CUDA_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    stamp = _stamp(_deps(), (NVCC_FLAGS, GXX_FLAGS, cuda_sources))
    if not force and os.path.exists(lib_gpu) and os.path.exists(lib_host) and os.path.exists(stamp_file) \
            and open(stamp_file).read() == stamp:
        return lib_gpu, lib_host
    objdir = os.path.join(ROOT, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    log = open(os.path.join(objdir, "ptxas.log"), "w")
    objs = []
    # one object per translation unit, recompiled only when that source, a shared header or the flags changed
    headers = [p for p in _deps() if p.endswith((".cuh", ".h"))]
    for src in cuda_sources:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        ostamp = _stamp(headers + [os.path.join(CSRC, src)], NVCC_FLAGS)
        if not force and os.path.exists(obj) and os.path.exists(obj + ".stamp") and open(obj + ".stamp").read() == ostamp:
            continue
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((cmd, obj, ostamp, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for cmd, obj, ostamp, p in procs:
        out, _ = p.communicate()
        log.write("$ " + " ".join(cmd) + "\n" + out + "\n")
        with open(obj + ".ptxas.log", "w") as f:
            f.write(out)
        if p.returncode != 0:
            failed = True
            sys.stderr.write(out)
            if os.path.exists(obj + ".stamp"):
                os.remove(obj + ".stamp")
        else:
            with open(obj + ".stamp", "w") as f:
                f.write(ostamp)
            if verbose:
                sys.stderr.write(out)
    log.close()
    if failed:
        raise RuntimeError("nvcc failed (see build/ptxas.log)")
    subprocess.check_call([nvcc, "-shared", "-Xlinker", "-Bsymbolic", "-o", lib_gpu] + objs)
    host = [os.path.join(CSRC, s) for s in HOST_SOURCES]
    subprocess.check_call(["g++"] + GXX_FLAGS + ["-shared", "-o", lib_host] + host +
                          ["-L" + HERE, "-lvsgpu", "-Wl,-rpath,$ORIGIN", "-Wl,-Bsymbolic"])
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return lib_gpu, lib_host


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
