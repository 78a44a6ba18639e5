"""Compile the CUDA extension (C-ABI shared library) in-tree for sm_100a.

The templated kernels are instantiated in separate translation units (csrc/inst.cu compiled once
per entry of VP_KERNEL_GROUPS in csrc/kernel_tables.h) which are built in parallel and linked with
the host side (csrc/vp_*.cu) into libvarpro_b200.so. Objects go to gpurun_out/_obj/ (scratch: git-ignored,
never shipped) and are reused when neither their source, nor a header, nor the flags changed; the .so is git-ignored but travels to the GPU box with the gpurun snapshot.
nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "..", "gpurun_out", "_obj")  # scratch: never pushed to the GPU box, never committed
LIB = os.path.join(HERE, "libvarpro_b200.so")
HEADER = os.path.join(HERE, "..", "include", "varpro_b200.h")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++"]
LINK_LIBS: list[str] = ["-ldl", "-lpthread"]  # NVTX v3 is header-only and dlopens its injection library
# host side of the C ABI (see csrc/vp_internal.h)
HOST_UNITS = ["vp_ctx", "vp_problem", "vp_fit", "vp_batch", "vp_hostbatch", "vp_diag"]


def _groups():
    txt = open(os.path.join(CSRC, "kernel_tables.h")).read()
    return re.findall(r"X\((\w+),\s*(\w+),\s*(\w+),\s*(\d+),\s*(\d+),\s*(\d+)\)", txt)


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".cu"))] + [HEADER]


def _units():
    """(object path, source, extra flags) of every translation unit."""
    units = [(os.path.join(OBJ, f"{u}.o"), os.path.join(CSRC, f"{u}.cu"), []) for u in HOST_UNITS]
    for tag, ctype, dt, n, p, part in _groups():
        variant = re.search(r"(\d+)$", tag)
        flags = [f"-DVP_INST_TAG={tag}", f"-DVP_INST_T={ctype}", f"-DVP_INST_DT={dt}", f"-DVP_INST_N={n}",
                 f"-DVP_INST_P={p}", f"-DVP_INST_PART={part}", f"-DVP_INST_VARIANT={variant.group(1) if variant else 0}"]
        units.append((os.path.join(OBJ, f"inst_{tag}.o"), os.path.join(CSRC, "inst.cu"), flags))
    return units


def _includes(path, part=None, seen=None):
    """Transitive closure of the local #include "..." of `path`. Only the `#if VP_INST_PART == k` ladder of
    csrc/inst.cu is interpreted (with `part`); every other conditional is taken as true."""
    seen = set() if seen is None else seen
    path = os.path.normpath(path)
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    active = [True]  # stack for the VP_INST_PART ladder
    taken = [True]
    for line in open(path):
        ls = line.strip()
        mi = re.match(r"#\s*(if|elif)\s+VP_INST_PART\s*==\s*(\d+)", ls)
        if mi and part is not None:
            hit = int(mi.group(2)) == int(part)
            if mi.group(1) == "if":
                active.append(hit)
                taken.append(hit)
            else:
                active[-1] = hit and not taken[-1]
                taken[-1] = taken[-1] or hit
            continue
        if part is not None and len(active) > 1 and re.match(r"#\s*else\b", ls):
            active[-1] = not taken[-1]
            continue
        if part is not None and len(active) > 1 and re.match(r"#\s*endif\b", ls):
            active.pop()
            taken.pop()
            continue
        m = re.match(r'#\s*include\s+"([^"]+)"', ls)
        if m and all(active):
            _includes(os.path.join(os.path.dirname(path), m.group(1)), None, seen)
    return seen


def _stamp(src=None, flags=()):
    """Hash of everything an object (src given: its source and the headers it includes) or the whole
    library (src None: every source and header) depends on, plus the flags."""
    h = hashlib.sha256()
    if src is None:
        deps = _deps()
    else:
        part = next((f.split("=")[1] for f in flags if f.startswith("-DVP_INST_PART=")), None)
        deps = _includes(src, part)
    for d in sorted(deps):
        h.update(os.path.basename(d).encode())
        h.update(open(d, "rb").read())
    h.update(" ".join(list(NVCC_FLAGS) + list(LINK_LIBS) + list(flags)).encode())
    return h.hexdigest()


def needs_build() -> bool:
    stamp = os.path.join(HERE, "libvarpro_b200.stamp")
    if not os.path.exists(LIB) or not os.path.exists(stamp):
        return True
    return open(stamp).read().strip() != _stamp()


def _compile(nvcc, obj, src, flags, verbose):
    stamp = _stamp(src, flags)
    if os.path.exists(obj) and os.path.exists(obj + ".stamp") and open(obj + ".stamp").read() == stamp and not verbose:
        return 0, "", obj
    r = _compile_run(nvcc, obj, src, flags, verbose)
    if r[0] == 0:
        open(obj + ".stamp", "w").write(stamp)
    return r


def _compile_run(nvcc, obj, src, flags, verbose):
    cmd = [nvcc, *NVCC_FLAGS, *flags, "-c", "-o", obj, src]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    return r.returncode, r.stdout + r.stderr, obj


def build(force: bool = False, verbose: bool = False, jobs: int | None = None) -> str:
    if not force and not needs_build():
        return LIB
    if force:
        for f in (os.listdir(OBJ) if os.path.isdir(OBJ) else []):
            if f.endswith(".stamp"):
                os.remove(os.path.join(OBJ, f))
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    units = _units()
    jobs = jobs or max(1, min(len(units), os.cpu_count() or 1))
    failed = []
    with cf.ThreadPoolExecutor(jobs) as ex:
        for rc, out, obj in ex.map(lambda u: _compile(nvcc, u[0], u[1], u[2], verbose), units):
            if rc != 0:
                failed.append(obj)
                sys.stderr.write(out)
            elif verbose:
                sys.stderr.write(f"== {os.path.basename(obj)}\n{out}")
    if failed:
        raise RuntimeError("nvcc failed building " + ", ".join(os.path.basename(f) for f in failed))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", "-o", LIB,
           *[u[0] for u in units], *LINK_LIBS]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libvarpro_b200.so")
    open(os.path.join(HERE, "libvarpro_b200.stamp"), "w").write(_stamp())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
