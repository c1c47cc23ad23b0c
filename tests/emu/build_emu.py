"""TEST-ONLY: compile dav1d_b200/csrc/*.cu with g++ against tests/emu/cuda_emu.h into
tests/emu/_build/libb200av1_emu.so (same C ABI, kernels run on CPU fibers). See cuda_emu.h.
Never loaded by the dav1d_b200 package."""
import os, subprocess, hashlib, json
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "dav1d_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
OUT = os.path.join(BUILD, "libb200av1_emu.so")
FLAGS = ["-std=c++17", "-O1", "-g", "-fPIC", "-DB200_EMU", "-include", os.path.join(HERE, "cuda_emu.h"),
         "-I" + HERE, "-w"]


def _stamp():
    h = hashlib.sha256()
    for root in (CSRC, HERE, os.path.join(ROOT, "include")):
        for f in sorted(os.listdir(root)):
            p = os.path.join(root, f)
            if os.path.isfile(p) and not f.endswith(".pyc"):
                h.update(f.encode()); h.update(open(p, "rb").read())
    return h.hexdigest()


def build(force=False):
    os.makedirs(BUILD, exist_ok=True)
    sf = os.path.join(BUILD, "stamp.json")
    stamp = _stamp()
    if not force and os.path.exists(OUT) and os.path.exists(sf) and json.load(open(sf)).get("stamp") == stamp:
        return OUT
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))

    def cc(src):
        o = os.path.join(BUILD, src[:-3] + ".o")
        r = subprocess.run(["g++"] + FLAGS + ["-x", "c++", "-c", os.path.join(CSRC, src), "-o", o],
                           capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError("g++ (emu) failed on %s:\n%s" % (src, r.stderr[-4000:]))
        return o

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(cc, srcs))
    o = os.path.join(BUILD, "cuda_emu.o")
    r = subprocess.run(["g++"] + FLAGS + ["-c", os.path.join(HERE, "cuda_emu.cpp"), "-o", o], capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(r.stderr[-4000:])
    r = subprocess.run(["g++", "-shared", "-o", OUT] + objs + [o], capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(r.stderr[-4000:])
    json.dump({"stamp": stamp}, open(sf, "w"))
    return OUT


if __name__ == "__main__":
    print(build())
