"""Build dav1d_b200/libb200av1.so: every .cu under csrc/ compiled by nvcc for sm_100a and
linked into ONE in-tree shared library (it travels to the GPU box with the repo snapshot).

    python -m dav1d_b200.build [-v] [--force]
"""
import os, subprocess, sys, hashlib, json
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libb200av1.so")
OBJ = os.path.join(HERE, "_obj")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
         "-Xcompiler", "-fPIC,-fvisibility=hidden", "--use_fast_math", "-Xptxas", "-v"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if not os.path.isfile(os.path.join(root, f)):
                continue
            with open(os.path.join(root, f), "rb") as fh:
                h.update(f.encode()); h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build_variant(name, defines):
    """Tuning aid: a second library `libb200av1_<name>.so` compiled with extra -D flags (selected at run time with
    B200AV1_LIB=<path>). Never used by the tests or the driver."""
    out = os.path.join(HERE, "libb200av1_%s.so" % name)
    obj = os.path.join(OBJ, "var_" + name)
    os.makedirs(obj, exist_ok=True)
    objs = []
    for src in _sources():
        o = os.path.join(obj, src[:-3] + ".o")
        r = subprocess.run([NVCC] + FLAGS + ["-D%s" % d for d in defines] + ["-c", os.path.join(CSRC, src), "-o", o], capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, r.stderr[-4000:]))
        objs.append(o)
    r = subprocess.run([NVCC, "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"], capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    return out


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    stamp_file = os.path.join(OBJ, "stamp.json")
    stamp = _stamp()
    if not force and os.path.exists(OUT) and os.path.exists(stamp_file):
        try:
            if json.load(open(stamp_file)).get("stamp") == stamp:
                return OUT
        except Exception:
            pass
    if not os.path.exists(NVCC):
        raise RuntimeError("nvcc not found at %s - cannot build libb200av1.so" % NVCC)
    srcs = _sources()

    def cc(src):
        o = os.path.join(OBJ, src[:-3] + ".o")
        cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(o + ".log", "w") as fh:
            fh.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, r.stderr[-4000:]))
        if verbose:
            print("[b200 build] %s ok" % src, file=sys.stderr)
        return o

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(cc, srcs))
    cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    json.dump({"stamp": stamp}, open(stamp_file, "w"))
    return OUT


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv, force="--force" in sys.argv)
    print(p)
