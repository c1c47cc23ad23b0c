"""bench.py on the CPU: what can be checked without a GPU (the reference arm runs here; the byte model and the stage tables
must cover every workload)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import refs
sys.path.insert(0, refs.ROOT)
import bench          # noqa: E402
from dav1d_b200 import synth


def test_byte_model_covers_every_stage():
    """SURVEY 8(d) accounting: every stage the bench times has algorithmic bytes, for pure inter, mixed and 10-bit grain frames"""
    S = synth.make_inter_frame(np.random.default_rng(5), 8, 264, 200, p_intra=0.1, p_obmc=0.2, p_warp=0.15, p_ii=0.15)
    alg = bench.frame_algorithmic_bytes(S)
    assert set(bench.STAGE_BYTES_KEY.values()) <= set(alg)
    for k in ("mc", "warp", "blend", "comp", "itx", "intra", "deblock", "cdef", "lr"):
        assert alg[k] > 0, k
    assert alg["fg"] == 0
    S = synth.make_inter_frame(np.random.default_rng(6), 10, 200, 136, film_grain=True)
    alg = bench.frame_algorithmic_bytes(S)
    assert alg["fg"] > 0 and alg["warp"] == 0 and alg["intra"] == 0
    # a stage list in stage_times() that the tables do not know would be a KeyError on the GPU box only
    src = open(os.path.join(refs.ROOT, "bench.py")).read()
    import re
    names = set(re.findall(r'stages\.append\(\("([a-z]+)"', src))
    assert names and names <= set(bench.STAGE_NAMES), names


def test_frame_traffic_file_is_per_workload():
    tj = json.load(open(os.path.join(refs.ROOT, "profiles", "frame_traffic.json")))
    for wl in ("4k8_inter", "4k10_full", "4k8_mixed"):
        assert isinstance(tj[wl], dict) and all(v > 0 for v in tj[wl].values())
        assert set(tj[wl]) <= set(bench.STAGE_NAMES) | {"expand"}


def test_host_threads_respects_affinity():
    n, info = bench.host_threads()
    assert 1 <= n <= len(os.sched_getaffinity(0)) and "affinity" in json.dumps(info)


@pytest.mark.skipif(not refs.have_ref(), reason="oracle/_ref not built")
def test_reference_arm_line():
    """`bench.py --impl reference` (runs on the host cores): one JSON line with the contract's keys, the same workload string
    as our arm would print, --steps honoured"""
    r = subprocess.run([sys.executable, os.path.join(refs.ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--workload", "4k8_mixed"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e", "impl"):
        assert k in d, k
    assert d["impl"] == "reference" and d["steps"] == 1 and d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference"
    assert d["config"]["workload"].startswith("4k8_mixed: ") and d["e2e"]["h2d_bytes_per_step"] == 0
