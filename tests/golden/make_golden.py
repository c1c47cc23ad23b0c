#!/usr/bin/env python3
"""Generate the committed golden fixtures from the UNMODIFIED reference C path.

Run in the build container (needs oracle/_ref/libdav1d_ref.so, i.e. /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/scans.npz (dav1d_scans, used by the checkasm-style generator when the
reference build is absent) and tests/golden/itx_golden.npz (one checkasm-style case per
defined itxfm_add[tx][txtp] slot and bit depth: inputs + the reference's outputs).
"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import refs
from dav1d_b200 import levels as L


def main():
    assert refs.have_ref()
    np.savez_compressed(os.path.join(HERE, "scans.npz"), **{"tx%d" % tx: refs.scan_table(tx).astype(np.uint16) for tx in range(19)})
    rng = np.random.default_rng(20260922)
    out = {}
    for bpc in (8, 10, 12):
        tbl = refs.ref_itx_table(bpc)
        bdmax = (1 << bpc) - 1
        for tx in range(19):
            w, h = L.TX_W[tx], L.TX_H[tx]
            smax = refs.SUBSH_ITERS[int(np.log2(max(w, h))) - 2]
            for tp in range(17):
                if not L.itx_defined(tx, tp):
                    continue
                subsh = int(rng.integers(1 if tp else 0, smax))
                coef, eob = refs.gen_itx_coefs(rng, tx, tp, subsh, bdmax)
                coef = coef.astype(refs.coef_dtype(bpc))
                dst = rng.integers(0, bdmax + 1, (h, w)).astype(refs.pixel_dtype(bpc))
                k = "b%d_t%d_p%d_" % (bpc, tx, tp)
                out[k + "coef"] = coef.copy(); out[k + "eob"] = np.int32(eob); out[k + "dst"] = dst.copy()
                d = dst.copy(); c = coef.copy()
                tbl[tx][tp](d, d.strides[0], c, eob)
                out[k + "out"] = d; out[k + "coef_out"] = c
    np.savez_compressed(os.path.join(HERE, "itx_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
