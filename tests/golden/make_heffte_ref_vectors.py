#!/usr/bin/env python
"""Generates tests/golden/heffte_ref_vectors.json by RUNNING the reference tree's heFFTe 2.1.0 (stock backend, built by
oracle/ref_heffte/Makefile from /root/reference/heffte/heffteBenchmark) in this container: forward spectra of small
world arrays over P slab ranks.  /root/reference does not exist on the GPU box; these committed vectors (and the
prebuilt oracle/_ref/libheffte_ref.so, which does travel) are what the tests there compare against.

    python tests/golden/make_heffte_ref_vectors.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import FORWARD, HeffteRef, minstd_uniform  # noqa: E402

CASES = [(2, 3, 4, 1, "alltoallv"), (4, 6, 8, 2, "p2p_plined"), (10, 9, 4, 3, "p2p"), (9, 10, 4, 3, "alltoall"), (16, 8, 8, 4, "p2p_plined"),
         (8, 8, 8, 8, "alltoallv"), (12, 10, 24, 2, "p2p_plined")]


def main():
    ref = HeffteRef()
    out = {"generator": "tests/golden/make_heffte_ref_vectors.py", "library": "heFFTe %d (stock backend) from /root/reference/heffte/heffteBenchmark" % ref.version(),
           "input": "heFFTe test input: std::minstd_rand(4242) -> U(0,1), real part only, world order (test/test_fft3d.h:19-27); case 0: 1..24 (test_units_nompi.cpp:92-98)",
           "layout": "world arrays A[x][y][z], z fastest; spectrum in the same natural order; values as [re, im] pairs", "cases": []}
    for n0, n1, n2, P, alg in CASES:
        cnt = n0 * n1 * n2
        if (n0, n1, n2) == (2, 3, 4):
            vals = np.arange(1, 25, dtype=np.float64)
        else:
            vals, _ = minstd_uniform(cnt, 4242)
        A = vals.astype(np.complex128).reshape(n0, n1, n2)
        S = ref.fft3d(A, P, FORWARD, alg)
        assert np.abs(S - np.fft.fftn(A)).max() <= 1e-12 * np.abs(S).max()   # sanity only; the vectors are the library's output
        out["cases"].append({"shape": [n0, n1, n2], "ranks": P, "algorithm": alg, "input_real": vals.tolist(),
                             "spectrum": np.stack([S.real.reshape(-1), S.imag.reshape(-1)], axis=1).tolist()})
    with open(os.path.join(ROOT, "tests", "golden", "heffte_ref_vectors.json"), "w") as f:
        json.dump(out, f)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
