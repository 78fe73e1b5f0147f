"""Generates tests/golden/ref3d_vectors.json by EXECUTING the reference's own hot-path sources on the CPU
(oracle/_ref/libref3dmpifft.so, built by oracle/ref_3dmpifft/Makefile from 3dmpifft_opt/include/fft_mpi_3d_api.cpp,
kernel_func.cpp and fast_transpose/kernels_{201,120}.cpp where they lie under /root/reference; only the JIT FFT engine is
replaced by a DFT).  Run in the build container (needs /root/reference):  python tests/golden/make_ref3d_vectors.py

Per case: the per-device inputs, the per-device outputs, BOTH plan buffers (bufferDev1, bufferDev2) after each of the four
stages in execution order, and the TransInfo tables the reference's plan creation filled.  Plus tables-only cases for
geometries too large to push through a DFT."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import BACKWARD, FORWARD, NumpySlab, Ref3dmpifft, SlabGeometry, minstd_uniform  # noqa: E402


def pairs(a):
    return [[float(v.real), float(v.imag)] for v in a]


def main():
    ref = Ref3dmpifft()
    cases, state = [], 4242
    for (n0, n1, n2, P) in [(4, 4, 8, 1), (6, 4, 4, 2), (5, 7, 4, 3)]:
        g = SlabGeometry(n0, n1, n2, P)
        vals, state = minstd_uniform(2 * n0 * n1 * n2, state)
        A = (vals[0::2] + 1j * vals[1::2]).reshape(n0, n1, n2)
        for direction in (FORWARD, BACKWARD):
            if direction == FORWARD:
                ins = NumpySlab(n0, n1, n2, P).scatter_input(A)          # x-slabs [x_l][y][z]
            else:
                ins = []
                for q in range(P):                                       # y-slabs [y_l][z][x]
                    b = np.zeros(g.max_count(q), dtype=np.complex128)
                    blk = A[:, q * g.yd: q * g.yd + g.n1l(q), :].transpose(1, 2, 0).reshape(-1)
                    b[: blk.size] = blk
                    ins.append(b)
            outs, tables, dumps = ref.execute(g, ins, direction, stages=True)
            cases.append({"shape": [n0, n1, n2], "devices": P, "direction": direction,
                          "inputs": [pairs(b) for b in ins], "outputs": [pairs(b) for b in outs],
                          "stages": [[{"buffer1": pairs(dumps[p][s][0]), "buffer2": pairs(dumps[p][s][1])} for s in range(4)] for p in range(P)],
                          "tables": tables.tolist()})
    table_cases = []
    for (n0, n1, n2, P) in [(64, 64, 64, 8), (100, 60, 16, 7), (30, 22, 24, 8), (96, 80, 32, 5), (33, 17, 8, 4)]:
        for direction in (FORWARD, BACKWARD):
            table_cases.append({"shape": [n0, n1, n2], "devices": P, "direction": direction, "tables": ref.tables(n0, n1, n2, P, direction).tolist(),
                                "max_data_count": [ref.max_data_count(n0, n1, n2, P, False), ref.max_data_count(n0, n1, n2, P, True)]})
    policy = [{"n0": n0, "wanted": w, "proper": ref.proper_device_num(n0, w)} for n0 in (512, 10, 9, 5, 7, 100, 33) for w in (1, 2, 3, 4, 8)]
    doc = {"generator": "tests/golden/make_ref3d_vectors.py",
           "library": "3dmpifft_opt/include/{fft_mpi_3d_api.cpp,kernel_func.cpp,fast_transpose/kernels_201.cpp,kernels_120.cpp} of /root/reference, "
                      "compiled in place and executed on the CPU (oracle/ref_3dmpifft)",
           "layout": "per-device buffers of getMaxDataCount elements as [re, im] pairs; stages in execution order (forward: fftZY, "
                     "localTransposeUneven, slabAlltoall, fftX; backward: fftX, slabAlltoall, localTransposeUneven, fftZY); "
                     "tables[p][q] = [scount, soffset, rcount, roffset]",
           "input": "std::minstd_rand(4242)-style U(0,1) real and imaginary parts (oracle.minstd_uniform), world order, one stream across the cases",
           "cases": cases, "table_cases": table_cases, "device_policy": policy}
    out = os.path.join(ROOT, "tests", "golden", "ref3d_vectors.json")
    with open(out, "w") as f:
        json.dump(doc, f)
    print("wrote", out, os.path.getsize(out), "bytes;", len(cases), "cases")


if __name__ == "__main__":
    main()
