"""Generates tests/golden/ref3d_vectors.json by EXECUTING the reference's own hot-path sources on the CPU
(oracle/_ref/libref3dmpifft.so + libtemplatefft_cpu.so, built by oracle/ref_3dmpifft/Makefile from
3dmpifft_opt/include/fft_mpi_3d_api.cpp, kernel_func.cpp, fast_transpose/kernels_{201,120}.cpp and the FFT engine
templateFFT/src/templateFFT.cpp where they lie under /root/reference; the kernels that engine generates at run time are compiled
with g++ and run on fibers).  Run in the build container (needs /root/reference):  python tests/golden/make_ref3d_vectors.py

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
    assert ref.engine == "templatefft", "the vectors must come from the reference's own FFT kernels"
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
    # the FFT engine alone: lines of the baseline's axis lengths (one line each, kept short) and a small 2-D plane
    engine_cases = []
    for n in (8, 9, 12, 49, 64, 100, 125, 243, 512, 768, 1024):
        vals, state = minstd_uniform(2 * n, state)
        x = vals[0::2] + 1j * vals[1::2]
        case = {"fftdim": 1, "shape": [n], "input": pairs(x), "forward": pairs(ref.engine_fft(x))}
        if n <= 125:
            case["backward"] = pairs(ref.engine_fft(x, inverse=True))
        engine_cases.append(case)
    vals, state = minstd_uniform(2 * 6 * 10, state)
    x = (vals[0::2] + 1j * vals[1::2]).reshape(6, 10)
    engine_cases.append({"fftdim": 2, "shape": [6, 10], "input": pairs(x.reshape(-1)), "forward": pairs(ref.engine_fft(x, 2).reshape(-1)),
                         "backward": pairs(ref.engine_fft(x, 2, inverse=True).reshape(-1))})
    policy = [{"n0": n0, "wanted": w, "proper": ref.proper_device_num(n0, w)} for n0 in (512, 10, 9, 5, 7, 100, 33) for w in (1, 2, 3, 4, 8)]
    doc = {"generator": "tests/golden/make_ref3d_vectors.py",
           "library": "3dmpifft_opt/include/{fft_mpi_3d_api.cpp,kernel_func.cpp,fast_transpose/kernels_201.cpp,kernels_120.cpp} of /root/reference, "
                      "and templateFFT/src/templateFFT.cpp (its generator + the kernels it generated), compiled in place and executed on the CPU (oracle/ref_3dmpifft)",
           "engine": ref.engine,
           "layout": "per-device buffers of getMaxDataCount elements as [re, im] pairs; stages in execution order (forward: fftZY, "
                     "localTransposeUneven, slabAlltoall, fftX; backward: fftX, slabAlltoall, localTransposeUneven, fftZY); "
                     "tables[p][q] = [scount, soffset, rcount, roffset]",
           "input": "std::minstd_rand(4242)-style U(0,1) real and imaginary parts (oracle.minstd_uniform), world order, one stream across the cases",
           "cases": cases, "table_cases": table_cases, "device_policy": policy, "engine_cases": engine_cases}
    # one oracle-sized 3-D case for the GPU test (binary, the JSON would be 0.5 MB): 8 x 16 x 32 on one device, input, forward
    # output ([y][z][x]) and the unnormalised backward transform of that output, all from the reference's executed code
    n0, n1, n2 = 8, 16, 32
    g = SlabGeometry(n0, n1, n2, 1)
    vals, state = minstd_uniform(2 * n0 * n1 * n2, state)
    x = vals[0::2] + 1j * vals[1::2]
    fwd, _, _ = ref.execute(g, [x], FORWARD)
    bwd, _, _ = ref.execute(g, [fwd[0]], BACKWARD)
    assert ref.engine_used(n0, n1, n2) == "templatefft"
    np.savez(os.path.join(ROOT, "tests", "golden", "ref3d_case_8x16x32.npz"), shape=np.array([n0, n1, n2]), input=x, forward=fwd[0], backward=bwd[0])
    out = os.path.join(ROOT, "tests", "golden", "ref3d_vectors.json")
    with open(out, "w") as f:
        json.dump(doc, f)
    print("wrote", out, os.path.getsize(out), "bytes;", len(cases), "cases")


if __name__ == "__main__":
    main()
