#!/usr/bin/env python
"""Writes tests/golden/heffte_vectors.json.

These are the only hard numbers the reference tree pins for an FFT at this boundary
(SURVEY.md section 8c).  They are restated here from the formulas / literals in
  heffte/heffteBenchmark/test/test_units_nompi.cpp:92-98   (input 1..24 on a 2x3x4 box, dim 0 fastest)
  heffte/heffteBenchmark/test/test_units_nompi.cpp:136-190 (make_fft0 / make_fft1 / make_fft2)
  heffte/heffteBenchmark/test/test_units_stock.cpp:229-255 (11-point DFT of 1..11)
  heffte/heffteBenchmark/test/test_common.h:136-140        (tolerances 1e-11 double / 5e-4 float)
No reference code is executed (it needs MPI); the script only re-types the pen-and-paper values.
Run:  python tests/golden/make_golden.py
"""
import json
import os


def cj(re, im=0.0):
    return [float(re), float(im)]


def main():
    inp = [cj(i + 1) for i in range(24)]
    # make_fft0: transforms over dim 0 (size 2, fastest)
    fft0 = [None] * 24
    for i in range(0, 24, 2):
        fft0[i] = cj(3 + 2 * i)
        fft0[i + 1] = cj(-1.0)
    # make_fft1: transforms over dim 1 (size 3, stride 2)
    fft1 = [None] * 24
    for j in range(4):
        for i in range(2):
            fft1[6 * j + i] = cj((2 * j + i + 1) * 9.0 - i * 6.0)
            fft1[6 * j + i + 2] = cj(-3.0, 1.73205080756888)
            fft1[6 * j + i + 4] = cj(-3.0, -1.73205080756888)
    # make_fft2: transforms over dim 2 (size 4, stride 6)
    fft2 = [None] * 24
    for i in range(6):
        fft2[i] = cj(40.0 + 4 * i)
        fft2[i + 6] = cj(-12.0, 12.0)
        fft2[i + 12] = cj(-12.0)
        fft2[i + 18] = cj(-12.0, -12.0)
    imag = [18.731279813890875, 8.55816705136493, 4.765777128986846, 2.5117658384695547, 0.790780616972353]
    dft11 = [None] * 11
    dft11[0] = cj(66, 0)
    for i in range(1, 6):
        dft11[i] = cj(-5.5, imag[i - 1])
        dft11[11 - i] = cj(-5.5, -imag[i - 1])
    out = {
        "source": "heffte/heffteBenchmark/test/test_units_nompi.cpp:92-190, test_units_stock.cpp:229-255",
        "box_shape_c_order": [4, 3, 2],
        "box_input": inp,
        "box_fft_dim0_axis2": fft0,
        "box_fft_dim1_axis1": fft1,
        "box_fft_dim2_axis0": fft2,
        "dft11_input": [cj(i + 1) for i in range(11)],
        "dft11_output": dft11,
        "tolerance_double": 1e-11,
        "tolerance_float": 5e-4,
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "heffte_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
