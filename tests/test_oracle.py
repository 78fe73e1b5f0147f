"""CPU tests: pin the oracle against the reference tree's golden vectors, against numpy's
pocketfft, and the two independent restatements (C and numpy) against each other."""
import json
import os

import numpy as np
import pytest

from oracle import BACKWARD, FORWARD, COracle, NumpySlab, SlabGeometry, minstd_uniform, proper_device_num

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def co():
    return COracle()


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(HERE, "golden", "heffte_vectors.json")) as f:
        return json.load(f)


def _c(v):
    a = np.asarray(v, dtype=np.float64)
    return a[..., 0] + 1j * a[..., 1]


def test_golden_box_per_axis(co, golden):
    """heFFTe pen-and-paper 2x3x4 box (test_units_nompi.cpp:92-190), each axis, fwd and back."""
    shape = tuple(golden["box_shape_c_order"])
    x = _c(golden["box_input"]).reshape(shape)
    for axis, key in ((2, "box_fft_dim0_axis2"), (1, "box_fft_dim1_axis1"), (0, "box_fft_dim2_axis0")):
        ref = _c(golden[key]).reshape(shape)
        got = co.fft_axis(x, axis, -1)
        assert np.abs(got - ref).max() < 1e-11
        back = co.fft_axis(got, axis, +1) / shape[axis]
        assert np.abs(back - x).max() < 1e-11
        # numpy twin agrees with the pen-and-paper values too
        assert np.abs(np.fft.fft(x, axis=axis) - ref).max() < 1e-11


def test_golden_dft11(co, golden):
    """11-point DFT of 1..11 (test_units_stock.cpp:229-255)."""
    x = _c(golden["dft11_input"])
    ref = _c(golden["dft11_output"])
    assert np.abs(co.fft_axis(x, 0, -1) - ref).max() < 1e-11


def test_radix_schedule(co):
    """templateFFT.cpp:3956-3963 + 4540-4550 + 4580-4588 (pure power-of-two rows of SURVEY App. B)."""
    assert co.radix_schedule(64) == [8, 8]
    assert co.radix_schedule(512) == [8, 8, 8]
    assert co.radix_schedule(1024) == [8, 8, 8, 2]
    assert np.prod(co.radix_schedule(768)) == 768
    assert co.radix_schedule(17) == []  # FFT_ERROR_UNSUPPORTED_RADIX in the reference


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 11, 13, 16, 35, 64, 243, 512, 625, 768, 1024, 17])
def test_engine_vs_pocketfft(co, n):
    rng = np.random.default_rng(n)
    a = rng.standard_normal((3, n, 4)) + 1j * rng.standard_normal((3, n, 4))
    for sign, ref in ((-1, np.fft.fft(a, axis=1)), (+1, np.fft.ifft(a, axis=1) * n)):
        got = co.fft_axis(a, 1, sign)
        assert np.abs(got - ref).max() <= 1e-12 * max(1.0, np.log2(max(n, 2))) * np.abs(ref).max()


def test_device_count_policy(co):
    """getProperDeviceNum / getMaxDataCount, fft_mpi_3d_api.cpp:232-316."""
    assert proper_device_num(512, 8) == 8
    assert proper_device_num(10, 4) == 4   # per=3 -> 3,3,3,1
    assert proper_device_num(9, 4) == 3    # per=3 -> 3 devices
    assert proper_device_num(5, 4) == 3    # per=2 -> 2,2,1
    for n0, w in ((512, 8), (10, 4), (9, 4), (5, 4), (7, 3)):
        assert co.lib.oracle_proper_device_num(n0, w) == proper_device_num(n0, w)
    g = SlabGeometry(10, 9, 4, 3)
    for p in range(3):
        assert co.lib.oracle_max_data_count(10, 9, 4, 3, int(p == 2)) == g.max_count(p)


def test_exchange_table(co):
    """fft_mpi_3d_api.cpp:84-133: counts are consistent between sender and receiver."""
    for (n0, n1, n2, P) in [(8, 8, 8, 2), (10, 9, 4, 3), (9, 10, 4, 3), (64, 64, 64, 8)]:
        for direction in (FORWARD, BACKWARD):
            tabs = [co.exchange_table(n0, n1, n2, P, d, direction) for d in range(P)]
            for s in range(P):
                for r in range(P):
                    assert tabs[s]["scount"][r] == tabs[r]["rcount"][s]
            g = SlabGeometry(n0, n1, n2, P)
            for d in range(P):
                tot = g.in_count(d) if direction == FORWARD else g.out_count(d)
                assert tabs[d]["scount"].sum() == tot


CASES = [(8, 8, 8, 1), (8, 8, 8, 2), (16, 8, 4, 4), (10, 9, 4, 3), (9, 10, 4, 3), (12, 7, 10, 4), (64, 64, 64, 8)]


@pytest.mark.parametrize("n0,n1,n2,P", CASES)
def test_slab_pipeline_stagewise(co, n0, n1, n2, P):
    """C oracle and numpy restatement agree after every stage (t0..t3, both directions) and the
    whole transform equals fftn; includes uneven splits."""
    rng = np.random.default_rng(n0 * 1000 + n1 * 10 + P)
    A = rng.standard_normal((n0, n1, n2)) + 1j * rng.standard_normal((n0, n1, n2))
    g = SlabGeometry(n0, n1, n2, P)
    ns = NumpySlab(n0, n1, n2, P)
    fwd_out = None
    for stop in range(4):
        a1 = ns.scatter_input(A); a2 = [np.zeros_like(b) for b in a1]
        b1 = ns.scatter_input(A); b2 = [np.zeros_like(b) for b in b1]
        ns.execute(a1, a2, FORWARD, stop)
        co.slab_execute(g, b1, b2, FORWARD, stop)
        scale = np.abs(np.fft.fftn(A)).max()
        for p in range(P):
            assert np.abs(a1[p] - b1[p]).max() <= 1e-12 * scale
            assert np.abs(a2[p] - b2[p]).max() <= 1e-12 * scale
        if stop == 3:
            fwd_out = b2
            S = ns.gather_forward_output(b2)
            assert np.abs(S - np.fft.fftn(A)).max() <= 1e-12 * np.log2(n0 * n1 * n2) * scale
    for stop in range(4):
        a1 = [b.copy() for b in fwd_out]; a2 = [np.zeros_like(b) for b in a1]
        b1 = [b.copy() for b in fwd_out]; b2 = [np.zeros_like(b) for b in b1]
        ns.execute(a1, a2, BACKWARD, stop)
        co.slab_execute(g, b1, b2, BACKWARD, stop)
        scale = np.abs(A).max() * n0 * n1 * n2
        for p in range(P):
            assert np.abs(a1[p] - b1[p]).max() <= 1e-12 * scale
            assert np.abs(a2[p] - b2[p]).max() <= 1e-12 * scale
        if stop == 3:
            B = ns.gather_natural(b2) / (n0 * n1 * n2)
            assert np.abs(B - A).max() <= 1e-11


def test_config_c1_roundtrip_64cube(co):
    """BASELINE config 1: 64^3 forward+inverse, 1 rank, CPU, round-trip max error <= 1e-11 on both
    the driver's ramp input (fftSpeed3d_c2c.cpp:61-63, 84-91) and heFFTe's U(0,1) input."""
    n = 64
    g = SlabGeometry(n, n, n, 1)
    for kind in ("ramp", "minstd"):
        a = np.zeros(n ** 3, dtype=np.complex128)
        if kind == "ramp":
            co.fill_ramp(a, 0)
        else:
            co.fill_minstd(a, 4242)
        b1 = [a.copy()]; b2 = [np.zeros_like(a)]
        co.slab_execute(g, b1, b2, FORWARD)
        c1 = [b2[0].copy()]; c2 = [np.zeros_like(a)]
        co.slab_execute(g, c1, c2, BACKWARD)
        drv, absolute = co.roundtrip_error(a, c2[0], float(n) ** 3)
        if kind == "ramp":
            assert drv <= 1e-11
        else:
            assert absolute <= 1e-11


def test_minstd_generator(co):
    """C and numpy restatements of the heFFTe input generator agree; values in [0,1)."""
    a = np.zeros(1000, dtype=np.complex128)
    st = co.fill_minstd(a, 4242)
    v, st2 = minstd_uniform(1000, 4242)
    assert st == st2
    assert np.array_equal(a.real, v) and np.all(a.imag == 0)
    assert v.min() >= 0.0 and v.max() < 1.0
    # first draw of std::minstd_rand(4242): 4242*48271 mod (2^31-1)
    assert (4242 * 48271) % 2147483647 == 204765582
