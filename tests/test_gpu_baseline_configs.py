"""BASELINE.json configs at FULL size against the CPU oracle (`-m gpu`; a case is skipped when the box has fewer GPUs):
  C2 512^3 double, 1 GPU          C3 512^3 double, 4 GPUs
  C4 1024^3 double, 8 GPUs        C5 768^3 single, 8 GPUs (mixed radix)
For every config the WHOLE forward spectrum of every device (its transposed y-slab [y_l][z][x]) is compared with the
oracle's slab pipeline on the same heFFTe-test input (std::minstd_rand(4242) -> U(0,1), test_fft3d.h:19-27), like
heFFTe's own test compares the whole world box (test/test_fft3d.h:99-115); then the backward transform of that
spectrum must return N^3 * input within the reference tolerance (1e-11 double / 5e-4 single, test_common.h:136-140).
The default plan flags are used, i.e. the production path of each device count (P2P exchange, stream-pipelined
forward for P > 1)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import distributedfft_b200 as dfft  # noqa: E402
from oracle import BACKWARD, FORWARD, COracle, NumpySlab, SlabGeometry  # noqa: E402
from gpu_helpers import run_slab  # noqa: E402

CONFIGS = [("C2", 512, 1, dfft.DOUBLE, 0), ("C3", 512, 4, dfft.DOUBLE, 0), ("C3b", 512, 8, dfft.DOUBLE, 0), ("C3c", 512, 2, dfft.DOUBLE, 0),
           # the pipelined schedules at full size, also where they are not the default (kernel chain for the cube)
           ("C3-chain", 512, 4, dfft.DOUBLE, dfft.FORCE_PIPELINE), ("C3c-chain", 512, 2, dfft.DOUBLE, dfft.FORCE_PIPELINE),
           ("C4", 1024, 8, dfft.DOUBLE, 0), ("C5", 768, 8, dfft.FLOAT, 0)]


@pytest.mark.parametrize("name,n,P,precision,flags", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_full_size_forward_spectrum_and_round_trip_vs_oracle(name, n, P, precision, flags):
    if torch.cuda.device_count() < P:
        pytest.skip(f"needs {P} GPUs, box has {torch.cuda.device_count()}")
    co = COracle()
    g = SlabGeometry(n, n, n, P)
    ns = NumpySlab(n, n, n, P)
    a = np.empty(n ** 3, dtype=np.complex128)
    co.fill_minstd(a, 4242)
    A = a.reshape(n, n, n)
    inputs = ns.scatter_input(A)
    b1 = [b.copy() for b in inputs]
    b2 = [np.zeros_like(b) for b in b1]
    co.slab_execute(g, b1, b2, FORWARD)          # oracle: reference stages t0..t3 on P "devices"
    del b1
    cast = (lambda x: x) if precision == dfft.DOUBLE else (lambda x: x.astype(np.complex64))
    res = run_slab(n, n, n, P, FORWARD, [cast(b) for b in inputs], precision=precision, repeat=2, refill=False, flags=flags)
    scale = max(np.abs(r[: g.out_count(q)]).max() for q, r in enumerate(b2))
    tol = 1e-12 * np.log2(float(n) ** 3) if precision == dfft.DOUBLE else 5e-6
    for q in range(P):
        cnt = g.out_count(q)
        err = np.abs(res[q]["buf2"][:cnt] - b2[q][:cnt]).max() / scale
        assert err <= tol, (name, "forward", q, err)
    if P > 1:
        assert res[0]["exchange"] == dfft.EXCHANGE_P2P
    spectra = [r["buf2"] for r in res]
    del res, b2
    back = run_slab(n, n, n, P, BACKWARD, spectra, precision=precision, flags=flags)
    rt_tol = 1e-11 if precision == dfft.DOUBLE else 5e-4
    for p in range(P):
        cnt = g.in_count(p)
        err = np.abs(back[p]["buf2"][:cnt] / float(n) ** 3 - inputs[p][:cnt]).max()
        assert err <= rt_tol, (name, "round trip", p, err)
