"""Pins the oracle -- stage by stage -- on the reference's OWN HOT-PATH CODE, executed.

oracle/ref_3dmpifft compiles 3dmpifft_opt/include/fft_mpi_3d_api.cpp (plan creation, TransInfo tables, fftZY,
localTransposeUneven, slabAlltoall, fftX, fft_mpi_execute_dft_3d_c2c), kernel_func.cpp (the pack / unpack kernels) and
fast_transpose/kernels_{201,120}.cpp (the cuTranspose tile kernels) from /root/reference, in place, against a HIP-on-CPU shim
and runs them on host memory (GPU threads are fibers; only the JIT FFT engine is replaced, by a DFT).

* committed vectors (tests/golden/ref3d_vectors.json, made by tests/golden/make_ref3d_vectors.py in the build container):
  BOTH plan buffers of every device after EVERY stage, the outputs and the exchange tables -- compared with both
  restatements (oracle_fft.c and the numpy twin) and, for the tables / counts / device policy, with the product library's
  host logic (libdfft.so; no GPU needed for those entry points);
* live, whenever oracle/_ref/libref3dmpifft.so is present (this container; shipped prebuilt to the GPU box): random inputs
  over even and uneven splits up to 8 devices, both directions, against numpy's fftn, against the oracle stage by stage, and
  against the product's recorded multi-device schedule interpreted on the CPU (tests/test_dry_run.py)."""
import json
import os

import numpy as np
import pytest

import distributedfft_b200 as dfft
from oracle import BACKWARD, FORWARD, COracle, NumpySlab, SlabGeometry, build_ref3d, proper_device_num

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 2e-13          # absolute, on O(1) inputs of <= 10^4 points: DFT vs Stockham vs pocketfft rounding


@pytest.fixture(scope="module")
def co():
    return COracle()


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(HERE, "golden", "ref3d_vectors.json")) as f:
        return json.load(f)


def c(v):
    a = np.asarray(v, dtype=np.float64)
    return a[:, 0] + 1j * a[:, 1]


def test_both_restatements_match_the_executed_reference_at_every_stage_boundary(co, gold):
    assert "fft_mpi_3d_api.cpp" in gold["library"] and "kernel_func.cpp" in gold["library"]
    for case in gold["cases"]:
        n0, n1, n2 = case["shape"]
        P, direction = case["devices"], case["direction"]
        g = SlabGeometry(n0, n1, n2, P)
        ns = NumpySlab(n0, n1, n2, P)
        ins = [c(b) for b in case["inputs"]]
        assert [b.size for b in ins] == [g.max_count(p) for p in range(P)]
        for s in range(4):
            for twin in (False, True):
                b1 = [b.copy() for b in ins]
                b2 = [np.zeros_like(b) for b in ins]
                if twin:
                    ns.execute(b1, b2, direction, stop_after=s)
                else:
                    co.slab_execute(g, b1, b2, direction, stop_after=s)
                for p in range(P):
                    st = case["stages"][p][s]
                    assert np.abs(b1[p] - c(st["buffer1"])).max() <= TOL, (case["shape"], P, direction, "stage", s, "bufferDev1", p, twin)
                    if not twin:      # (the numpy twin only keeps the buffer a stage produces; oracle_fft.c mirrors both)
                        assert np.abs(b2[p] - c(st["buffer2"])).max() <= TOL, (case["shape"], P, direction, "stage", s, "bufferDev2", p)
                    else:
                        live = b2[p] != 0
                        assert np.abs(b2[p][live] - c(st["buffer2"])[live]).max(initial=0) <= TOL
        # the result the caller sees
        b1 = [b.copy() for b in ins]
        b2 = [np.zeros_like(b) for b in ins]
        co.slab_execute(g, b1, b2, direction)
        for p in range(P):
            assert np.abs(b2[p] - c(case["outputs"][p])).max() <= TOL
        # ... and it IS the 3-D transform (forward: y-slabs [y_l][z][x]; backward: x-slabs [x_l][y][z], unnormalised)
        if direction == FORWARD:
            A = np.concatenate([ins[p][: g.in_count(p)] for p in range(P)]).reshape(n0, n1, n2)
            F = np.fft.fftn(A)
            for q in range(P):
                ref = F[:, q * g.yd: q * g.yd + g.n1l(q), :].transpose(1, 2, 0).reshape(-1)
                assert np.abs(c(case["outputs"][q])[: ref.size] - ref).max() <= TOL * 10


def test_exchange_tables_counts_and_device_policy_match_the_executed_reference(co, gold):
    """Integers, so exact: the oracle's AND the product library's TransInfo tables (dfft_exchange_table), getMaxDataCount
    (dfft_max_data_count) and the device-count policy (dfft_init) against what the reference's own functions returned."""
    for case in gold["cases"] + gold["table_cases"]:
        n0, n1, n2 = case["shape"]
        P, direction = case["devices"], case["direction"]
        t = np.asarray(case["tables"])
        for p in range(P):
            mine = dfft.exchange_table(n0, n1, n2, P, p, direction)
            orc = co.exchange_table(n0, n1, n2, P, p, direction)
            for j, k in enumerate(("scount", "soffset", "rcount", "roffset")):
                assert list(mine[k]) == list(t[p, :, j]) == list(orc[k]), (case["shape"], P, direction, p, k)
        if "max_data_count" in case:
            for last in (False, True):
                assert dfft.getMaxDataCount(n0, n1, n2, P, last) == case["max_data_count"][int(last)] == co.lib.oracle_max_data_count(n0, n1, n2, P, int(last))
    for row in gold["device_policy"]:
        assert proper_device_num(row["n0"], row["wanted"]) == row["proper"]
        assert dfft.fft_mpi_init([row["n0"], 64, 4], row["wanted"])[0] == row["proper"]


@pytest.fixture(scope="module")
def ref():
    if build_ref3d() is None:
        pytest.skip("oracle/_ref/libref3dmpifft.so not built and /root/reference absent")
    from oracle import Ref3dmpifft
    return Ref3dmpifft()


def _inputs(g, A, direction):
    if direction == FORWARD:
        return NumpySlab(g.n0, g.n1, g.n2, g.P).scatter_input(A)
    ins = []
    for q in range(g.P):
        b = np.zeros(g.max_count(q), dtype=np.complex128)
        blk = A[:, q * g.yd: q * g.yd + g.n1l(q), :].transpose(1, 2, 0).reshape(-1)
        b[: blk.size] = blk
        ins.append(b)
    return ins


LIVE = [(1, 8, 8, 8), (2, 8, 8, 8), (3, 10, 9, 4), (3, 9, 10, 4), (4, 12, 10, 8), (4, 16, 8, 8), (8, 16, 16, 16), (3, 15, 22, 26), (5, 14, 9, 6), (8, 24, 16, 16), (7, 20, 27, 4)]


@pytest.mark.parametrize("P,n0,n1,n2", LIVE)
def test_live_executed_reference_vs_numpy_oracle_and_the_products_schedule(co, ref, P, n0, n1, n2):
    from test_dry_run import simulate
    g = SlabGeometry(n0, n1, n2, P)
    rng = np.random.default_rng(n0 * 1000 + n1 * 10 + P)
    A = rng.standard_normal((n0, n1, n2)) + 1j * rng.standard_normal((n0, n1, n2))
    for direction in (FORWARD, BACKWARD):
        ins = _inputs(g, A, direction)
        outs, tables, dumps = ref.execute(g, ins, direction, stages=True)
        whole, _, _ = ref.execute(g, ins, direction)          # the reference's own fft_mpi_execute_dft_3d_c2c
        scale = np.abs(np.fft.fftn(A)).max()
        # (1) numpy: forward = fftn in y-slabs; backward = unnormalised ifftn in x-slabs
        if direction == FORWARD:
            F = np.fft.fftn(A)
            want = [F[:, q * g.yd: q * g.yd + g.n1l(q), :].transpose(1, 2, 0).reshape(-1) for q in range(P)]
        else:
            B = np.fft.ifftn(A) * A.size
            want = [B[p * g.xd: p * g.xd + g.n0l(p)].reshape(-1) for p in range(P)]
        for p in range(P):
            assert np.array_equal(outs[p], whole[p])
            assert np.abs(outs[p][: want[p].size] - want[p]).max() <= 1e-13 * np.log2(A.size) * scale
        # (2) the oracle, both buffers at every stage boundary
        for s in range(4):
            b1 = [b.copy() for b in ins]
            b2 = [np.zeros_like(b) for b in ins]
            co.slab_execute(g, b1, b2, direction, stop_after=s)
            for p in range(P):
                assert np.abs(b1[p] - dumps[p][s][0]).max() <= 1e-13 * np.log2(A.size) * scale, (s, p, "bufferDev1")
                assert np.abs(b2[p] - dumps[p][s][1]).max() <= 1e-13 * np.log2(A.size) * scale, (s, p, "bufferDev2")
        # (3) the product's recorded schedule (P2P fused, P2P two-sweep, NCCL), interpreted on the CPU
        for flags in (dfft.EXCHANGE_P2P, dfft.EXCHANGE_P2P | dfft.NO_FUSE, dfft.EXCHANGE_NCCL):
            got, _, _ = simulate(n0, n1, n2, P, direction, ins, flags)
            for p in range(P):
                n = want[p].size
                assert np.abs(got[p][:n] - outs[p][:n]).max() <= 1e-13 * np.log2(A.size) * scale, (flags, p)
        # (4) tables
        for p in range(P):
            mine = dfft.exchange_table(n0, n1, n2, P, p, direction)
            for j, k in enumerate(("scount", "soffset", "rcount", "roffset")):
                assert list(mine[k]) == list(tables[p, :, j])


def test_live_counts_and_policy_sweep(ref):
    rng = np.random.default_rng(5)
    for _ in range(200):
        P = int(rng.integers(1, 9))
        n0, n1, n2 = (int(rng.integers(1, 200)) for _ in range(3))
        if (P - 1) * -(-n0 // P) >= n0 or (P - 1) * -(-n1 // P) >= n1:
            continue
        for last in (False, True):
            assert dfft.getMaxDataCount(n0, n1, n2, P, last) == ref.max_data_count(n0, n1, n2, P, last)
    for n0 in range(1, 70):
        for w in range(1, 9):
            assert proper_device_num(n0, w) == ref.proper_device_num(n0, w), (n0, w)
