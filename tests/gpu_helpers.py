"""Helpers for the -m gpu parity tests: drive libdfft.so through its C ABI (via the ctypes
binding) on torch-allocated device buffers, one host thread per device like the reference driver
(3dmpifft_opt/fftSpeed3d_c2c.cpp:49-51)."""
import threading

import numpy as np
import torch

import distributedfft_b200 as dfft
from oracle import SlabGeometry

CDT = {dfft.DOUBLE: (np.complex128, torch.complex128), dfft.FLOAT: (np.complex64, torch.complex64)}


def run_slab(n0, n1, n2, P, direction, inputs, precision=dfft.DOUBLE, flags=0, repeat=1, stages=None, inplace=False, refill=True):
    """inputs[p]: numpy array (max_count) for device p's bufferDev1.  Returns dict with per-device
    'buf1', 'buf2' numpy copies after execution (and after each stage when `stages`)."""
    g = SlabGeometry(n0, n1, n2, P)
    npdt, tdt = CDT[precision]
    comm = dfft.LocalComm(P) if P > 1 else None
    res = [None] * P
    errs = []

    def worker(p):
        try:
            torch.cuda.set_device(p)
            dev = torch.device("cuda", p)
            mc = g.max_count(p)
            tin = torch.zeros(mc, dtype=tdt, device=dev)
            tin[: inputs[p].size] = torch.from_numpy(np.ascontiguousarray(inputs[p], dtype=npdt)).to(dev)
            tout = tin if inplace else torch.zeros(mc, dtype=tdt, device=dev)
            torch.cuda.synchronize(dev)
            plan = dfft.fft_mpi_plan_dft_c2c_3d(n0, n1, n2, tin.data_ptr(), None if inplace else tout.data_ptr(), comm, p, P,
                                                direction, precision, flags)
            out = {}

            def fetch(ptr, count):
                host = np.empty(count, dtype=npdt)
                dfft.memcpy_dtoh(host.ctypes.data, ptr, host.nbytes)
                return host

            if stages is not None:
                out["stages"] = []
                for s in stages:
                    plan.execute_stage(s)
                    out["stages"].append((fetch(plan.bufferDev1, mc), fetch(plan.bufferDev2, mc)))
            else:
                for _ in range(repeat):
                    if _ > 0 and refill:  # refill bufferDev1 like the reference driver (fftSpeed3d_c2c.cpp:78)
                        plan.synchronize()
                        torch.cuda.synchronize(dev)
                        h = np.zeros(mc, dtype=npdt); h[: inputs[p].size] = inputs[p]
                        dfft.memcpy_htod(plan.bufferDev1, h.ctypes.data, h.nbytes)
                    plan.execute()
                plan.synchronize()
                out["timings"] = plan.timings()
                out["launches"] = plan.launches
                out["exchange"] = plan.exchange
                out["fused"] = plan.fused
            out["buf1"] = fetch(plan.bufferDev1, mc)
            out["buf2"] = fetch(plan.bufferDev2, mc)
            out["counts"] = (plan.in_count, plan.out_count, plan.maxDataCountInDevice)
            plan.destroy()
            res[p] = out
        except Exception as exc:  # noqa
            import traceback
            errs.append((p, traceback.format_exc()))

    if P == 1:
        worker(0)
    else:
        th = [threading.Thread(target=worker, args=(p,)) for p in range(P)]
        for t in th: t.start()
        for t in th: t.join()
    if comm is not None:
        comm.destroy()
    if errs:
        raise RuntimeError("device thread failed:\n" + "\n".join(e for _, e in errs))
    return res
