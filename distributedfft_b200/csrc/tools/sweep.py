#!/usr/bin/env python
"""Developer tool: time forward transforms for several (size, precision, DFFT_VARIANT, flags) in one process.
Single GPU:   python sweep.py 512:double:0 512:double:1 ...
Multi GPU:    torchrun --nproc-per-node P sweep.py ...     (P2P exchange, process per GPU)
An item is  size:precision:variant[:fuse|nofuse|nopipe|nccl|...][;ENV=value;ENV=value...]   (flags joined with '+')"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", ".."))
import torch
import distributedfft_b200 as dfft

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
comm = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("gloo")
    def ag(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out
    comm = dfft.BootstrapComm(rank, world, ag)

FLAGS = {"fuse": dfft.FORCE_FUSE, "nofuse": dfft.NO_FUSE, "nopipe": dfft.NO_PIPELINE, "nccl": dfft.EXCHANGE_NCCL, "p2p": dfft.EXCHANGE_P2P,
         "overlap": dfft.OVERLAP_X, "": 0}
SWEEP_ENV = set()
for item in sys.argv[1:]:
    head, *envs = item.split(";")
    for k in SWEEP_ENV:
        os.environ.pop(k, None)
    for kv in envs:
        k, v = kv.split("=")
        os.environ[k] = v
        SWEEP_ENV.add(k)
    parts = head.split(":")
    n, precs, var = int(parts[0]), parts[1], parts[2]
    flags = 0
    if len(parts) > 3:
        for f in parts[3].split("+"):
            flags |= FLAGS[f]
    os.environ["DFFT_VARIANT"] = var
    prec = dfft.DOUBLE if precs == "double" else dfft.FLOAT
    tdt = torch.complex128 if prec == dfft.DOUBLE else torch.complex64
    maxc = dfft.getMaxDataCount(n, n, n, world, rank == world - 1)
    tin = torch.empty(maxc, dtype=tdt, device=dev)
    torch.view_as_real(tin).uniform_(0.0, 1.0)
    tout = torch.empty(maxc, dtype=tdt, device=dev)
    torch.cuda.synchronize(dev)
    plan = dfft.fft_mpi_plan_dft_c2c_3d(n, n, n, tin.data_ptr(), tout.data_ptr(), comm, rank, world, dfft.FORWARD, prec, flags)
    stream = torch.cuda.ExternalStream(plan.stream, device=dev)
    for _ in range(3):
        plan.execute()
    plan.synchronize()
    if world > 1:
        dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    steps = 20
    e0.record(stream)
    for _ in range(steps):
        plan.execute()
    e1.record(stream)
    plan.synchronize()
    ms = e0.elapsed_time(e1) / steps
    pt = plan.pass_timings(); st = plan.timings()
    if world > 1:
        t = torch.tensor([ms] + pt + st, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, pt, st = t[0].item(), t[1:4].tolist(), t[4:].tolist()
    if rank == 0:
        print(f"{item:60s} P={world} fused={int(plan.fused)} parts={plan.pipeline_parts} ms/step {ms:8.4f}  pass {[round(x, 4) for x in pt]}  stage {[round(x, 4) for x in st[:4]]}", flush=True)
    if os.environ.get("DFFT_DEBUG_TIMELINE") and plan.overlapped:
        tl = plan.debug_timeline()
        print(f"   rank {rank} timeline us: phase0_end {tl[0]:.1f} first_X {tl[1]:.1f} end {tl[2]:.1f} | us/tile Z {tl[3]:.2f} Y {tl[4]:.2f} X {tl[5]:.2f} | tiles {tl[6]:.0f} {tl[7]:.0f} {tl[8]:.0f} | wait mean {tl[9]:.1f} max {tl[10]:.1f}", flush=True)
    plan.destroy()
    del tin, tout
    torch.cuda.empty_cache()
if comm is not None:
    comm.destroy()
    dist.barrier()
    dist.destroy_process_group()
