#!/usr/bin/env python3
"""NVLink/NVSwitch microbenchmarks for the data-parallel step (torchrun, one rank per GPU): per-rank inbound/outbound
GB/s of the access patterns a reduce-scatter / all-gather can be built from. All ranks run every pattern at once
(as the real step does). Build the helper first: nvcc ... -o tools/libp2p_probe.so tools/p2p_probe.cu"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libp2p_probe.so"))
lib.p2p_probe.restype = ctypes.c_int
lib.p2p_probe.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int, ctypes.c_ulonglong, ctypes.c_ulonglong,
                          ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_int, ctypes.c_void_p]
N = 12_800_000                       # floats (51 MB), like the gradient arena
buf = symm.empty(N, dtype=torch.float32, device=dev)
buf.fill_(1.0)
hdl = symm.rendezvous(buf, dist.group.WORLD)
ptrs = [int(p) for p in hdl.buffer_ptrs]
mc = int(hdl.multicast_ptr) if hdl.multicast_ptr else 0
loc = torch.ones(N, dtype=torch.float32, device=dev)
shard = N // world // 1024 * 1024
off4, n4 = rank * shard // 4, shard // 4
if rank == 0:
    print(f"world {world}; multicast supported: {bool(mc)}; shard {shard * 4 / 1e6:.1f} MB per rank")


def run(name, kind, srcs, nbytes_in, nbytes_out, grid=592, use_mc=False, n4_=None, off4_=None):
    if use_mc and not mc:
        return
    arr = (ctypes.c_ulonglong * 8)(*(srcs + [0] * (8 - len(srcs))))
    st = torch.cuda.current_stream().cuda_stream
    a, b = (n4 if n4_ is None else n4_), (off4 if off4_ is None else off4_)
    for _ in range(3):
        assert lib.p2p_probe(kind, arr, len(srcs), mc, loc.data_ptr(), a, b, grid, st) == 0
    torch.cuda.synchronize(); dist.barrier()
    ts = []
    for _ in range(10):
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lib.p2p_probe(kind, arr, len(srcs), mc, loc.data_ptr(), a, b, grid, st); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = torch.tensor(sorted(ts)[len(ts) // 2], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        us = float(t.item())
        print(f"  {name:58s} {us:7.1f} us   in {nbytes_in / us / 1e3:6.0f} GB/s   out {nbytes_out / us / 1e3:6.0f} GB/s per rank")


remote = [ptrs[p] for p in range(world) if p != rank]
rs_bytes = shard * 4 * (world - 1)
if rank == 0:
    print("reduce-scatter patterns (each rank reads its shard from every peer):")
run("ld.relaxed.sys.v4 from all peers (+local)", 0, ptrs, rs_bytes, rs_bytes)
run("ld.global.nc.v4 from all peers (+local)", 1, ptrs, rs_bytes, rs_bytes)
run("ld.global.nc.v4, grid 148*8", 1, ptrs, rs_bytes, rs_bytes, grid=1184)
run("multimem.ld_reduce.add.v4.f32 (in-switch reduction)", 2, [], shard * 4, rs_bytes, use_mc=True)
if rank == 0:
    print("all-gather patterns (each rank writes its shard into every peer):")
run("st.global.v4 to all peers (+local)", 3, ptrs, rs_bytes, rs_bytes)
run("st.global.v2 (8 B per thread) to all peers (+local)", 5, ptrs, rs_bytes // 2, rs_bytes // 2)
run("multimem.st.v4 (in-switch multicast)", 4, [], rs_bytes, shard * 4, use_mc=True)
# copy engines: push my piece for every peer
if rank == 0:
    print("copy engines:")
for streams in (1, 4):
    ss = [torch.cuda.Stream(dev) for _ in range(streams)]
    ts = []
    for _ in range(8):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        evs = []
        for i, p in enumerate(q for q in range(world) if q != rank):
            s = ss[i % streams]
            s.wait_event(e0)
            with torch.cuda.stream(s):
                dst = hdl.get_buffer(p, (shard,), torch.float32, rank * shard) if hasattr(hdl, "get_buffer") else None
                dst.copy_(loc[p * shard:(p + 1) * shard], non_blocking=True)
                ev = torch.cuda.Event(); ev.record(s); evs.append(ev)
        for ev in evs:
            torch.cuda.current_stream().wait_event(ev)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = torch.tensor(sorted(ts)[len(ts) // 2], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        us = float(t.item())
        print(f"  cudaMemcpyAsync pushes of one shard to every peer, {streams} stream(s)   {us:7.1f} us   out {rs_bytes / us / 1e3:6.0f} GB/s per rank")
dist.destroy_process_group()
