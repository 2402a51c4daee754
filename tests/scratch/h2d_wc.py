"""Host-to-device copy rate of N ranks at once from ordinary pinned memory and from write-combined pinned memory
(cudaHostAllocWriteCombined): is the box's host-link ceiling (bench.py: e2e.copy_only_*) a property of the memory type?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tests/scratch/h2d_wc.py
"""
import ctypes as C
import glob
import os
import site
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
path = None
for sp in site.getsitepackages():
    g = glob.glob(sp + "/nvidia/cuda_runtime/lib/libcudart.so*")
    if g:
        path = g[0]
rt = C.CDLL(path)
rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
SIZE = 211 << 20
d = torch.empty(SIZE, dtype=torch.uint8, device="cuda")
d2 = torch.empty(36 << 20, dtype=torch.uint8, device="cuda")
out = {}
for name, flags in (("pinned", 0), ("write_combined", 4), ("pinned", 0), ("write_combined", 4)):
    p, q = C.c_void_p(), C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), SIZE, flags) == 0
    assert rt.cudaHostAlloc(C.byref(q), 36 << 20, 0) == 0
    C.memset(p, 1, SIZE)
    s2 = torch.cuda.Stream()
    for it in range(3):
        rt.cudaMemcpyAsync(d.data_ptr(), p, SIZE, 1, None)
        rt.cudaDeviceSynchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 20
    for it in range(K):
        rt.cudaMemcpyAsync(d.data_ptr(), p, SIZE, 1, None)                      # H2D on the null stream
        rt.cudaMemcpyAsync(q, d2.data_ptr(), 36 << 20, 2, C.c_void_p(s2.cuda_stream))  # D2H alongside, like a step's records
        rt.cudaDeviceSynchronize()
    if world > 1:
        dist.barrier()
    dt = (time.perf_counter() - t0) / K
    t = torch.tensor([dt], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{name:15s} {1e3 * t.item():7.3f} ms per step (211 MB in + 36 MB out per rank), {world * (SIZE + (36 << 20)) / t.item() / 1e9:6.1f} GB/s for the box", flush=True)
    rt.cudaFreeHost(p)
    rt.cudaFreeHost(q)
if world > 1:
    dist.destroy_process_group()
