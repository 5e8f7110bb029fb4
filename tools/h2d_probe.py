"""Aggregate host<->device bandwidth of a multi-GPU box with every rank transferring at once (development aid):

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 tools/h2d_probe.py

Variants of the host buffer: cudaHostAlloc (portable), write-combined, mmap + cudaHostRegister, torch pinned."""
import ctypes
import mmap
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fds_b200 import _lib  # noqa: E402

rank = int(os.environ['RANK']); local = int(os.environ.get('LOCAL_RANK', rank)); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
lib = _lib.load()
rt = ctypes.CDLL('libcudart.so.12') if os.path.exists('/usr/local/cuda/lib64/libcudart.so.12') else None
NB = 2 << 30
dev = torch.empty(NB, dtype=torch.uint8, device='cuda')


def host_tensor(kind):
    if kind == 'torch':
        return torch.empty(NB, dtype=torch.uint8).pin_memory(), None
    if kind in ('alloc', 'wc'):
        p = (lib.fds_host_alloc_wc if kind == 'wc' else lib.fds_host_alloc)(NB)
        a = np.ctypeslib.as_array((ctypes.c_uint8 * NB).from_address(p))
        if kind == 'alloc':
            a[::4096] = 1
        return torch.from_numpy(a), p
    if kind == 'register':
        t = torch.empty(NB, dtype=torch.uint8)
        t[::4096] = 1
        torch.cuda.cudart().cudaHostRegister(t.data_ptr(), NB, 0)
        return t, None
    raise ValueError(kind)


def run(kind, direction):
    h, keep = host_tensor(kind)
    s = torch.cuda.Stream()
    for it in range(3):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s):
            if direction == 'h2d':
                dev.copy_(h, non_blocking=True)
            elif direction == 'd2h':
                h.copy_(dev, non_blocking=True)
        s.synchronize()
        dt = time.perf_counter() - t0
        dist.barrier()
    t = torch.tensor([dt], device='cuda', dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print('%-9s %s: slowest rank %.3f s -> %.1f GB/s per GPU, %.0f GB/s aggregate' %
              (kind, direction, t.item(), NB / t.item() / 1e9, world * NB / t.item() / 1e9), flush=True)
    if kind == 'register':
        torch.cuda.cudart().cudaHostUnregister(h.data_ptr())
    elif keep is not None:
        lib.fds_host_free(keep)
    del h


for kind in ('alloc', 'wc', 'register', 'torch'):
    for direction in (('h2d',) if kind == 'wc' else ('h2d', 'd2h')):
        try:
            run(kind, direction)
        except Exception as e:                       # noqa: BLE001
            if rank == 0:
                print(kind, direction, 'failed:', repr(e)[:200], flush=True)
dist.barrier()
dist.destroy_process_group()
