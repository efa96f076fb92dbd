#!/usr/bin/env python
"""PCIe ceiling probe for the e2e (host-buffer) path: H2D / D2H bandwidth with pinned and write-combined host memory,
alone and simultaneously, and the e2e encode call at several chunk sizes.  Run on a GPU box."""
import ctypes as C
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from summerset_b200.api import Context, ReedSolomon  # noqa: E402
from summerset_b200._lib import check  # noqa: E402

ctx = Context(0)
dev = torch.device("cuda", 0)
N = 1 << 32
d = torch.empty(N, dtype=torch.uint8, device=dev)
d2 = torch.empty(N, dtype=torch.uint8, device=dev)
hp = torch.empty(N, dtype=torch.uint8, pin_memory=True)
ho = torch.empty(N, dtype=torch.uint8, pin_memory=True)
p = C.c_void_p()
check(ctx.lib.ss_host_alloc_wc(ctx.h, N, C.byref(p)))
s1 = torch.cuda.Stream(); s2 = torch.cuda.Stream()


def t_copy(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    return N / best / 1e9


print("H2D pinned      %.1f GB/s" % t_copy(lambda: d.copy_(hp, non_blocking=True)))
print("D2H pinned      %.1f GB/s" % t_copy(lambda: ho.copy_(d, non_blocking=True)))
print("H2D write-comb  %.1f GB/s" % t_copy(lambda: check(ctx.lib.ss_copy_h2d(ctx.h, d.data_ptr(), p.value, N))))


def both():
    with torch.cuda.stream(s1):
        d.copy_(hp, non_blocking=True)
    with torch.cuda.stream(s2):
        ho.copy_(d2, non_blocking=True)


print("H2D + D2H simultaneous, each direction  %.1f GB/s" % t_copy(both))
rs = ReedSolomon(ctx, 3, 2)
n = 1 << 20
hv = hp.numpy().reshape(n, 4096)
par = torch.empty((2, n, 1376), dtype=torch.uint8, pin_memory=True).numpy()
for mb in (16, 64, 256):
    os.environ["SS_E2E_CHUNK_MB"] = str(mb)
    rs.encode_uniform_host(hv, 4096, par)
    t = time.perf_counter()
    for _ in range(3):
        rs.encode_uniform_host(hv, 4096, par)
    dt = (time.perf_counter() - t) / 3
    print("ss_rs_encode_uniform host call, chunk %3d MiB: %.1f ms  (H2D %.1f GB/s equivalent)" % (mb, dt * 1e3, n * 4096 / dt / 1e9))
