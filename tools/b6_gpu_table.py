#!/usr/bin/env python
"""B6 on the GPU side (benches/rse_bench.rs:17-26,161-203): RS(3,2) from_data + compute_parity of ONE
codeword at 4 KB ... 4 MB, next to the CPU figures of tests/cpu_baseline_table.py.

Three numbers per size:
  call     ss_rs_encode with HOST slices -- what the GpuReedSolomon drop-in of INTEGRATION.md section 3 executes per request:
           d H2D copies, one kernel, p D2H copies, one stream synchronisation (latency per call, median of many)
  kernel   the same single codeword already resident in HBM, device-timed (CUDA events)
  batch    the batch size at which the batched device-resident call (ss_rs_encode_uniform_dev) takes less time per codeword
           than the CPU's single-thread AVX2 figure measured in the same process (oracle; baseline only)
Writes a text table to stdout.   python tools/b6_gpu_table.py
"""
from __future__ import annotations

import statistics
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import pyoracle as oracle  # noqa: E402  (CPU baseline column only)
from summerset_b200 import workloads as wl  # noqa: E402
from summerset_b200.api import Context, ReedSolomon, round_up, shard_len  # noqa: E402


def main():
    assert torch.cuda.is_available(), "needs a GPU"
    torch.cuda.set_device(0)
    ctx = Context(0)
    rs = ReedSolomon(ctx, 3, 2)
    mode = 1 if oracle.have_avx2() else 0
    print("# RS(3,2) from_data + compute_parity of ONE codeword (mirrors benches/rse_bench.rs sizes)")
    print(f"# {'size':>9s}  {'CPU 1 thread':>13s}  {'GPU call (host slices)':>23s}  {'GPU kernel (resident)':>22s}  {'kernel GB/s':>11s}  "
          f"{'kernel':>34s}  batch at which the GPU wins per codeword")
    for size in [4096, 16 << 10, 64 << 10, 256 << 10, 1 << 20, 4 << 20]:
        one = wl.payload_uniform(1, size, alphanumeric=True, seed_extra=size)
        L = shard_len(size, 3)
        # CPU, one thread
        reps = max(5, (64 << 20) // size // 4)
        t0 = time.perf_counter()
        for _ in range(reps):
            oracle.rs_encode_uniform(3, 2, one, size, mode=mode, threads=1)
        cpu_ms = (time.perf_counter() - t0) / reps * 1e3
        # GPU call with host slices
        padded = np.zeros(3 * L, dtype=np.uint8)
        padded[:size] = one[0, :size]
        shards = [np.ascontiguousarray(padded[i * L:(i + 1) * L]) for i in range(3)] + [np.zeros(L, dtype=np.uint8) for _ in range(2)]
        for _ in range(5):
            rs.encode(shards)
        lat = []
        for _ in range(200 if size <= (256 << 10) else 50):
            t0 = time.perf_counter(); rs.encode(shards); lat.append(time.perf_counter() - t0)
        want = oracle.rs_encode_uniform(3, 2, one, size)
        assert (shards[3] == want[0, 0, :L]).all() and (shards[4] == want[1, 0, :L]).all()
        call_ms = statistics.median(lat) * 1e3
        # resident single codeword, device-timed
        ds = round_up(L, 16)
        d_data = torch.from_numpy(one).cuda()
        par = torch.empty((2, 1, ds), dtype=torch.uint8, device="cuda")
        for _ in range(5):
            rs.encode_uniform(d_data, size, parity=par)
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        iters = 200
        for _ in range(iters):
            rs.encode_uniform(d_data, size, parity=par)
        b.record(); torch.cuda.synchronize()
        k_ms = a.elapsed_time(b) / iters
        kernel = rs.last_kernel()
        assert (par.cpu().numpy()[:, 0, :L] == want[:, 0, :L]).all()
        # batch size at which the batched call beats the CPU per codeword
        win = None
        for n in [1, 2, 4, 8, 16, 32, 64, 128, 256, 1024, 4096]:
            if n * size > (2 << 30):
                break
            dd = d_data.repeat(n, 1).contiguous()
            pp = torch.empty((2, n, ds), dtype=torch.uint8, device="cuda")
            for _ in range(3):
                rs.encode_uniform(dd, size, parity=pp)
            torch.cuda.synchronize()
            a.record()
            for _ in range(20):
                rs.encode_uniform(dd, size, parity=pp)
            b.record(); torch.cuda.synchronize()
            per = a.elapsed_time(b) / 20 / n
            if per < cpu_ms:
                win = (n, per)
                break
        print(f"  {size:9d}  {cpu_ms:10.4f} ms  {call_ms:20.4f} ms  {k_ms:19.4f} ms  {5 * L / (k_ms * 1e-3) / 1e9:11.1f}  {kernel:>34s}  "
              + (f"n >= {win[0]} ({win[1] * 1e3:.2f} us per codeword)" if win else "never within 4096"))


if __name__ == "__main__":
    main()
