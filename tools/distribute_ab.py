"""A/B of the Crossword distribute kernels on the full cfg-4 workload (one GPU): times each variant alone with CUDA events and
checks that every byte of the five replica logs is identical between them.  Usage: python tools/distribute_ab.py [variants ...]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from summerset_b200 import workloads as wl  # noqa: E402
from summerset_b200.api import Context, ReedSolomon, cw_slot_pitch  # noqa: E402


def main():
    import os
    variants = [int(a) for a in sys.argv[1:]] or [0, 5]
    n = 1 << 20
    fixed = int(os.environ.get("SS_AB_LEN", "0"))
    dev = torch.device("cuda", 0)
    ctx = Context(0)
    D, P, NREP = (int(v) for v in os.environ.get("SS_AB_CODE", "3,2,5").split(","))      # code and population
    rs = ReedSolomon(ctx, D, P)
    lens, spr = wl.cfg4_lengths(n, seed_extra=0)
    if fixed:
        n = min(n, (12 << 30) // fixed)
        lens, spr = np.full(n, fixed, dtype=lens.dtype), spr[:n]
        print(f"all payloads {fixed} B, n = {n}")
    dj = (D + P) // NREP
    if (D, P, NREP) != (3, 2, 5):
        choices = np.arange(dj, D + 1, dj)
        spr = choices[np.random.default_rng(7).integers(0, len(choices), n)].astype(np.uint8)
        n = min(n, 1 << 19)
        lens, spr = lens[:n], spr[:n]
    lay = wl.ragged_layout(lens, D)
    gen = torch.Generator(device=dev); gen.manual_seed(1234)
    arena = torch.randint(0, 256, (lay["data_bytes"] + 256,), dtype=torch.uint8, device=dev, generator=gen)
    doff = torch.from_numpy(lay["data_off"].astype(np.int64)).to(dev)
    dlen = torch.from_numpy(lens.astype(np.int32)).to(dev)
    Lp = cw_slot_pitch(lay["L"].astype(np.int64)) if os.environ.get("SS_AB_PITCH16", "0") != "1" else (lay["L"].astype(np.int64) + 15) // 16 * 16
    slot_bytes = spr.astype(np.int64) * Lp
    rep_off = torch.from_numpy(np.concatenate([[0], np.cumsum(slot_bytes)[:-1]]).astype(np.int64)).to(dev)
    region = int(slot_bytes.sum() + 255) // 256 * 256
    spr_t = torch.from_numpy(spr).to(dev)
    alg = int((lay["L"].astype(np.int64) * (D + NREP * spr.astype(np.int64))).sum()) + n * 31
    logs = {}
    base = variants[0]
    for v in variants:
        logs[v] = torch.full((NREP, region), 0x5a, dtype=torch.uint8, device=dev) if (v == base or os.environ.get('SS_AB_CHECK', '1') == '1') else logs[base]
        ptrs = [logs[v][r].data_ptr() for r in range(NREP)]
        rs.set_variant(v)
        for _ in range(3):
            rs.crossword_distribute(arena, doff, dlen, spr_t, rep_off, ptrs)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(10):
            rs.crossword_distribute(arena, doff, dlen, spr_t, rep_off, ptrs)
        ev[1].record(); torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 10
        print(f"variant {v}: {rs.last_kernel():48s} {ms:8.3f} ms  {alg / ms / 1e6:8.1f} GB/s", flush=True)
    base = variants[0]
    for v in variants[1:]:
        if logs[v] is logs[base]:
            continue
        same = all(torch.equal(logs[base][r], logs[v][r]) for r in range(NREP))
        print(f"variant {v} vs {base}: all five logs identical = {same}", flush=True)
        assert same


if __name__ == "__main__":
    main()
