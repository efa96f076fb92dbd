#!/usr/bin/env python
"""bench.py -- throughput of the quorum-tally + Reed-Solomon accept path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2|cfg5]

A "step" is one pass of the hot path over one batch of synthetic input.  Default workload (cfg3,
BASELINE.json configs[2], the configuration the north-star target is quoted on): the fused RSPaxos accept
step -- RS(3,2) encode of 2^20 groups' 4096-byte request batches + quorum tally (4 of 5) of each group's
64-slot ack window -- ONE kernel launch per step, inputs resident in HBM.  The JSON line also carries the
cfg2 (MultiPaxos tally only) measurement under "cfg2".

  value      RS shard GB/s = (d+p)*L bytes per codeword * codewords / time   (whole job, all ranks)
  e2e        same metric through the host-buffer C-ABI calls (pinned host memory, H2D + D2H inside)
  roofline   algorithmic bytes per launch / CUDA-event kernel time vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference: the CPU oracle port of the reference path on this box's cores
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

D, P, DATA_LEN = 3, 2, 4096
R, THRESH_RSPAXOS, THRESH_MULTIPAXOS = 5, 4, 3
G_PER_GPU = 1 << 20
METRIC = "RS shard GB/s on the fused RSPaxos accept step (RS(3,2) encode + quorum tally); consensus slots committed/s in slots_committed_per_s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg2", "cfg5", "cfg3b", "cfg4"])
    ap.add_argument("--groups", type=int, default=G_PER_GPU, help="groups per GPU")
    ap.add_argument("--variant", type=int, default=0, help="encode kernel variant (tuning)")
    ap.add_argument("--replicas", type=int, default=5, help="cfg3 variant: population n (RS(majority, n-majority), f=(n//2)//2)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="N>1: how shard planes reach the simulated peers")
    ap.add_argument("--no-tally", action="store_true", help="tuning: time the encode alone")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def measured_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        j = json.loads(f.read_text())
        return float(j["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def copy_bandwidth_here(torch, dev):
    """STREAM-style copy on THIS box, measured the way MEASURED_PEAKS.json was (b.copy_(a), 1 Gi bf16 elements,
    read+write bytes, best of 10): boxes of the pool differ by several percent, so the live figure is reported
    beside the pool-wide denominator."""
    try:
        a = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev)
        b = torch.empty_like(a)
        a.fill_(1.0)
        best = 0.0
        for _ in range(10):
            s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
            s.record(); b.copy_(a); e.record(); torch.cuda.synchronize()
            best = max(best, 2 * a.numel() * 2 / (s.elapsed_time(e) * 1e-3) / 1e9)
        del a, b
        return best
    except Exception:
        return None


def profile_traffic(workload: str):
    """dram bytes per launch from the committed ncu --set full capture (profiles/), or None."""
    f = ROOT / "profiles" / "traffic.json"
    if f.exists():
        try:
            return json.loads(f.read_text()).get(workload)
        except Exception:
            return None
    return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nme, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# =================================================================================================
# CPU arm: the oracle port of the reference path, all host threads
# =================================================================================================
def cpu_arm(steps: int, warmup: int, sample_cw: int, workload: str):
    from oracle import pyoracle as oracle
    from summerset_b200 import workloads as wl
    # hardware threads this process may run on -- NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1
    hw_threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    threads = hw_threads
    mode = 1 if oracle.have_avx2() else 0
    L = oracle.cw_shard_len(DATA_LEN, D)
    ds = (L + 15) // 16 * 16
    data = wl.payload_uniform(sample_cw, DATA_LEN, seed_extra=7)
    planes = wl.cfg2_planes(sample_cw, R, 0.9, seed_extra=7)
    parity = np.zeros((P, sample_cw, ds), dtype=np.uint8)
    off = np.arange(sample_cw, dtype=np.uint64) * np.uint64(data.shape[1])
    lens = np.full(sample_cw, DATA_LEN, dtype=np.uint32)
    poff = np.arange(sample_cw, dtype=np.uint64) * np.uint64(ds)

    def step():
        if workload != "cfg2":
            oracle.rs_encode_batch(D, P, data.reshape(-1), off, lens, parity.reshape(-1), sample_cw * ds, poff, mode, threads)
        oracle.tally_planes(planes, THRESH_RSPAXOS if workload != "cfg2" else THRESH_MULTIPAXOS, threads)

    # "all the host threads it can use": oversubscribing SMT siblings / a second socket can be slower than fewer
    # threads for this memory-bound loop, so the thread count is calibrated (best of max, max/2, max/4) and stated.
    step()
    best = None
    for cand in sorted({hw_threads, max(1, hw_threads // 2), max(1, hw_threads // 4), max(1, hw_threads // 8)}):
        threads = cand
        step()                                  # the first call at a new team size pays for creating the team
        dtc = 1e30
        for _ in range(3):
            t0 = time.perf_counter(); step(); dtc = min(dtc, time.perf_counter() - t0)
        if best is None or dtc < best[0]:
            best = (dtc, cand)
    threads = best[1]

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    gbs = (D + P) * L * sample_cw / dt / 1e9
    slots = sample_cw * 64 / dt
    return dict(ms_per_step=dt * 1e3, gbs=gbs, slots_per_s=slots, threads=threads,
                path="AVX2 vpshufb nibble tables (what rse-simd enables)" if mode else "scalar MUL_TABLE",
                sample=f"{sample_cw} codewords x {DATA_LEN} B (1/{max(1, G_PER_GPU // sample_cw)} of the workload) + their 64-slot ack windows, per step")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 1 << 18   # 1 GiB of payload per step: large enough to amortise thread wake-ups in a VM
    r = cpu_arm(max(1, args.steps), max(1, args.warmup), sample, args.workload)
    value = r["gbs"] if args.workload != "cfg2" else r["slots_per_s"]
    unit = "GB/s" if args.workload != "cfg2" else "slots/s"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config_dict(args, 1),
        "slots_committed_per_s": r["slots_per_s"],
        "cpu_baseline": {"value": value, "unit": unit, "cores": r["threads"], "kind": "port", "sample": r["sample"],
                         "path": r["path"],
                         "note": "C restatement of the reference's Rust path (oracle/ss_oracle.c); the reference "
                                 "itself cannot be built here (no cargo/rustc)"},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def config_dict(args, world):
    L = (DATA_LEN + D - 1) // D
    return {"workload": {"cfg3": f"cfg3: RSPaxos ({D},{R}) fused RS({D},{P}) encode + quorum tally ({THRESH_RSPAXOS} of {R}), 2^20 groups x 4096 B "
                                 "request batch + 64-slot ack window per GPU",
                         "cfg2": "cfg2: MultiPaxos 5-replica quorum tally (3 of 5), 2^20 groups x 64 slots per GPU",
                         "cfg5": "cfg5: Raft 7-replica match-index commit scan, 2^22 groups, 64-slot term window"}[args.workload],
            "groups_per_gpu": args.groups, "data_len": DATA_LEN, "rs": [D, P], "shard_len": L, "replicas": R,
            "sharding": f"groups x{world} (independent shards)" + ("" if world == 1 else " + shard planes written to the simulated peers' GPUs over NVLink by the encode kernel (ack planes copied back)"),
            "l2": "inputs (4 GiB payload + 2.9 GB parity per GPU) exceed the 126 MB L2; no flush needed"}


# =================================================================================================
# GPU arm
# =================================================================================================
def run_ours(args):
    import torch
    import torch.distributed as dist
    from summerset_b200 import sharding, workloads as wl
    from summerset_b200.api import Context, ReedSolomon, SS_RS_OUT_PADDED16
    from summerset_b200._lib import check

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = Context(local)
    rs = ReedSolomon(ctx, D, P)
    rs.set_variant(args.variant)
    n = args.groups
    L, ds, ps = rs.parity_layout(DATA_LEN, n)
    peak, peak_src = measured_peaks()
    if args.workload == "cfg3b":
        return run_cfg3b(args, ctx, rs, dev, world, rank, peak, peak_src)
    if args.workload == "cfg4":
        return run_cfg4(args, ctx, rs, dev, world, rank, peak, peak_src)

    # ---- synthetic inputs, generated on the device (seeded) ----
    gen = torch.Generator(device=dev); gen.manual_seed(wl.SEED_BASE + 3 + 1000 * rank)
    data = torch.randint(0, 256, (n, DATA_LEN), dtype=torch.uint8, device=dev, generator=gen)
    shards = torch.zeros((D + P, n, ds), dtype=torch.uint8, device=dev) if world > 1 else None
    parity = shards[D:] if world > 1 else torch.empty((P, n, ds), dtype=torch.uint8, device=dev)
    # ack planes: leader always acks, followers with p = 0.9 (230/256)
    u = torch.randint(0, 256, (R, n, 64), dtype=torch.uint8, device=dev, generator=gen) < 230
    w = torch.tensor([1 << i for i in range(63)] + [-(1 << 63)], dtype=torch.int64, device=dev)
    planes = (u.to(torch.int64) * w).sum(dim=2)
    planes[0] = -1
    del u
    committed = torch.empty(n, dtype=torch.int64, device=dev)
    bar = torch.empty(n, dtype=torch.int32, device=dev)
    recv = torch.empty((R, n, ds), dtype=torch.uint8, device=dev) if world > 1 else None
    ack_recv = torch.empty((R, n), dtype=torch.int64, device=dev) if world > 1 else None
    rounds = sharding.exchange_rounds(R, world, rank) if world > 1 else []
    flags = SS_RS_OUT_PADDED16 | (2 if world > 1 else 0)
    p2p = None
    if world > 1 and args.exchange == "p2p":
        # every rank hosts R shard planes (replica r of the groups led from rank (me - r) % world) and the R ack
        # planes of its own groups; peers map both through CUDA IPC and the kernels store straight into them.
        log = ctx.dev_alloc(R * n * ds)
        acks = ctx.dev_alloc(R * n * 8)
        handles = [None] * world
        dist.all_gather_object(handles, (ctx.ipc_export(log), ctx.ipc_export(acks)))
        peer_log, peer_acks = {}, {}
        for q in range(world):
            if q == rank:
                peer_log[q], peer_acks[q] = log, acks
            else:
                peer_log[q] = ctx.ipc_open(handles[q][0], R * n * ds)
                peer_acks[q] = ctx.ipc_open(handles[q][1], R * n * 8)
        shard_ptrs = [peer_log[sharding.replica_rank(rank, r, world)].ptr + r * n * ds for r in range(R)]
        acks_t = acks.tensor().view(torch.int64).view(R, n)
        acks_t.copy_(planes)
        tiny = torch.zeros(1, dtype=torch.int32, device=dev)
        p2p = dict(log=log, acks=acks, peer_log=peer_log, peer_acks=peer_acks, shard_ptrs=shard_ptrs, acks_t=acks_t)
        dist.barrier()

    def exchange():
        # shard plane r of my groups -> rank (rank + r) % world ; then each simulated follower acks:
        # its ack plane (seeded drop mask = my `planes[r]` of the home rank, sent along) returns home.
        for rd in rounds:
            ins = [shards[rd["send"][dst]] if rd["send"][dst] >= 0 else shards[0][:0] for dst in range(world)]
            outs = [recv[rd["recv"][src]] if rd["recv"][src] >= 0 else recv[0][:0] for src in range(world)]
            dist.all_to_all(outs, ins)
        for rd in rounds:
            # follower on rank dst accepted shard r from home `src`; its ack plane travels back the other way
            ins = [planes[rd["recv"][src]] if rd["recv"][src] >= 0 else planes[0][:0] for src in range(world)]
            outs = [ack_recv[rd["send"][dst]] if rd["send"][dst] >= 0 else ack_recv[0][:0] for dst in range(world)]
            dist.all_to_all(outs, ins)

    def step():
        if args.workload == "cfg2":
            ctx.tally_planes(planes, THRESH_MULTIPAXOS, True, committed, bar)
            return
        if args.no_tally:
            check(ctx.lib.ss_rs_encode_uniform_dev(rs.h, data.data_ptr(), DATA_LEN, DATA_LEN, n, parity.data_ptr(), ps, ds, flags))
            return
        if p2p is not None:
            # ONE kernel: tally the ack planes the followers wrote last step, RS-encode, and store shard r of
            # every local group into the HBM of the GPU that simulates replica r (NVLink stores)
            rs.accept_step_replicate(data, DATA_LEN, p2p["shard_ptrs"], ds, p2p["acks_t"], THRESH_RSPAXOS, committed, bar)
            # the simulated follower (home h, replica r) on this rank acks: its ack plane goes to h's ack buffer
            for r in range(R):
                h = (rank - r) % world
                ctx.copy_d2d(p2p["peer_acks"][h].ptr + r * n * 8, planes[r].data_ptr(), n * 8)
            dist.all_reduce(tiny)          # stream-ordered barrier: every rank's stores of this step are done
            return
        check(ctx.lib.ss_accept_step_fused_dev(rs.h, data.data_ptr(), DATA_LEN, DATA_LEN, n, parity.data_ptr(), ps, ds,
                                               flags, (ack_recv if world > 1 else planes).data_ptr(), R,
                                               THRESH_RSPAXOS, committed.data_ptr(), bar.data_ptr()))
        if world > 1:
            exchange()

    if world > 1:
        ack_recv.copy_(planes)

    # ---- cfg5 (Raft), cfg3b (reconstruct) and cfg4 (Crossword ragged) are separate small harnesses ----
    if args.workload == "cfg5":
        return run_cfg5(args, ctx, dev, world, rank, peak, peak_src)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = ctx.launches
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    k_evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(args.steps):
        step()
    t_end.record()
    barrier()
    total_ms = t_start.elapsed_time(t_end)
    launches = ctx.launches - launches0
    clocks = sampler.stop() if sampler else None
    # kernel-only duration (CUDA events around the single launch), measured in a second pass so the events
    # do not perturb the whole-step timing above
    kt = []
    for i in range(args.steps):
        a, b = k_evs[i]
        a.record()
        if world == 1:
            step()
        elif p2p is not None:
            rs.accept_step_replicate(data, DATA_LEN, p2p["shard_ptrs"], ds, p2p["acks_t"], THRESH_RSPAXOS, committed, bar)
        else:
            check(ctx.lib.ss_accept_step_fused_dev(rs.h, data.data_ptr(), DATA_LEN, DATA_LEN, n, parity.data_ptr(), ps, ds,
                                                   flags, ack_recv.data_ptr(), R, THRESH_RSPAXOS, committed.data_ptr(),
                                                   bar.data_ptr()))
        b.record()
    torch.cuda.synchronize()
    kt = [a.elapsed_time(b) for a, b in k_evs]
    kernel_ms = sum(kt) / len(kt)
    if world > 1:
        t = torch.tensor([total_ms, kernel_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, kernel_ms = float(t[0]), float(t[1])
    ms_per_step = total_ms / args.steps

    # ---- parity spot-check of what was just timed.  The oracle is used here ONLY as the checker of the
    #      timed outputs (outside the timed region); nothing measured or shipped routes through it. ----
    check_note = None
    if rank == 0:
        from oracle import pyoracle as oracle
        idx = torch.arange(0, n, max(1, n // 128), device=dev)
        if args.workload != "cfg2":
            want = oracle.rs_encode_uniform(D, P, data[idx].cpu().numpy(), DATA_LEN)
            if p2p is not None:
                # read my groups' shards back from wherever the kernel put them (peer HBM through the IPC mapping)
                got = np.stack([p2p["peer_log"][sharding.replica_rank(rank, r, world)].tensor().view(R, n, ds)[r][idx].cpu().numpy()
                                for r in range(D, D + P)])
                dshard = p2p["peer_log"][sharding.replica_rank(rank, 1, world)].tensor().view(R, n, ds)[1][idx].cpu().numpy()
                assert (dshard[:, :L] == data[idx][:, L:2 * L].cpu().numpy()).all(), "bench data-shard check failed"
            else:
                got = parity[:, idx].cpu().numpy()
            assert (got == want).all(), "bench parity check failed"
        src = (p2p["acks_t"] if p2p is not None else (ack_recv if world > 1 else planes))
        if args.no_tally:
            committed.zero_()
            ctx.tally_planes(src, THRESH_RSPAXOS if args.workload != "cfg2" else THRESH_MULTIPAXOS, True, committed, bar)
        cw, bw = oracle.tally_planes(src[:, idx].cpu().numpy().view(np.uint64),
                                     THRESH_RSPAXOS if args.workload != "cfg2" else THRESH_MULTIPAXOS)
        assert (committed[idx].cpu().numpy().view(np.uint64) == cw).all(), "bench commit check failed"
        check_note = f"{len(idx)} sampled groups bit-exact vs oracle"

    alg_rs = (D + P) * L                       # 6830 B / codeword
    alg_tally = (R + 1) * 8 + 4                # 48 B planes+commit word, +4 B commit_bar
    if args.workload == "cfg2":
        alg = alg_tally
        value = n * world * 64 / (ms_per_step * 1e-3)
        unit = "slots/s"
    else:
        alg = alg_rs + alg_tally
        value = alg_rs * n * world / (ms_per_step * 1e-3) / 1e9
        unit = "GB/s"
    achieved = alg * n / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": profile_traffic(args.workload if R == 5 else f"{args.workload}_r{R}"),
                "kernel": rs.last_kernel() if args.workload != "cfg2" else "tally_planes_kernel",
                "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": alg * n, "peak_source": peak_src + " (burst figure; kernel timed alone)"}
    if world > 1 and args.workload == "cfg3":
        # the same fused kernel WITHOUT the replicate stores (parity to local HBM only): per-GPU compute is flat in N
        lp = torch.empty((P, n, ds), dtype=torch.uint8, device=dev)
        src_planes = p2p["acks_t"] if p2p is not None else ack_recv
        local_ms = _time_steps(torch, lambda: check(ctx.lib.ss_accept_step_fused_dev(
            rs.h, data.data_ptr(), DATA_LEN, DATA_LEN, n, lp.data_ptr(), ps, ds, SS_RS_OUT_PADDED16, src_planes.data_ptr(), R,
            THRESH_RSPAXOS, committed.data_ptr(), bar.data_ptr())), 10, 3)
        del lp
        roofline["local_only_kernel_ms"] = local_ms
        roofline["local_only_frac"] = alg * n / (local_ms * 1e-3) / 1e9 / peak
        remote = sum(1 for r in range(R) if sharding.replica_rank(rank, r, world) != rank)
        nv_bytes = remote * n * L
        nv_ms = nv_bytes / 770e9 * 1e3
        roofline["comm"] = {"exchange": args.exchange if p2p is not None or args.exchange == "nccl" else "nccl",
                            "remote_planes_per_rank": remote, "nvlink_bytes_per_rank_per_step": nv_bytes,
                            "nvlink_ref_gbs": 770.0, "nvlink_bound_ms": nv_ms, "hbm_bound_ms": alg * n / (peak * 1e9) * 1e3,
                            "step_ms": ms_per_step, "frac_of_slower_bound": max(nv_ms, alg * n / (peak * 1e9) * 1e3) / ms_per_step,
                            "note": "target time = slower of HBM bytes / measured copy bandwidth and NVLink bytes / 770 GB/s "
                                    "(measured peer-copy reference, B200_PROFILING.md); remote shards are written by the encode "
                                    "kernel itself into peer HBM" if p2p is not None else "NCCL all-to-all baseline"}
    if rank == 0:
        here = copy_bandwidth_here(torch, dev)
        if here:
            roofline["copy_gbs_this_box"] = here
            roofline["frac_of_copy_this_box"] = achieved / here

    extra = {}
    if rank == 0 and args.workload == "cfg3" and world == 1:
        # cfg2 (MultiPaxos tally only) measured beside it; 4 rotated plane sets > L2
        extra["cfg2"] = bench_cfg2(ctx, torch, dev, n, peak)

    e2e = None
    if not args.no_e2e and world == 1:
        e2e = bench_e2e(ctx, rs, torch, n, args)
    elif world > 1 and not args.no_e2e:
        e2e = bench_e2e(ctx, rs, torch, n, args, dist=dist, world=world, dev=dev)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_arm(3, 1, 1 << 18, args.workload)
        cpu = {"value": r["gbs"] if args.workload != "cfg2" else r["slots_per_s"], "unit": unit, "cores": r["threads"],
               "kind": "port", "sample": r["sample"] + " x 3 steps", "path": r["path"], "slots_per_s": r["slots_per_s"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": config_dict(args, world),
            "slots_committed_per_s": n * world * 64 / (ms_per_step * 1e-3),
            "payload_GBps": DATA_LEN * n * world / (ms_per_step * 1e-3) / 1e9,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "parity_check": check_note,
        }
        if world > 1 and args.workload == "cfg3":
            line["scaling_note"] = ("N=1 simulates all 5 replicas of a group on one GPU (no exchange: the shard planes are the "
                                    "follower logs). At N>1 replica r of a group led from rank h lives on rank (h+r)%N and the encode "
                                    "kernel stores its shard there over NVLink, so the step becomes NVLink-bound (roofline.comm); the "
                                    "same kernel without the remote stores takes roofline.local_only_kernel_ms on every N.")
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_cfg2(ctx, torch, dev, n, peak):
    sets = 4
    gen = torch.Generator(device=dev); gen.manual_seed(1234)
    planes = torch.randint(-(1 << 62), 1 << 62, (sets, R, n), dtype=torch.int64, device=dev, generator=gen)
    planes[:, 0] = -1
    committed = torch.empty((sets, n), dtype=torch.int64, device=dev)
    bar = torch.empty((sets, n), dtype=torch.int32, device=dev)
    for i in range(8):
        ctx.tally_planes(planes[i % sets], THRESH_MULTIPAXOS, True, committed[i % sets], bar[i % sets])
    torch.cuda.synchronize()
    iters = 40
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        ctx.tally_planes(planes[i % sets], THRESH_MULTIPAXOS, True, committed[i % sets], bar[i % sets])
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    alg = ((R + 1) * 8 + 4) * n
    return {"workload": "cfg2: MultiPaxos quorum tally 3 of 5, 2^20 groups x 64 slots, 4 rotated plane sets (208 MB > L2)",
            "slots_committed_per_s": n * 64 / (ms * 1e-3), "ms_per_step": ms,
            "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                         "note": "52 B/group; ~8 us of work per launch, launch-latency bound"}}


def bench_e2e(ctx, rs, torch, n, args, dist=None, world=1, dev=None):
    """Same step through the host-buffer C-ABI entry points: pinned host inputs, H2D + kernel + D2H inside."""
    L, ds, ps = rs.parity_layout(DATA_LEN, n)
    if args.workload == "cfg2":
        planes = torch.empty((R, n), dtype=torch.int64, pin_memory=True)
        planes.random_(-(1 << 62), 1 << 62)
        planes[0] = -1
        pn = planes.numpy().view(np.uint64)
        steps = max(3, min(args.steps, 10))
        ctx.tally_planes_host(pn, THRESH_MULTIPAXOS)
        t0 = time.perf_counter()
        for _ in range(steps):
            ctx.tally_planes_host(pn, THRESH_MULTIPAXOS)
        dt = (time.perf_counter() - t0) / steps
        val = torch.tensor([dt], dtype=torch.float64)
        return {"value": n * world * 64 / dt, "unit": "slots/s", "h2d_bytes_per_step": R * n * 8,
                "d2h_bytes_per_step": n * 12, "steps": steps, "ms_per_step": dt * 1e3}
    hdata = torch.empty((n, DATA_LEN), dtype=torch.uint8, pin_memory=True)
    rng = np.random.Generator(np.random.Philox(key=99))
    hv = hdata.numpy()
    chunk = 1 << 16
    for a in range(0, n, chunk):
        hv[a:a + chunk] = rng.integers(0, 256, size=(min(chunk, n - a), DATA_LEN), dtype=np.uint8)
    hpar = torch.empty((P, n, ds), dtype=torch.uint8, pin_memory=True)
    hplanes = torch.empty((R, n), dtype=torch.int64, pin_memory=True)
    hplanes.random_(-(1 << 62), 1 << 62)
    hplanes[0] = -1
    pn = hplanes.numpy().view(np.uint64)
    steps = max(3, min(args.steps, 10))

    def step():
        rs.encode_uniform_host(hv, DATA_LEN, hpar.numpy())
        return ctx.tally_planes_host(pn, THRESH_RSPAXOS)

    step()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        c, b = step()
    dt = (time.perf_counter() - t0) / steps
    if dist is not None:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t[0])
    # check the e2e result too
    from oracle import pyoracle as oracle
    idx = np.arange(0, n, max(1, n // 64))
    want = oracle.rs_encode_uniform(D, P, hv[idx], DATA_LEN)
    assert (hpar.numpy()[:, idx] == want).all(), "e2e parity check failed"
    return {"value": (D + P) * L * n * world / dt / 1e9, "unit": "GB/s",
            "h2d_bytes_per_step": n * DATA_LEN + R * n * 8, "d2h_bytes_per_step": P * n * ds + n * 12,
            "steps": steps, "ms_per_step": dt * 1e3, "slots_committed_per_s": n * world * 64 / dt,
            "timing": "host wall clock around blocking C-ABI calls (they return when results are in host memory)"}


def run_cfg5(args, ctx, dev, world, rank, peak, peak_src):
    import torch
    import torch.distributed as dist
    from summerset_b200 import workloads as wl
    G = 1 << 22
    w = wl.cfg5_raft(1 << 16, 7, 64, seed_extra=rank)
    rep = G // (1 << 16)
    t = lambda a, r: torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).to(dev).repeat(*r)
    match = t(w["match"], (1, rep)); lc = t(w["last_commit"], (rep,)); le = t(w["log_end"], (rep,))
    ct = t(w["curr_term"], (rep,)); terms = t(w["terms"], (rep, 1))
    out = torch.empty(G, dtype=torch.int32, device=dev)
    for _ in range(max(3, args.warmup)):
        ctx.raft_commit_scan(match, lc, le, ct, terms, 4, out)
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    l0 = ctx.launches
    a.record()
    for _ in range(args.steps):
        ctx.raft_commit_scan(match, lc, le, ct, terms, 4, out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt[0])
    alg = 296 * G
    if rank == 0:
        print(json.dumps({"metric": "Raft groups scanned/s (7 replicas, 64-slot window)", "value": G * world / (ms * 1e-3),
                          "unit": "groups/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                          "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "u32", "data": "synthetic", "config": config_dict(args, world),
                          "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src},
                          "gpu_launches": int(ctx.launches - l0), "e2e": None, "cpu_baseline": None}))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def _time_steps(torch, fn, steps, warmup):
    for _ in range(max(3, warmup)):
        fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def run_cfg3b(args, ctx, rs, dev, world, rank, peak, peak_src):
    """cfg 3b: batched reconstruct_data of 2^20 RS(3,2) codewords x 4 KB under the seeded erasure mix."""
    import torch
    from summerset_b200 import workloads as wl
    n = args.groups
    L, ds, ps = rs.parity_layout(DATA_LEN, n)
    gen = torch.Generator(device=dev); gen.manual_seed(wl.SEED_BASE + 3 + rank)
    data = torch.randint(0, 256, (n, DATA_LEN), dtype=torch.uint8, device=dev, generator=gen)
    sh = torch.zeros((D + P, n, ds), dtype=torch.uint8, device=dev)
    padded = torch.zeros((n, D * L), dtype=torch.uint8, device=dev)
    padded[:, :DATA_LEN] = data
    for i in range(D):
        sh[i, :, :L] = padded[:, i * L:(i + 1) * L]
    del padded
    rs.encode_uniform(data, DATA_LEN, parity=sh[D:])
    keep = sh.clone()
    present = wl.erasure_patterns(n, D, P, seed_extra=rank)
    pm = torch.from_numpy(present.astype(np.int32)).to(dev)
    for j in range(D + P):
        sh[j][((pm >> j) & 1) == 0] = 0x5A
    off = torch.arange(n, dtype=torch.int64, device=dev) * ds
    lens = torch.full((n,), DATA_LEN, dtype=torch.int32, device=dev)
    l0 = ctx.launches
    ms = _time_steps(torch, lambda: rs.reconstruct_batch(sh, n * ds, off, lens, pm, True), args.steps, args.warmup)
    launches = (ctx.launches - l0)
    ok = all(torch.equal(sh[i], keep[i]) for i in range(D))
    miss_data = sum(((present >> i) & 1) == 0 for i in range(D)).astype(np.int64)
    alg = int(((miss_data > 0) * D * L + miss_data * L).sum()) + n * 20
    if rank == 0:
        print(json.dumps({"metric": "RS reconstruct GB/s (reconstruct_data, RS(3,2), 4 KB payloads, cfg-3b erasure mix)",
                          "value": alg / (ms * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
                          "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                          "config": {"workload": "cfg3b: reconstruct_data, 2^20 codewords, 50% intact / 25% one data shard / 25% two shards missing"},
                          "codewords_per_s": n / (ms * 1e-3), "roundtrip_bit_exact": bool(ok), "kernel": rs.last_kernel(),
                          "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                                       "algorithmic_bytes_per_launch": alg, "peak_source": peak_src},
                          "gpu_launches": int(launches * args.steps // max(1, args.steps + max(3, args.warmup))), "e2e": None, "cpu_baseline": None}))


def run_cfg4(args, ctx, rs, dev, world, rank, peak, peak_src):
    """cfg 4: Crossword n=5,d=3,T=5,f=2: ragged RS(3,2) encode of mixed 256 B..64 KB payloads + coverage tally."""
    import torch
    from oracle import pyoracle as oracle
    from summerset_b200 import workloads as wl
    n = args.groups
    lens, spr = wl.cfg4_lengths(n, seed_extra=0)      # same sizes on every rank (payload bytes differ by rank)
    lay = wl.ragged_layout(lens, D)
    gen = torch.Generator(device=dev); gen.manual_seed(wl.SEED_BASE + 4 + rank)
    arena = torch.randint(0, 256, (lay["data_bytes"] + 256,), dtype=torch.uint8, device=dev, generator=gen)
    parity = torch.empty((P, lay["plane_bytes"]), dtype=torch.uint8, device=dev)
    doff = torch.from_numpy(lay["data_off"].astype(np.int64)).to(dev)
    poff = torch.from_numpy(lay["par_off"].astype(np.int64)).to(dev)
    dlen = torch.from_numpy(lens.astype(np.int32)).to(dev)
    masks = torch.randint(0, 32, (n,), dtype=torch.uint8, device=dev, generator=gen) | 1     # leader always acks
    pidx = torch.from_numpy((spr - 1).astype(np.uint8)).to(dev)
    from summerset_b200.api import crossword_brr_assignment
    policies = [crossword_brr_assignment(5, 5, s) for s in (1, 2, 3)]

    # replica logs for the distribute step: replica r of my groups lives on rank (rank + r) % world
    import torch.distributed as dist
    from summerset_b200 import sharding
    Lp = (lay["L"].astype(np.int64) + 15) // 16 * 16
    slot_bytes = spr.astype(np.int64) * Lp
    rep_off_np = np.concatenate([[0], np.cumsum(slot_bytes)[:-1]]).astype(np.int64)
    region = int(slot_bytes.sum() + 255) // 256 * 256
    log = ctx.dev_alloc(5 * region)
    peer_log = {rank: log}
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, ctx.ipc_export(log))
        for q in range(world):
            if q != rank:
                peer_log[q] = ctx.ipc_open(handles[q], 5 * region)
        tiny = torch.zeros(1, dtype=torch.int32, device=dev)
    rep_ptrs = [peer_log[sharding.replica_rank(rank, r, world)].ptr + r * region for r in range(5)]
    rep_off = torch.from_numpy(rep_off_np).to(dev)
    spr_t = torch.from_numpy(spr).to(dev)

    def step():
        # lens were drawn with the same seed on every rank, so all regions have the same size
        rs.crossword_distribute(arena, doff, dlen, spr_t, rep_off, rep_ptrs)
        out = ctx.tally_crossword(masks, pidx, policies, 5, 3, 3, 2, True)
        if world > 1:
            dist.all_reduce(tiny)
        return out

    l0 = ctx.launches
    ms = _time_steps(torch, step, args.steps, args.warmup)
    launches = ctx.launches - l0
    dist_kernel = rs.last_kernel()
    ms_enc = _time_steps(torch, lambda: rs.encode_batch(arena, doff, dlen, parity, lay["plane_bytes"], poff), args.steps, 1)
    if world > 1:
        tt = torch.tensor([ms, ms_enc], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_enc = float(tt[0]), float(tt[1])
    # the follower view: replica 1 of my groups (wherever it lives) holds shard (1 + k) % 5 in slot k
    chk = peer_log[sharding.replica_rank(rank, 1, world)].tensor()[region:2 * region]
    # spot check vs oracle
    idx = np.arange(0, n, max(1, n // 64))
    sub_len = lens[idx]; sub_lay = wl.ragged_layout(sub_len, D)
    sub = np.zeros(sub_lay["data_bytes"] + 64, dtype=np.uint8)
    for j, g in enumerate(idx):          # gather the sampled payloads / parities on the device, copy only those
        o = int(lay["data_off"][g]); so = int(sub_lay["data_off"][j]); ln = int(lens[g])
        sub[so:so + ln] = arena[o:o + ln].cpu().numpy()
    want = np.zeros((P, sub_lay["plane_bytes"]), dtype=np.uint8)
    oracle.rs_encode_batch(D, P, sub, sub_lay["data_off"], sub_len, want.reshape(-1), sub_lay["plane_bytes"], sub_lay["par_off"])
    for j, g in enumerate(idx):
        Lg = int(lay["L"][g]); o = int(lay["par_off"][g]); so = int(sub_lay["par_off"][j])
        got = parity[:, o:o + Lg].cpu().numpy()
        assert (got == want[:, so:so + Lg]).all(), "cfg4 parity check failed"
        if spr[g] == 3:     # replica 1, slot 2 = shard 3 = parity 0
            o2 = int(rep_off_np[g]) + 2 * int(Lp[g])
            assert (chk[o2:o2 + Lg].cpu().numpy() == want[0, so:so + Lg]).all(), "cfg4 distribute check failed"
    note = f"{len(idx)} sampled codewords bit-exact vs oracle (parity planes and replica-1 log)"
    alg = int((lay["L"].astype(np.int64) * (D + P)).sum()) + n * 22
    alg_dist = int((lay["L"].astype(np.int64) * (D + 5 * spr.astype(np.int64))).sum()) + n * 31
    remote = sum(1 for r in range(5) if sharding.replica_rank(rank, r, world) != rank)
    nv_bytes = int((lay["L"].astype(np.int64) * spr.astype(np.int64)).sum()) * remote
    if rank == 0:
        print(json.dumps({"metric": "RS shard GB/s, Crossword ragged encode + distribute-by-assignment + coverage tally",
                          "value": alg * world / (ms * 1e-3) / 1e9,
                          "distribute": {"kernel": dist_kernel, "step_ms": ms, "hbm_bytes_per_rank": alg_dist,
                                         "hbm_GBps_per_rank": alg_dist / (ms * 1e-3) / 1e9,
                                         "nvlink_bytes_per_rank": nv_bytes, "nvlink_bound_ms": nv_bytes / 770e9 * 1e3,
                                         "hbm_bound_ms": alg_dist / (peak * 1e9) * 1e3,
                                         "frac_of_slower_bound": max(nv_bytes / 770e9 * 1e3, alg_dist / (peak * 1e9) * 1e3) / ms},
                          "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                          "config": {"workload": "cfg4: Crossword n=5 d=3 T=5 f=2, 2^20 codewords, data_len uniform over {256..65536}, spr uniform {1,2,3}",
                                     "payload_bytes": int(lens.astype(np.int64).sum())},
                          "slots_committed_per_s": n * world / (ms * 1e-3), "encode_only_ms": ms_enc, "parity_check": note,
                          "kernel": "rs32_encode_ragged_kernel",
                          "roofline": {"bound": "hbm", "achieved": alg / (ms_enc * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": alg / (ms_enc * 1e-3) / 1e9 / peak, "traffic": None,
                                       "algorithmic_bytes_per_launch": alg, "peak_source": peak_src},
                          "gpu_launches": int(launches), "e2e": None, "cpu_baseline": None}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    global D, P, R, THRESH_RSPAXOS, THRESH_MULTIPAXOS
    args = parse_args()
    # The contract is ONE JSON line on stdout.  Native libraries print there too (NCCL writes "NCCL version ..." to fd 1
    # when the first communicator comes up), so fd 1 is pointed at stderr for the whole run and Python's sys.stdout
    # keeps a private duplicate of the real stdout for the result line.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if args.replicas != 5:      # scripts/local_cluster.py:41-50 defaults: fault_tolerance = (n//2)//2
        R = args.replicas
        D = R // 2 + 1
        P = R - D
        THRESH_MULTIPAXOS = D
        THRESH_RSPAXOS = D + (R // 2) // 2
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
