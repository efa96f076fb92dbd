"""The multi-GPU accept step -- ss_accept_step_replicate_dev (five-pointer plane mode of the row kernel),
ss_follower_ack_dev, the step flags and the CUDA-IPC path -- against the oracle.

One-GPU tests use five distinct LOCAL buffers as the five replicas' logs (the kernel cannot tell local from peer
memory); the two-process test maps real peer memory through ss_ipc_export/open and needs two GPUs (skips otherwise).
Reference: rspaxos/request.rs:72-77,127-142 (encode + shard r to peer r), rspaxos/durability.rs:101-118 (follower
ack), rspaxos/messages.rs:438-440 (tally).
"""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from summerset_b200 import workloads as wl

DEV = "cuda:0"
ROOT = Path(__file__).resolve().parent.parent


def _want_planes(oracle, data, data_len, d, p):
    """all d+p shard planes [d+p, n, L] of uniform codewords, from the oracle"""
    n = data.shape[0]
    L = oracle.cw_shard_len(data_len, d)
    out = np.zeros((d + p, n, L), dtype=np.uint8)
    for g in range(n):
        out[:d, g] = oracle.cw_split(data[g, :data_len].tobytes(), d)
    par = oracle.rs_encode_uniform(d, p, data, data_len)
    out[d:] = par[:, :, :L]
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("d,p", [(3, 2), (4, 3), (2, 1), (6, 4)])
@pytest.mark.parametrize("data_len", [1, 17, 4096, 12288, 30000])
@pytest.mark.parametrize("with_tally", [False, True])
def test_replicate_separate_local_buffers(ctx, oracle, d, p, data_len, with_tally):
    """d+p distinct buffers as the replicas' logs: every plane of every group, guard bands, padding, fused tally.  RS(3,2)
    takes its hand-specialised row kernel (segmented for payloads above 12 KB); the other codes the generalised row /
    packed kernels (compile-time cluster codes, or specialised by NVRTC), which also walk codewords wider than one CTA."""
    from summerset_b200._lib import SummersetError
    from summerset_b200.api import ReedSolomon, round_up, shard_len
    rs = ReedSolomon(ctx, d, p)
    n = 1531
    T = d + p
    stride = round_up(max(data_len, 16), 16)
    data = wl.payload_uniform(n, stride, seed_extra=data_len)
    L = shard_len(data_len, d); ds = round_up(L, 16)
    # separate allocations, each with a guard band that must stay untouched
    bufs = [torch.full((n * ds + 64,), 0xC3, dtype=torch.uint8, device=DEV) for _ in range(T)]
    ptrs = [b.data_ptr() + 32 for b in bufs]
    assert all(q % 16 == 0 for q in ptrs)
    planes = wl.cfg2_planes(n, 5, 0.9, seed_extra=3)
    pl = torch.from_numpy(planes.view(np.int64)).to(DEV)
    committed = torch.zeros(n, dtype=torch.int64, device=DEV)
    bar = torch.zeros(n, dtype=torch.int32, device=DEV)
    rs.accept_step_replicate(torch.from_numpy(data).to(DEV), data_len, ptrs, ds, pl if with_tally else None, 4,
                             committed if with_tally else None, bar if with_tally else None)
    torch.cuda.synchronize()
    assert rs.last_kernel().startswith("rs32_encode_row_kernel" if (d, p) == (3, 2) else "horner_encode_")
    want = _want_planes(oracle, data, data_len, d, p)
    for j in range(T):
        got = bufs[j].cpu().numpy()
        assert (got[:32] == 0xC3).all() and (got[32 + n * ds:] == 0xC3).all(), f"plane {j}: guard band written"
        body = got[32:32 + n * ds].reshape(n, ds)
        assert (body[:, :L] == want[j]).all(), f"plane {j} differs from the oracle"
        assert (body[:, L:] == 0).all(), f"plane {j}: padding not zero"
    if with_tally:
        cw, bw = oracle.tally_planes(planes, 4)
        assert (committed.cpu().numpy().view(np.uint64) == cw).all()
        assert (bar.cpu().numpy().view(np.uint32) == bw).all()


@pytest.mark.gpu
def test_replicate_rejects_what_it_does_not_support(ctx):
    from summerset_b200._lib import SS_ERR_UNSUPPORTED, SummersetError
    from summerset_b200.api import ReedSolomon
    d = torch.zeros((4, 4096), dtype=torch.uint8, device=DEV)
    out = torch.zeros(5 * 4 * 1376, dtype=torch.uint8, device=DEV)
    rs = ReedSolomon(ctx, 3, 2)
    with pytest.raises(SummersetError) as e:       # misaligned plane pointer
        rs.accept_step_replicate(d, 4096, [out.data_ptr() + 1 + j * 4 * 1376 for j in range(5)], 1376, None, 4, None, None)
    assert e.value.code < 0


@pytest.mark.gpu
def test_follower_ack_and_flags_single_gpu(ctx):
    """ss_follower_ack_dev copies every ack plane to its destination, skips NULL destinations, honours the wait flags
    (already satisfied here) and raises the signal flags afterwards."""
    from summerset_b200.api import StepSync
    R, n = 5, 4098
    src = torch.randint(-(1 << 62), 1 << 62, (R, n), dtype=torch.int64, device=DEV)
    dst = torch.zeros((R, n), dtype=torch.int64, device=DEV)
    flags = torch.zeros(32, dtype=torch.int64, device=DEV)
    flags[:R] = 7                                                  # wait flags already at step 7
    ptrs = [dst[r].data_ptr() if r != 2 else 0 for r in range(R)]
    sync = StepSync(flags.data_ptr(), R, 7, [flags[16 + r].data_ptr() for r in range(R)], 7)
    ctx.follower_ack(src, ptrs, sync)
    torch.cuda.synchronize()
    for r in range(R):
        if r == 2:
            assert int(dst[r].abs().sum()) == 0
        else:
            assert torch.equal(dst[r], src[r])
    assert flags[16:16 + R].tolist() == [7] * R
    assert ctx.device_status() == 0


@pytest.mark.gpu
def test_flag_wait_times_out_instead_of_hanging(ctx):
    """A wait that can never be satisfied ends after the bounded spin and reports it through the status word."""
    from summerset_b200.api import StepSync
    R, n = 3, 64
    src = torch.ones((R, n), dtype=torch.int64, device=DEV)
    dst = torch.zeros((R, n), dtype=torch.int64, device=DEV)
    flags = torch.zeros(8, dtype=torch.int64, device=DEV)
    ctx.follower_ack(src, [dst[r].data_ptr() for r in range(R)], StepSync(flags.data_ptr(), R, 5, [], 0))
    torch.cuda.synchronize()
    assert ctx.device_status() == 1          # SS_DEV_STATUS_FLAG_TIMEOUT, cleared by the read
    assert ctx.device_status() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("lag,mode", [(1, "p2p"), (2, "p2p"), (2, "ce")])
def test_replicated_step_protocol_single_gpu(ctx, oracle, lag, mode):
    """The whole E_k / A_k flag protocol with world = 1 (every replica hosted locally): shards of the last step in
    every plane, tally of step k == oracle tally of the acks delivered at step k - lag."""
    from summerset_b200.api import ReedSolomon
    from summerset_b200.replicate import ReplicatedAcceptStep
    n, data_len, R = 2050, 4096, 5
    rs = ReedSolomon(ctx, 3, 2)
    st = ReplicatedAcceptStep(ctx, rs, n, data_len, R, 1, 0, lambda o: [o], lag=lag, mode=mode)
    committed = torch.zeros(n, dtype=torch.int64, device=DEV)
    bar = torch.zeros(n, dtype=torch.int32, device=DEV)
    empty = np.zeros((R, n), dtype=np.uint64)
    st.fill_acks(torch.from_numpy(empty.view(np.int64)).to(DEV))
    datas = [wl.payload_uniform(n, data_len, seed_extra=100 + k) for k in range(4)]
    ackss = [wl.cfg2_planes(n, R, 0.8, seed_extra=200 + k) for k in range(4)]
    keep = []              # the copy-engine mode reads the ack planes on a second stream: keep the tensors alive
    for k in range(1, 5):
        keep.append((torch.from_numpy(datas[k - 1]).to(DEV), torch.from_numpy(ackss[k - 1].view(np.int64)).to(DEV)))
        st.step(keep[-1][0], keep[-1][1], 4, committed, bar)
        st.drain()
        torch.cuda.synchronize()
        src = ackss[k - lag - 1] if k - lag >= 1 else empty
        cw, bw = oracle.tally_planes(src, 4)
        assert (committed.cpu().numpy().view(np.uint64) == cw).all(), f"step {k}"
        assert (bar.cpu().numpy().view(np.uint32) == bw).all()
        want = _want_planes(oracle, datas[k - 1], data_len, 3, 2)
        for r, plane in enumerate(st.my_shards()):
            assert (plane.cpu().numpy()[:, :st.L] == want[r]).all(), f"step {k} plane {r}"
    assert ctx.device_status() == 0
    flags = st.flags.tensor().view(torch.int64)
    assert flags[:R].tolist() == [4] * R and flags[16:16 + R].tolist() == [4] * R
    st.close()


@pytest.mark.gpu
def test_two_process_cuda_ipc_replicate():
    """Two processes, two GPUs: shard planes and ack planes cross NVLink through CUDA-IPC mappings, ordered only by
    the step flags.  Each rank checks every plane of every group of its own groups (read back from the peer's HBM)
    and its commit words against the oracle."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ)
    env["MASTER_ADDR"] = "127.0.0.1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(ROOT / "tests" / "ipc_replicate_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("IPC_REPLICATE_OK") == 2, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("n,data_len,stride", [(5000, 4096, 4096), (4099, 1000, 1000), (3, 18, 32), (70001, 256, 256)])
def test_accept_step_fused_host_buffers(ctx, oracle, n, data_len, stride, monkeypatch):
    """ss_accept_step_fused: HOST payloads + ack planes in, parity + commit words + commit_bar out, chunked through
    the fused kernel (several chunks with a ragged last one: 1 MiB chunks via SS_E2E_CHUNK_MB)."""
    from summerset_b200.api import ReedSolomon, round_up, shard_len
    monkeypatch.setenv("SS_E2E_CHUNK_MB", "1")
    rs = ReedSolomon(ctx, 3, 2)
    data = wl.payload_uniform(n, data_len, stride=stride, seed_extra=n)
    planes = wl.cfg2_planes(n, 5, 0.85, seed_extra=n)
    L = shard_len(data_len, 3); ds = round_up(L, 16)
    parity = np.full((2, n, ds), 0xEE, dtype=np.uint8)
    committed = np.zeros(n, dtype=np.uint64)
    bar = np.zeros(n, dtype=np.uint32)
    rs.accept_step_fused_host(data, data_len, parity, planes, 4, committed, bar)
    want = oracle.rs_encode_uniform(3, 2, data, data_len)
    assert (parity == want).all()
    cw, bw = oracle.tally_planes(planes, 4)
    assert (committed == cw).all() and (bar == bw).all()
