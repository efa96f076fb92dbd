"""Two-rank worker of tests/test_gpu_replicate.py::test_two_process_cuda_ipc_replicate (one process per GPU).

Runs a few steps of the multi-GPU accept step (summerset_b200.replicate) with shard planes and ack planes crossing
NVLink through CUDA-IPC mappings (ss_ipc_export / ss_ipc_open) and ordering carried only by the step flags, then
checks on EVERY rank: all five shard planes of all of its groups (read back from the peer GPU's HBM) and the commit
words / commit_bar against the oracle.  Also exercises ss_copy_d2d into peer memory.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import pyoracle as oracle  # noqa: E402  (the checker)
from summerset_b200 import workloads as wl  # noqa: E402
from summerset_b200.api import Context, ReedSolomon  # noqa: E402
from summerset_b200.replicate import ReplicatedAcceptStep  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("gloo")

    def exchange(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    ctx = Context(local)
    rs = ReedSolomon(ctx, 3, 2)
    n, data_len, R, steps = 4098, 4096, 5, 5
    for lag, mode in ((1, "p2p"), (2, "p2p"), (2, "ce")):
        st = ReplicatedAcceptStep(ctx, rs, n, data_len, R, world, rank, exchange, lag=lag, mode=mode)
        assert st.remote_planes() == (2 if world == 2 else 0)
        empty = np.zeros((R, n), dtype=np.uint64)
        st.fill_acks(torch.from_numpy(empty.view(np.int64)).to(dev))
        dist.barrier()
        committed = torch.zeros(n, dtype=torch.int64, device=dev)
        bar = torch.zeros(n, dtype=torch.int32, device=dev)
        # the acks my followers send at step k are seeded by (the LEADER's rank, step): plane r belongs to leader_of[r]
        def follower_acks(k):
            pl = np.zeros((R, n), dtype=np.uint64)
            for r in range(R):
                pl[r] = wl.cfg2_planes(n, R, 0.8, seed_extra=1000 * st.leader_of[r] + k)[r]
            return pl
        data = None
        keep = []          # the copy-engine mode reads the ack planes on a second stream: the tensors must outlive the call
        for k in range(1, steps + 1):
            data = wl.payload_uniform(n, data_len, seed_extra=10 * rank + k)
            keep.append((torch.from_numpy(data).to(dev), torch.from_numpy(follower_acks(k).view(np.int64)).to(dev)))
            st.step(keep[-1][0], keep[-1][1], 4, committed, bar)
            # NO synchronisation between steps: the flags order everything
        st.drain()
        torch.cuda.synchronize()
        dist.barrier()
        assert ctx.device_status() == 0, "a flag wait timed out"
        # tally of the last step = acks my followers sent at step (steps - lag), seeded by MY rank
        src = wl.cfg2_planes(n, R, 0.8, seed_extra=1000 * rank + (steps - lag))
        cw, bw = oracle.tally_planes(src, 4)
        assert (committed.cpu().numpy().view(np.uint64) == cw).all(), f"rank {rank} lag {lag} {mode}: commit words"
        assert (bar.cpu().numpy().view(np.uint32) == bw).all()
        # every plane of every group of the last step, wherever it lives
        L = st.L
        want_par = oracle.rs_encode_uniform(3, 2, data, data_len)
        for r, plane in enumerate(st.my_shards()):
            got = plane.cpu().numpy()
            if r < 3:
                want = np.stack([oracle.cw_split(data[g, :data_len].tobytes(), 3)[r] for g in range(n)])
            else:
                want = want_par[r - 3][:, :L]
            assert (got[:, :L] == want).all(), f"rank {rank} lag {lag} {mode}: plane {r}"
            assert (got[:, L:] == 0).all()
        # ss_copy_d2d into peer memory: overwrite the peer's ack buffer 0 plane 0 and read it back
        peer = (rank + 1) % world
        if peer != rank:
            mark = torch.full((n,), 0x1234 + rank, dtype=torch.int64, device=dev)
            ctx.copy_d2d(st.peer_acks[peer].ptr, mark.data_ptr(), n * 8)
            torch.cuda.synchronize()
            dist.barrier()
            mine = st.acks_view[0, 0]
            assert int(mine[0]) == 0x1234 + peer and int(mine[-1]) == 0x1234 + peer
        dist.barrier()
        st.close()
    print("IPC_REPLICATE_OK", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
