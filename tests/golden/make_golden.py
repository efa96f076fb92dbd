"""Generates the committed golden fixtures from the CPU oracle.

The reference cannot be compiled or imported in this environment (Rust, no toolchain), so
these vectors come from oracle/ss_oracle.c -- itself pinned in tests/test_oracle_kat.py against the
reference's own tests and the upstream crate's known answers.  Run: python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import pyoracle as oracle  # noqa: E402
from summerset_b200 import workloads as wl  # noqa: E402

out = Path(__file__).resolve().parent
arrs = {}
for d, p, dl, n in [(3, 2, 4096, 5), (3, 2, 18, 3), (3, 2, 1, 2), (4, 3, 1000, 4), (5, 4, 257, 4), (6, 4, 100, 3),
                    # the cluster codes (population 3 / 7 / 9 / 6 / 4) at the benchmark payload size
                    (2, 1, 4096, 3), (4, 3, 4096, 3), (5, 4, 4096, 3), (4, 2, 4096, 2), (3, 1, 4096, 2)]:
    data = wl.payload_uniform(n, dl, seed_extra=1000 + d)
    arrs[f"data_{d}_{p}_{dl}"] = data
    arrs[f"parity_{d}_{p}_{dl}"] = oracle.rs_encode_uniform(d, p, data, dl)
np.savez_compressed(out / "rs_golden.npz", **arrs)

planes = wl.cfg2_planes(512, 5, 0.9, seed_extra=5)
c, bar = oracle.tally_planes(planes, 3)
w = wl.cfg5_raft(256, 7, 64, seed_extra=5)
nc = oracle.raft_scan_batch(w["match"], w["last_commit"], w["log_end"], w["curr_term"], w["terms"], 4)
np.savez_compressed(out / "tally_golden.npz", planes=planes, threshold=3, committed=c, commit_bar=bar,
                    raft_match=w["match"], raft_last_commit=w["last_commit"], raft_log_end=w["log_end"],
                    raft_curr_term=w["curr_term"], raft_terms=w["terms"], raft_threshold=4, raft_new_commit=nc)
print("wrote", [p.name for p in out.glob("*.npz")])
