"""GPU parity tests for the Reed-Solomon path: CUDA kernels (through the C ABI) vs the CPU oracle on the
same seeded inputs, bit-exact.  The first block restates src/utils/rscoding.rs:685-877 against the
GPU-backed coder; the rest covers geometry edge cases, every erasure pattern, ragged batches and
size-independent properties at BASELINE.json's full sizes."""
import itertools

import numpy as np
import pytest
import torch

from summerset_b200 import workloads as wl
from summerset_b200.api import Bitmap, ReedSolomon, RSCodeword, SummersetError, bincode_string, round_up, shard_len

pytestmark = pytest.mark.gpu
DATA = bincode_string("interesting_value")
DEV = "cuda:0"


# ---------------------------------------------------------------------------------------------
# the reference's own tests (rscoding.rs:685-877) on the GPU coder
# ---------------------------------------------------------------------------------------------
def test_ref_compute_verify(ctx):
    rs32 = ReedSolomon(ctx, 3, 2)
    cw_null = RSCodeword.from_null(3, 2)
    with pytest.raises(SummersetError):
        cw_null.compute_parity(rs32)
    with pytest.raises(SummersetError):
        cw_null.verify_parity(rs32)
    cw_part = RSCodeword.from_data(DATA, 3, 2)
    cw_part.shards[1] = None
    with pytest.raises(SummersetError):
        cw_part.compute_parity(rs32)
    with pytest.raises(SummersetError):
        cw_part.verify_parity(rs32)
    cw = RSCodeword.from_data(DATA, 3, 0)
    cw.compute_parity(None)
    assert cw.avail_parity_shards() == 0 and cw.verify_parity(None)
    cw = RSCodeword.from_data(DATA, 3, 2)
    cw.compute_parity(rs32)
    assert cw.avail_parity_shards() == 2
    assert cw.verify_parity(rs32)
    # derived KAT
    assert bytes(cw.shards[3]).hex() == "2b6c7b717e70" and bytes(cw.shards[4]).hex() == "2ffb9ccc5da8"
    rs53 = ReedSolomon(ctx, 5, 3)
    for fn in (cw.compute_parity, cw.verify_parity):
        with pytest.raises(SummersetError):
            fn(None)
        with pytest.raises(SummersetError):
            fn(rs53)
    # a corrupted parity byte must fail verification
    cw.shards[4][2] ^= 1
    assert not cw.verify_parity(rs32)


def test_ref_reconstruction(ctx):
    rs32 = ReedSolomon(ctx, 3, 2)
    cw_null = RSCodeword.from_null(3, 2)
    with pytest.raises(SummersetError):
        cw_null.reconstruct_all(rs32)
    with pytest.raises(SummersetError):
        cw_null.reconstruct_data(rs32)
    cw_part = RSCodeword.from_data(DATA, 3, 2)
    cw_part.shards[1] = None
    with pytest.raises(SummersetError):
        cw_part.reconstruct_all(rs32)       # only 2 of 5 present
    with pytest.raises(SummersetError):
        cw_part.reconstruct_data(rs32)
    cw = RSCodeword.from_data(DATA, 3, 2)
    cw.reconstruct_all(rs32)
    assert cw.avail_shards() == 5
    golden = [s.copy() for s in cw.shards]
    cw.shards[1] = None; cw.shards[3] = None
    cw.reconstruct_all(rs32)
    assert cw.avail_shards() == 5 and all((a == b).all() for a, b in zip(cw.shards, golden))
    cw.shards[0] = None; cw.shards[2] = None
    cw.reconstruct_data(rs32)
    assert cw.avail_data_shards() == 3 and all((a == b).all() for a, b in zip(cw.shards[:3], golden[:3]))
    cw.shards[0] = None; cw.shards[1] = None; cw.shards[4] = None
    with pytest.raises(SummersetError) as ei:
        cw.reconstruct_all(rs32)
    assert ei.value.code == -10                                   # TooFewShardsPresent
    with pytest.raises(SummersetError):
        cw.reconstruct_data(rs32)
    assert cw.shards[0] is None and cw.shards[1] is None          # never partial
    rs53 = ReedSolomon(ctx, 5, 3)
    with pytest.raises(SummersetError):
        cw.reconstruct_all(None)
    with pytest.raises(SummersetError):
        cw.reconstruct_all(rs53)


def test_ref_get_data(ctx):
    rs32 = ReedSolomon(ctx, 3, 2)
    cw = RSCodeword.from_data(DATA, 3, 2)
    assert cw.get_data() == DATA
    cw.compute_parity(rs32)
    cw.shards[0] = None
    cw.data_copy = None
    with pytest.raises(SummersetError):
        cw.get_data()
    cw.reconstruct_data(rs32)
    assert cw.get_data() == DATA


def test_coder_new_errors_and_matrix(ctx, oracle):
    for d, p, code in [(0, 1, -3), (3, 0, -5), (200, 57, -2)]:
        with pytest.raises(SummersetError) as ei:
            ReedSolomon(ctx, d, p)
        assert ei.value.code == code == oracle.rs_new_rc(d, p)
    for d, p in [(3, 2), (4, 2), (4, 3), (5, 4), (6, 4), (9, 6), (12, 8), (2, 1), (1, 1), (17, 3)]:
        rs = ReedSolomon(ctx, d, p)
        assert (rs.matrix() == oracle.rs_matrix(d, p)).all()
        assert (rs.data_shard_count(), rs.parity_shard_count(), rs.total_shard_count()) == (d, p, d + p)


def test_upstream_one_encode_on_gpu(ctx):
    rs = ReedSolomon(ctx, 5, 5)
    shards = [np.array(x, dtype=np.uint8) for x in ([0, 1], [4, 5], [2, 3], [6, 7], [8, 9])]
    shards += [np.zeros(2, dtype=np.uint8) for _ in range(5)]
    rs.encode(shards)
    assert [s.tolist() for s in shards[5:]] == [[12, 13], [10, 11], [14, 15], [90, 91], [94, 95]]
    assert rs.verify(shards)


def test_single_shape_errors(ctx):
    rs = ReedSolomon(ctx, 3, 2)
    sh = [np.zeros(8, np.uint8) for _ in range(5)]
    with pytest.raises(SummersetError) as ei:
        rs.encode(sh[:4])
    assert ei.value.code == -1
    with pytest.raises(SummersetError) as ei:
        rs.encode(sh + [np.zeros(8, np.uint8)])
    assert ei.value.code == -2
    with pytest.raises(SummersetError) as ei:
        rs.encode([np.zeros(0, np.uint8) for _ in range(5)])
    assert ei.value.code == -11
    with pytest.raises(SummersetError) as ei:
        rs.encode(sh[:4] + [np.zeros(7, np.uint8)])
    assert ei.value.code == -9


# ---------------------------------------------------------------------------------------------
# batched encode vs oracle
# ---------------------------------------------------------------------------------------------
def _gpu_encode_uniform(rs, data_np, data_len):
    data = torch.from_numpy(data_np).to(DEV)
    par = rs.encode_uniform(data, data_len)
    torch.cuda.synchronize()
    return par.cpu().numpy()


@pytest.mark.parametrize("d,p", [(3, 2), (4, 3), (5, 4), (2, 1), (6, 4), (9, 6), (1, 1), (12, 8), (17, 3)])
@pytest.mark.parametrize("data_len", [1, 2, 3, 15, 16, 17, 47, 48, 49, 255, 4096, 4097, 5000])
def test_encode_uniform_matches_oracle(ctx, oracle, d, p, data_len):
    rs = ReedSolomon(ctx, d, p)
    n = 37
    data = wl.payload_uniform(n, data_len, stride=round_up(data_len, 16) + 16, seed_extra=d * 1000 + p)
    got = _gpu_encode_uniform(rs, data, data_len)
    want = oracle.rs_encode_uniform(d, p, data, data_len)
    assert got.shape == want.shape
    assert (got == want).all()
    assert rs.last_kernel().startswith("rs32_" if (d, p) == (3, 2) else ("horner_" if d <= 8 else "generic_"))


@pytest.mark.parametrize("variant", [1, 2, 3, 5, 6, 8])
def test_encode_kernel_variants(ctx, oracle, variant):
    """every selectable kernel variant (flat v1, register budgets of the row kernel, bit-plane generic) is bit-exact"""
    for d, p, data_len, n in [(3, 2, 4096, 2500), (3, 2, 100, 700), (3, 2, 16 * 200, 33), (4, 3, 4096, 300), (5, 4, 1000, 300)]:
        rs = ReedSolomon(ctx, d, p)
        rs.set_variant(variant)
        data = wl.payload_uniform(n, data_len, seed_extra=variant)
        got = _gpu_encode_uniform(rs, data, data_len)
        assert (got == oracle.rs_encode_uniform(d, p, data, data_len)).all(), (variant, d, p, data_len, rs.last_kernel())
    # reconstruct with the bit-plane kernel (5) and the global-table Horner kernel (8) too
    if variant in (5, 8):
        d, p, data_len = 3, 2, 777
        rs = ReedSolomon(ctx, d, p); rs.set_variant(variant)
        n = 32
        data = wl.payload_uniform(n, data_len, seed_extra=1)
        full, L, ds = _planes_from(oracle, d, p, data, data_len)
        present = np.arange(n, dtype=np.uint32)
        sh = torch.from_numpy(full.copy()).to(DEV)
        off = torch.arange(n, dtype=torch.int64, device=DEV) * ds
        st = rs.reconstruct_batch(sh, n * ds, off, torch.full((n,), data_len, dtype=torch.int32, device=DEV),
                                  torch.from_numpy(present.astype(np.int32)).to(DEV), False)
        torch.cuda.synchronize()
        assert rs.last_kernel() == ("generic_reconstruct_kernel" if variant == 5 else "horner_reconstruct_kernel")
        assert (sh.cpu().numpy()[:, :, :L] == full[:, :, :L]).all()
        assert [int(x) for x in st.cpu()] == [0 if bin(pt).count("1") >= d else -10 for pt in range(n)]


def test_row_kernel_emit_data_and_wide_codewords(ctx, oracle):
    """SS_RS_EMIT_DATA (all d+p planes in one pass) on the row kernel and the general kernels; codewords wider
    than the row kernel's 256-column limit fall back to the flat kernel"""
    from summerset_b200._lib import check
    for d, p, data_len, n in [(3, 2, 4096, 1000), (3, 2, 20000, 64), (3, 2, 5, 40), (4, 3, 1000, 100), (4, 3, 4096, 300),
                              (4, 3, 40000, 37), (5, 4, 30001, 21), (6, 4, 4096, 100), (6, 4, 50000, 9), (2, 1, 9000, 50)]:
        rs = ReedSolomon(ctx, d, p)
        data = wl.payload_uniform(n, data_len, seed_extra=3)
        L, ds, ps = rs.parity_layout(data_len, n)
        sh = torch.full((d + p, n, ds), 0x11, dtype=torch.uint8, device=DEV)
        dt = torch.from_numpy(data).to(DEV)
        check(ctx.lib.ss_rs_encode_uniform_dev(rs.h, dt.data_ptr(), data.shape[1], data_len, n, sh[d:].data_ptr(), ps, ds, 3))
        torch.cuda.synchronize()
        full, L2, ds2 = _planes_from(oracle, d, p, data, data_len)
        assert (sh.cpu().numpy() == full).all(), (d, p, data_len, rs.last_kernel())


def test_encode_unaligned_stride_and_exact_output(ctx, oracle):
    """payload stride not a multiple of 16 (every codeword starts at a different alignment) and
    byte-exact (non padded) parity slots with stride L."""
    from summerset_b200._lib import check
    for d, p, data_len in [(3, 2, 4096), (3, 2, 1000), (4, 3, 777), (5, 4, 99)]:
        rs = ReedSolomon(ctx, d, p)
        n = 29
        stride = data_len + 5
        data = wl.payload_uniform(n, data_len, stride=stride, seed_extra=11)
        L = shard_len(data_len, d)
        dt = torch.from_numpy(data).to(DEV)
        par = torch.full((p, n, L), 0xEE, dtype=torch.uint8, device=DEV)
        guard = par.clone()
        check(ctx.lib.ss_rs_encode_uniform_dev(rs.h, dt.data_ptr(), stride, data_len, n, par.data_ptr(), n * L, L, 0))
        torch.cuda.synchronize()
        want = oracle.rs_encode_uniform(d, p, data, data_len, shard_stride=L)
        assert (par.cpu().numpy() == want).all()
        del guard


def test_encode_ragged_matches_oracle(ctx, oracle):
    rng = np.random.default_rng(5)
    for d, p in [(3, 2), (4, 3), (5, 4)]:
        rs = ReedSolomon(ctx, d, p)
        lens = np.concatenate([rng.integers(1, 3000, 200), [0, 0, 1, 2, 65536, 70001, 16, 48],
                               wl.CFG4_SIZES]).astype(np.uint32)
        rng.shuffle(lens)
        lay = wl.ragged_layout(lens, d)
        arena = rng.integers(0, 256, lay["data_bytes"] + 64, dtype=np.uint8)
        n = len(lens)
        want = np.zeros((p, lay["plane_bytes"]), dtype=np.uint8)
        oracle.rs_encode_batch(d, p, arena, lay["data_off"], lens, want.reshape(-1), lay["plane_bytes"], lay["par_off"])
        par = torch.zeros((p, lay["plane_bytes"]), dtype=torch.uint8, device=DEV)
        rs.encode_batch(torch.from_numpy(arena).to(DEV), torch.from_numpy(lay["data_off"].astype(np.int64)).to(DEV),
                        torch.from_numpy(lens.astype(np.int32)).to(DEV), par, lay["plane_bytes"],
                        torch.from_numpy(lay["par_off"].astype(np.int64)).to(DEV))
        torch.cuda.synchronize()
        assert (par.cpu().numpy() == want).all()
        assert rs.last_kernel().endswith("ragged_kernel")


def test_encode_against_golden_fixture(ctx):
    """committed fixture (tests/golden/make_golden.py): protects against oracle and GPU drifting together"""
    from pathlib import Path
    z = np.load(Path(__file__).parent / "golden" / "rs_golden.npz")
    for key in [k[5:] for k in z.files if k.startswith("data_")]:
        d, p, data_len = (int(x) for x in key.split("_"))
        rs = ReedSolomon(ctx, d, p)
        got = _gpu_encode_uniform(rs, z["data_" + key], data_len)
        assert (got == z["parity_" + key]).all(), key


# ---------------------------------------------------------------------------------------------
# batched reconstruct vs oracle
# ---------------------------------------------------------------------------------------------
def _planes_from(oracle, d, p, data, data_len):
    """all d+p shards of each codeword in plane layout [t, n, ds] (ds = round_up(L,16))"""
    n = data.shape[0]
    L = shard_len(data_len, d)
    ds = round_up(L, 16)
    planes = np.zeros((d + p, n, ds), dtype=np.uint8)
    for g in range(n):
        planes[:d, g, :L] = oracle.cw_split(data[g, :data_len].tobytes(), d)
    planes[d:] = oracle.rs_encode_uniform(d, p, data, data_len)
    return planes, L, ds


@pytest.mark.parametrize("d,p", [(3, 2), (4, 3), (5, 4), (2, 1), (6, 4)])
@pytest.mark.parametrize("data_only", [True, False])
def test_reconstruct_every_pattern(ctx, oracle, d, p, data_only):
    _check_reconstruct_every_pattern(ctx, oracle, d, p, data_only, 1000 + d, 0)


@pytest.mark.parametrize("d,p,data_len", [(3, 2, 4096), (3, 2, 9001), (4, 3, 4096), (2, 1, 3000), (3, 2, 16 * 3 * 64)])
@pytest.mark.parametrize("variant", [0, 1 << 16])
def test_reconstruct_long_codewords(ctx, oracle, d, p, data_len, variant):
    """codewords of more than 32 / 64 columns: the small-code kernel's one-column passes and (bit 16) its paired-column
    passes, full and partial second columns"""
    _check_reconstruct_every_pattern(ctx, oracle, d, p, False, data_len, variant)
    _check_reconstruct_every_pattern(ctx, oracle, d, p, True, data_len, variant)


def _check_reconstruct_every_pattern(ctx, oracle, d, p, data_only, data_len, variant):
    rs = ReedSolomon(ctx, d, p)
    rs.set_variant(variant)
    t = d + p
    pats = list(range(1 << t))
    n = len(pats)
    data = wl.payload_uniform(n, data_len, seed_extra=77 + d)
    full, L, ds = _planes_from(oracle, d, p, data, data_len)
    present = np.array(pats, dtype=np.uint32)
    damaged = full.copy()
    for g, pat in enumerate(pats):
        for j in range(t):
            if not (pat >> j) & 1:
                damaged[j, g, :] = 0xCC                    # garbage where the shard is missing
    off = (np.arange(n, dtype=np.uint64) * np.uint64(ds))
    lens = np.full(n, data_len, dtype=np.uint32)
    want = damaged.copy()
    st_want = oracle.rs_reconstruct_batch(d, p, want.reshape(-1), n * ds, off, lens, present, data_only)
    sh = torch.from_numpy(damaged).to(DEV)
    st = rs.reconstruct_batch(sh, n * ds, torch.from_numpy(off.astype(np.int64)).to(DEV),
                              torch.from_numpy(lens.astype(np.int32)).to(DEV),
                              torch.from_numpy(present.astype(np.int32)).to(DEV), data_only)
    torch.cuda.synchronize()
    got = sh.cpu().numpy()
    assert (st.cpu().numpy() == st_want).all()
    for g, pat in enumerate(pats):
        npres = bin(pat).count("1")
        if npres < d:
            assert st_want[g] == -10
            assert (got[:, g] == damaged[:, g]).all()       # nothing written
            continue
        upto = d if data_only else t
        for j in range(upto):
            assert (got[j, g, :L] == full[j, g, :L]).all(), (pat, j)
        for j in range(upto, t):
            assert (got[j, g] == damaged[j, g]).all()       # untouched
    # oracle agrees on the bytes it defines
    assert (got[:, :, :L] == want[:, :, :L]).all()


@pytest.mark.parametrize("d,p", [(3, 2), (4, 3), (2, 1)])
@pytest.mark.parametrize("data_len", [1, 17, 100, 4096, 12288, 20000])
@pytest.mark.parametrize("data_only", [True, False])
def test_reconstruct_uniform_every_pattern(ctx, oracle, d, p, data_len, data_only):
    """ss_rs_reconstruct_uniform_dev: RS(3,2) takes the pattern-specialised row kernel (compile-time decode rows per
    present mask, incl. codewords wider than one CTA pass), other codes the program kernels -- every present mask, several
    times over so CTAs walk mixed patterns, data-only and full, against the oracle (rscoding.rs:490-537)."""
    rs = ReedSolomon(ctx, d, p)
    t = d + p
    reps = 7
    pats = [pat for _ in range(reps) for pat in range(1 << t)]
    rng = np.random.default_rng(data_len + d)
    rng.shuffle(pats)
    n = len(pats)
    data = wl.payload_uniform(n, data_len, seed_extra=99 + d)
    full, L, ds = _planes_from(oracle, d, p, data, data_len)
    present = np.array(pats, dtype=np.uint32)
    damaged = full.copy()
    for g, pat in enumerate(pats):
        for j in range(t):
            if not (pat >> j) & 1:
                damaged[j, g, :] = 0xCC
    sh = torch.from_numpy(damaged).to(DEV)
    st = rs.reconstruct_uniform(sh, data_len, torch.from_numpy(present.astype(np.int32)).to(DEV), data_only)
    torch.cuda.synchronize()
    if (d, p) == (3, 2):
        assert rs.last_kernel() == "rs32_reconstruct_row_kernel"
    got = sh.cpu().numpy()
    st = st.cpu().numpy()
    upto = d if data_only else t
    for g, pat in enumerate(pats):
        if bin(pat).count("1") < d:
            assert st[g] == -10 and (got[:, g] == damaged[:, g]).all()      # refused, nothing written
            continue
        assert st[g] == 0
        for j in range(t):
            if j < upto:
                assert (got[j, g, :L] == full[j, g, :L]).all(), (pat, j)
                if not (pat >> j) & 1:
                    assert (got[j, g, L:] == 0).all(), "regenerated shards carry zero padding"
            else:
                assert (got[j, g] == damaged[j, g]).all()                    # untouched


def test_reconstruct_null_codeword_and_ragged(ctx, oracle):
    d, p = 3, 2
    rs = ReedSolomon(ctx, d, p)
    rng = np.random.default_rng(9)
    lens = np.array([0, 5, 4096, 1, 0, 333, 65536, 17], dtype=np.uint32)
    lay = wl.ragged_layout(lens, d)
    arena = rng.integers(0, 256, lay["data_bytes"] + 64, dtype=np.uint8)
    n = len(lens)
    pb = lay["plane_bytes"]
    planes = np.zeros((d + p, pb), dtype=np.uint8)
    for g in range(n):
        if lens[g] == 0:
            continue
        L = int(lay["L"][g]); o = int(lay["par_off"][g]); do = int(lay["data_off"][g])
        planes[:d, o:o + L] = oracle.cw_split(arena[do:do + int(lens[g])].tobytes(), d)
    par = np.zeros((p, pb), dtype=np.uint8)
    oracle.rs_encode_batch(d, p, arena, lay["data_off"], lens, par.reshape(-1), pb, lay["par_off"])
    planes[d:] = par
    present = np.array([31, 0b11100, 0b10110, 0b01011, 7, 0b11001, 0b01110, 0b00011], dtype=np.uint32)
    damaged = planes.copy()
    want = damaged.copy()
    st_want = oracle.rs_reconstruct_batch(d, p, want.reshape(-1), pb, lay["par_off"], lens, present, False)
    sh = torch.from_numpy(damaged).to(DEV)
    st = rs.reconstruct_batch(sh, pb, torch.from_numpy(lay["par_off"].astype(np.int64)).to(DEV),
                              torch.from_numpy(lens.astype(np.int32)).to(DEV),
                              torch.from_numpy(present.astype(np.int32)).to(DEV), False)
    torch.cuda.synchronize()
    assert st.cpu().tolist() == st_want.tolist() == [-20, 0, 0, 0, -20, 0, 0, -10]
    assert (sh.cpu().numpy() == planes).all()              # consistent codewords: reconstruction is the identity


# ---------------------------------------------------------------------------------------------
# fused accept step (config 3) and host-buffer entry points
# ---------------------------------------------------------------------------------------------
def test_accept_step_fused(ctx, oracle):
    d, p, data_len, n = 3, 2, 4096, 3000
    rs = ReedSolomon(ctx, d, p)
    data = wl.payload_uniform(n, data_len, seed_extra=3)
    planes = wl.cfg2_planes(n, 5, 0.8, seed_extra=3)
    L, ds, ps = rs.parity_layout(data_len, n)
    par = torch.empty((p, n, ds), dtype=torch.uint8, device=DEV)
    committed = torch.empty(n, dtype=torch.int64, device=DEV)
    bar = torch.empty(n, dtype=torch.int32, device=DEV)
    before = ctx.launches
    rs.accept_step_fused(torch.from_numpy(data).to(DEV), data_len, par, torch.from_numpy(planes.view(np.int64)).to(DEV),
                         4, committed, bar)
    torch.cuda.synchronize()
    assert ctx.launches == before + 1                       # ONE launch
    assert (par.cpu().numpy() == oracle.rs_encode_uniform(d, p, data, data_len)).all()
    c_want, b_want = oracle.tally_planes(planes, 4)
    assert (committed.cpu().numpy().view(np.uint64) == c_want).all()
    assert (bar.cpu().numpy().view(np.uint32) == b_want).all()


@pytest.mark.parametrize("d,p,data_len,n", [(4, 3, 4096, 3000), (5, 4, 4096, 2000), (2, 1, 777, 1000), (6, 4, 4090, 500),
                                            (8, 8, 1, 100), (4, 3, 100, 400000), (7, 2, 16 * 7 * 3, 900), (4, 2, 5000, 300), (3, 1, 3333, 300),
                                            (5, 4, 64, 1000), (2, 1, 4096, 1000)])
def test_generic_row_kernel_fused_step(ctx, oracle, d, p, data_len, n):
    """the d <= 8 Horner row kernel: fused step with the tally, grid-stride wrap (n > resident CTAs x waves), ragged tails"""
    rs = ReedSolomon(ctx, d, p)
    data = wl.payload_uniform(n, data_len, seed_extra=d * 31 + p)
    R = d + p
    planes = wl.cfg2_planes(n, R, 0.8, seed_extra=d)
    L, ds, ps = rs.parity_layout(data_len, n)
    par = torch.full((p, n, ds), 0x5a, dtype=torch.uint8, device=DEV)
    committed = torch.empty(n, dtype=torch.int64, device=DEV)
    bar = torch.empty(n, dtype=torch.int32, device=DEV)
    thr = R // 2 + 1
    before = ctx.launches
    rs.accept_step_fused(torch.from_numpy(data).to(DEV), data_len, par, torch.from_numpy(planes.view(np.int64)).to(DEV),
                         thr, committed, bar)
    torch.cuda.synchronize()
    assert ctx.launches == before + 1
    static = (d, p) in [(2, 1), (4, 3), (5, 4), (4, 2), (3, 1)]      # population 3/7/9/6/4 cluster codes
    tag = "<static code>" if static else "<nvrtc>"                  # any other code: specialised at run time (jit.cu)
    if not static:
        status = ctx.lib.ss_rs_jit_status(rs.h).decode()
        assert status.startswith("specialised by NVRTC"), status
    vpc = -(-L // 16)
    last = max(0, data_len - (d - 1) * L)
    packed = (min(L, last) // 16 != vpc) or vpc % 32 != 0          # some lanes would idle or be masked: packed layout
    assert rs.last_kernel() == ("horner_encode_packed_kernel" if packed else "horner_encode_row_kernel") + tag + "+tally"
    want = oracle.rs_encode_uniform(d, p, data, data_len)
    assert (par.cpu().numpy() == want).all()
    # run-time masks; forced one-codeword-per-pass layout; forced packed layout; pipelined packed loop; no NVRTC
    for v in (1 << 11, 1 << 13, 1 << 14, (1 << 14) | (1 << 11), (1 << 14) | (1 << 15), (1 << 13) | (1 << 11), 1 << 17, (1 << 17) | (1 << 14)):
        rs.set_variant(v)
        par3 = rs.encode_uniform(torch.from_numpy(data).to(DEV), data_len)
        torch.cuda.synchronize()
        is_packed = (packed or (v >> 14) & 1) and not (v >> 13) & 1
        is_static = static and not (v >> 11) & 1
        is_jit = not static and not (v >> 11) & 1 and not (v >> 17) & 1
        assert rs.last_kernel() == ("horner_encode_packed_kernel" if is_packed else "horner_encode_row_kernel") + \
            ("<static code>" if is_static else "<nvrtc>" if is_jit else ""), (v, rs.last_kernel())
        assert (par3.cpu().numpy() == want).all(), v
    c_want, b_want = oracle.tally_planes(planes, thr)
    assert (committed.cpu().numpy().view(np.uint64) == c_want).all()
    assert (bar.cpu().numpy().view(np.uint32) == b_want).all()
    # the flat kernel (variant 1) gives the same bytes
    rs.set_variant(1)
    par2 = rs.encode_uniform(torch.from_numpy(data).to(DEV), data_len)
    torch.cuda.synchronize()
    assert rs.last_kernel().startswith("horner_encode_uniform")
    assert torch.equal(par2, par)


def test_encode_uniform_host_buffers(ctx, oracle):
    for d, p, data_len, n in [(3, 2, 4096, 40000), (4, 3, 1000, 5000)]:
        rs = ReedSolomon(ctx, d, p)
        data = wl.payload_uniform(n, data_len, seed_extra=21)
        L = shard_len(data_len, d)
        for ss in (round_up(L, 16), L):                    # padded and exact host layouts
            par = np.full((p, n, ss), 0x77, dtype=np.uint8)
            rs.encode_uniform_host(data, data_len, par)
            want = oracle.rs_encode_uniform(d, p, data, data_len, shard_stride=ss)
            assert (par[:, :, :L] == want[:, :, :L]).all()


# ---------------------------------------------------------------------------------------------
# full-size, size-independent properties (BASELINE config 3 / 3b at 2^20 codewords x 4 KB)
# ---------------------------------------------------------------------------------------------
def test_full_size_encode_erase_decode_roundtrip(ctx, oracle):
    d, p, data_len, n = 3, 2, 4096, 1 << 20
    rs = ReedSolomon(ctx, d, p)
    L, ds, ps = rs.parity_layout(data_len, n)
    g = torch.Generator(device=DEV); g.manual_seed(wl.SEED_BASE + 3)
    data = torch.randint(0, 256, (n, data_len), dtype=torch.uint8, device=DEV, generator=g)
    # shard planes [5, n, ds]: data shards are views of the payload, copied out for the erase test
    sh = torch.zeros((d + p, n, ds), dtype=torch.uint8, device=DEV)
    padded = torch.zeros((n, d * L), dtype=torch.uint8, device=DEV)
    padded[:, :data_len] = data
    for i in range(d):
        sh[i, :, :L] = padded[:, i * L:(i + 1) * L]
    del padded
    rs.encode_uniform(data, data_len, parity=sh[d:])
    torch.cuda.synchronize()
    # (1) spot-check parity against the oracle on a sample of codewords
    idx = np.random.default_rng(1).integers(0, n, 257)
    idx[0], idx[1] = 0, n - 1
    sample = data[torch.from_numpy(idx).to(DEV)].cpu().numpy()
    want = oracle.rs_encode_uniform(d, p, sample, data_len)
    got = sh[d:][:, torch.from_numpy(idx).to(DEV)].cpu().numpy()
    assert (got == want).all()
    # (2) linearity checksum: XOR over all codewords of parity == parity of XOR over all codewords
    def xor_reduce(t):
        x = t.view(torch.int64) if t.shape[-1] % 8 == 0 else t
        while x.shape[0] > 1:
            h = x.shape[0] // 2
            r = x[:h] ^ x[h:2 * h]
            x = torch.cat([r, x[2 * h:]]) if x.shape[0] % 2 else r
        return x[0].view(torch.uint8) if x.dtype == torch.int64 else x[0]
    xd = xor_reduce(data).cpu().numpy()[None, :]
    xp = torch.stack([xor_reduce(sh[d + j]) for j in range(p)]).cpu().numpy()
    assert (oracle.rs_encode_uniform(d, p, np.ascontiguousarray(xd), data_len)[:, 0] == xp).all()
    # (3) erase per the cfg-3b distribution, reconstruct, compare with the original planes
    present = wl.erasure_patterns(n, d, p)
    keep = sh.clone()
    pm = torch.from_numpy(present.astype(np.int32)).to(DEV)
    for j in range(d + p):
        missing = ((pm >> j) & 1) == 0
        sh[j][missing] = 0x5A
    off = torch.arange(n, dtype=torch.int64, device=DEV) * ds
    lens = torch.full((n,), data_len, dtype=torch.int32, device=DEV)
    st = rs.reconstruct_batch(sh, n * ds, off, lens, pm, False)
    torch.cuda.synchronize()
    assert int(st.abs().sum()) == 0
    assert torch.equal(sh, keep)


@pytest.mark.gpu
def test_context_and_handles_may_be_destroyed_in_any_order(oracle):
    """ss_ctx_destroy before / after the coders and engines created on it (header: 'either order is safe'): the context
    holds one reference to itself and every handle one more; a coder that outlives its context refuses work but can still
    be destroyed, and a fresh context works afterwards."""
    from summerset_b200._lib import SummersetError
    from summerset_b200.api import Context, ReedSolomon
    from summerset_b200.engine import LeaderEngine
    rng = np.random.default_rng(11)
    data = [rng.integers(0, 256, 64, dtype=np.uint8) for _ in range(3)]
    want = [d.copy() for d in data] + [np.zeros(64, dtype=np.uint8) for _ in range(2)]
    assert oracle.rs_encode(3, 2, want) == 0

    def fresh():
        return [d.copy() for d in data] + [np.zeros(64, dtype=np.uint8) for _ in range(2)]

    for order in ("handles_first", "context_first"):
        c = Context(0, own_stream=True)
        rs1, rs2 = ReedSolomon(c, 3, 2), ReedSolomon(c, 4, 3)
        eng = LeaderEngine(c, "multipaxos", 128, 5)
        shards = fresh()
        rs1.encode(shards)
        assert all((a == b).all() for a, b in zip(shards, want))
        if order == "handles_first":
            rs1.close(); eng.close(); rs2.close(); c.close()
        else:
            c.close()
            with pytest.raises(SummersetError):
                rs1.encode(fresh())                     # the context is closed: calls fail loudly, nothing dangles
            rs2.close(); eng.close(); rs1.close()       # the last one frees the context
    c = Context(0, own_stream=True)
    rs = ReedSolomon(c, 3, 2)
    shards = fresh()
    rs.encode(shards)
    assert all((a == b).all() for a, b in zip(shards, want)) and rs.verify(shards)
