"""CPU-side checks: the C-ABI library loads and exports every symbol include/summerset_b200.h declares,
fails loudly without a GPU, and the host-side mirrors (Bitmap, RSCodeword bookkeeping) behave like the
reference's (src/utils/bitmap.rs:312-420, src/utils/rscoding.rs:697-786 host-only parts)."""
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from summerset_b200 import _lib
from summerset_b200.api import Bitmap, RSCodeword, SummersetError, bincode_string, shard_len

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = (ROOT / "include" / "summerset_b200.h").read_text()
    declared = set(re.findall(r"\b(ss_[a-z0-9_]+)\s*\(", header))
    declared -= {"ss_ctx", "ss_rs_coder"}
    assert len(declared) >= 38
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert lib.ss_version() == 200
    assert b"no CPU fallback" in lib.ss_strerror(_lib.SS_ERR_NO_DEVICE)


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_gpu_fails_loudly():
    import ctypes as C
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.ss_ctx_create(0, C.byref(h))
    assert rc == _lib.SS_ERR_NO_DEVICE and not h
    assert b"no CPU fallback" in lib.ss_last_error()
    from summerset_b200.api import Context
    with pytest.raises(SummersetError):
        Context(0)


# ---- src/utils/bitmap.rs:312-420 ---------------------------------------------------------------
def test_bitmap_new_invalid():
    with pytest.raises(AssertionError, match="invalid bitmap size 0"):
        Bitmap(0, True)


def test_bitmap_conversions():
    ref = Bitmap.from_indices(5, range(1, 4))
    assert Bitmap.from_indices(5, [1, 2, 3]) == ref
    assert Bitmap.from_indices(5, {1, 2, 3}) == ref
    assert [i for i, f in ref.iter() if f] == [1, 2, 3]


def test_bitmap_set_get():
    m = Bitmap(7, False)
    m.set(0, True); m.set(1, False); m.set(2, True)
    with pytest.raises(SummersetError):
        m.set(7, True)
    assert m.get(0) is True and m.get(1) is False and m.get(2) is True and m.get(3) is False
    with pytest.raises(SummersetError):
        m.get(7)


def test_bitmap_flip_union_count_iter():
    m = Bitmap(5, False); m.set(1, True); m.flip()
    assert m == Bitmap.from_indices(5, [0, 2, 3, 4])
    a = Bitmap.from_indices(5, [0, 1, 3]); a.union(Bitmap.from_indices(5, [0, 4]))
    assert a == Bitmap.from_indices(5, [0, 1, 3, 4])
    with pytest.raises(SummersetError):
        a.union(Bitmap(6, False))
    c = Bitmap(7, False)
    assert c.count() == 0
    c.set(0, True); c.set(2, True); c.set(3, True)
    assert c.count() == 3
    ref = [True, True, False, True, True]
    m = Bitmap(5, True); m.set(2, False)
    for i, f in m.iter():
        assert ref[i] == f


def test_bitmap_bincode():
    m = Bitmap.from_indices(10, [0, 2, 3, 9]); m.set(5, True)
    # bit length 10, one usize block, value 0b1000101101 = 557 -> varint 0xFB + u16
    assert m.encode() == bytes([10, 1, 251]) + (557).to_bytes(2, "little")
    m = Bitmap.from_indices(24, [1, 5, 7, 12, 17, 23])
    v = sum(1 << i for i in [1, 5, 7, 12, 17, 23])
    assert m.encode() == bytes([24, 1, 252]) + v.to_bytes(4, "little")


# ---- src/utils/rscoding.rs host-only bookkeeping ------------------------------------------------
DATA = bincode_string("interesting_value")


def test_rscodeword_new_from_data():
    assert DATA == bytes([17]) + b"interesting_value"
    data_len = len(DATA)
    L = shard_len(data_len, 3)
    with pytest.raises(SummersetError):
        RSCodeword.from_data(DATA, 0, 0)
    cw = RSCodeword.from_data(DATA, 3, 0)
    assert (cw.num_data_shards(), cw.num_parity_shards(), cw.num_shards()) == (3, 0, 3)
    assert (cw.avail_data_shards(), cw.avail_parity_shards(), cw.avail_shards()) == (3, 0, 3)
    assert cw.avail_shards_map() == Bitmap.from_indices(3, [0, 1, 2])
    assert cw.data_len() == data_len and cw.shard_len() == L
    cw = RSCodeword.from_data(DATA, 3, 2)
    assert (cw.num_data_shards(), cw.num_parity_shards(), cw.num_shards()) == (3, 2, 5)
    assert (cw.avail_data_shards(), cw.avail_parity_shards(), cw.avail_shards()) == (3, 0, 3)
    assert cw.avail_shards_map() == Bitmap.from_indices(5, [0, 1, 2])
    assert cw.data_len() == data_len and cw.shard_len() == L


def test_rscodeword_new_from_null():
    with pytest.raises(SummersetError):
        RSCodeword.from_null(0, 0)
    cw = RSCodeword.from_null(3, 2)
    assert (cw.num_data_shards(), cw.num_parity_shards(), cw.num_shards()) == (3, 2, 5)
    assert (cw.avail_data_shards(), cw.avail_parity_shards(), cw.avail_shards()) == (0, 0, 0)
    assert cw.avail_shards_map() == Bitmap(5, False)
    assert cw.data_len() == 0 and cw.shard_len() == 0


def test_rscodeword_subset_absorb():
    cwa = RSCodeword.from_data(DATA, 3, 2)
    with pytest.raises(SummersetError):
        cwa.subset_copy(Bitmap.from_indices(6, [0, 5]), False)
    cw01 = cwa.subset_copy(Bitmap.from_indices(5, [0, 1]), False)
    assert cw01.avail_data_shards() == 2
    cw02 = cwa.subset_copy(Bitmap.from_indices(5, [0, 2]), True)
    assert cw02.avail_data_shards() == 2 and cw02.data_copy is not None
    cwb = RSCodeword.from_null(3, 2)
    cwb.absorb_other(cw02)
    assert cwb.avail_shards() == 2 and cwb.avail_shards_map() == Bitmap.from_indices(5, [0, 2])
    cwb.absorb_other(cw01)
    assert cwb.avail_shards() == 3 and cwb.avail_shards_map() == Bitmap.from_indices(5, [0, 1, 2])
    assert cwb.get_data() == DATA
    with pytest.raises(SummersetError):
        cwb.absorb_other(RSCodeword.from_data(DATA, 5, 3))


def test_rscodeword_p0_and_null_paths_need_no_coder():
    """rscoding.rs:454-456,498-507,549-557: p == 0 is decided on the host, never reaching the coder."""
    cw = RSCodeword.from_data(DATA, 3, 0)
    cw.compute_parity(None)
    assert cw.avail_parity_shards() == 0 and cw.verify_parity(None) is True
    cw.reconstruct_all(None)
    cw.shards[1] = None
    with pytest.raises(SummersetError):
        cw.reconstruct_all(None)
    with pytest.raises(SummersetError):
        cw.reconstruct_data(None)
    null = RSCodeword.from_null(3, 2)
    for fn in (null.compute_parity, null.verify_parity, null.reconstruct_all, null.reconstruct_data):
        with pytest.raises(SummersetError, match="null"):
            fn(None)
    cw = RSCodeword.from_data(DATA, 3, 2)
    with pytest.raises(SummersetError, match="None"):
        cw.compute_parity(None)


def test_crossword_brr_assignment_matches_oracle(oracle):
    """host-side policy helper (crossword/mod.rs:866-888) against the oracle restatement"""
    from summerset_b200.api import crossword_brr_assignment
    for n, T in [(5, 5), (3, 3), (7, 7), (5, 10), (9, 9), (4, 8)]:
        for spr in range(1, T + 1):
            assert [int(x) for x in oracle.cw_brr_assignment(n, T, spr)] == crossword_brr_assignment(n, T, spr)


def test_nvrtc_specialisation_compiles_without_a_gpu():
    """jit.cu: the row / packed encode kernels compile under NVRTC for a code that has no compile-time table (Crossword's
    RS(6,4)), offline -- the same source, options and name expressions a coder uses on the GPU box."""
    import ctypes as C
    from summerset_b200 import _lib
    lib = _lib.load()
    log = C.create_string_buffer(16384)
    n = lib.ss_jit_selftest(6, 4, log, 16384)
    assert n > 50000, (n, log.value.decode())
    assert lib.ss_jit_selftest(9, 3, log, 16384) < 0 and b"d <= 8" in log.value      # outside the specialised range


def test_crossword_slot_pitch_matches_the_header():
    """api.cw_slot_pitch is SS_CW_SLOT_PITCH of include/summerset_b200.h: the shard length rounded up to 32 bytes (DRAM
    sector), for ints and arrays; the macro text is checked so the two cannot drift apart silently."""
    import re
    from pathlib import Path
    import numpy as np
    from summerset_b200.api import cw_slot_pitch
    hdr = (Path(__file__).resolve().parent.parent / "include" / "summerset_b200.h").read_text()
    m = re.search(r"#define SS_CW_SLOT_PITCH\(L\) \(\(\(\(uint64_t\)\(L\)\) \+ (\d+)u\) & ~\(uint64_t\)(\d+)u\)", hdr)
    assert m and m.group(1) == m.group(2) == "31"
    for L, want in [(0, 0), (1, 32), (31, 32), (32, 32), (33, 64), (86, 96), (171, 192), (1366, 1376), (21846, 21856)]:
        assert cw_slot_pitch(L) == want
    arr = np.array([1, 32, 33, 10923], dtype=np.int64)
    assert (cw_slot_pitch(arr) == np.array([32, 32, 64, 10944])).all()
