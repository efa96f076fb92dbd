"""A THIRD implementation of the crate's arithmetic, sharing nothing with oracle/ss_oracle.c or
summerset_b200/csrc/gf256.hpp (same author, same session -> common-mode risk): GF(2^8) as polynomials over
GF(2) reduced modulo x^8+x^4+x^3+x^2+1 with sympy.polys.galoistools -- no log/antilog tables, no byte tricks -- and the
coding matrix M = vandermonde(d+p, d) * inverse(top d x d) built by plain Gauss-Jordan on top of it.

It pins, against BOTH the oracle and the product's host code (and through tests/cpp/test_static_codes.cpp the GPU tables):
the parity rows of every code the protocols construct, the survey's derived known-answer test, the upstream 5+5 vector,
and 1000 random products / 255 inverses.  Parity stays "unpinned vs the reference" until someone runs
tests/golden/verify_with_cargo.rs on a box with Rust; this test only removes the shared-author risk.
"""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

sympy = pytest.importorskip("sympy")
from sympy.polys.domains import ZZ  # noqa: E402
from sympy.polys.galoistools import gf_gcdex, gf_mul, gf_rem  # noqa: E402

ROOT = Path(__file__).resolve().parent.parent
POLY = [1, 0, 0, 0, 1, 1, 1, 0, 1]            # x^8 + x^4 + x^3 + x^2 + 1  (0x11D)


def _bits(b):
    out = [(b >> i) & 1 for i in range(7, -1, -1)]
    while out and out[0] == 0:
        out.pop(0)
    return out


def _byte(poly):
    v = 0
    for c in poly:
        v = (v << 1) | (int(c) & 1)
    return v


def pmul(a, b):
    return _byte(gf_rem(gf_mul(_bits(a), _bits(b), 2, ZZ), POLY, 2, ZZ))


def pinv(a):
    s, _, g = gf_gcdex(_bits(a), POLY, 2, ZZ)           # s*a + t*POLY = 1
    assert g == [1]
    return _byte(gf_rem(s, POLY, 2, ZZ))


def ppow(a, n):
    """the crate's galois::exp: a^0 = 1 (also for a = 0), 0^n = 0"""
    if n == 0:
        return 1
    r = 1
    for _ in range(n):
        r = pmul(r, a)
    return r


def pmat_inv(m):
    n = len(m)
    a = [row[:] + [1 if i == j else 0 for j in range(n)] for i, row in enumerate(m)]
    for col in range(n):
        piv = next(r for r in range(col, n) if a[r][col])
        a[col], a[piv] = a[piv], a[col]
        s = pinv(a[col][col])
        a[col] = [pmul(x, s) for x in a[col]]
        for r in range(n):
            if r != col and a[r][col]:
                f = a[r][col]
                a[r] = [x ^ pmul(f, y) for x, y in zip(a[r], a[col])]
    return [row[n:] for row in a]


def pcoding_matrix(d, p):
    v = [[ppow(r, c) for c in range(d)] for r in range(d + p)]
    ti = pmat_inv([row[:] for row in v[:d]])
    return [[_xor_sum(pmul(v[r][k], ti[k][c]) for k in range(d)) for c in range(d)] for r in range(d + p)]


def _xor_sum(it):
    r = 0
    for x in it:
        r ^= x
    return r


def _gf_dump(tmp_path, *args):
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    exe = tmp_path / "gf_dump"
    if not exe.exists():
        subprocess.run([gxx, "-std=c++17", "-O1", str(ROOT / "tests" / "cpp" / "gf_dump.cpp"), "-o", str(exe)], check=True)
    return subprocess.run([str(exe), *map(str, args)], capture_output=True, text=True, check=True).stdout


CODES = [(2, 1), (3, 2), (4, 2), (4, 3), (5, 4), (3, 1), (6, 4)]
SURVEY_ROWS = {   # SURVEY.md 8c table (Backblaze / klauspost matrices)
    (3, 2): ["01 01 01", "0f 08 06"],
    (4, 2): ["1b 1c 12 14", "1c 1b 14 12"],
    (4, 3): ["1b 1c 12 14", "1c 1b 14 12", "12 14 1b 1c"],
    (5, 4): ["07 07 06 06 01", "09 08 09 08 01", "0f 0e 0e 0f 01", "02 7d 95 fd 16"],
}


@pytest.mark.parametrize("d,p", CODES)
def test_matrix_three_ways(oracle, tmp_path, d, p):
    mine = pcoding_matrix(d, p)
    assert [row for row in mine[:d]] == [[1 if i == j else 0 for j in range(d)] for i in range(d)], "systematic top block"
    assert (np.array(mine, dtype=np.uint8) == oracle.rs_matrix(d, p)).all(), "oracle matrix differs from the polynomial construction"
    prod = [[int(x, 16) for x in line.split()] for line in _gf_dump(tmp_path, "matrix", d, p).strip().splitlines()]
    assert prod == mine, "product (gf256.hpp) matrix differs from the polynomial construction"
    if (d, p) in SURVEY_ROWS:
        assert [" ".join(f"{x:02x}" for x in row) for row in mine[d:]] == SURVEY_ROWS[(d, p)]


def test_random_products_and_inverses(oracle, tmp_path):
    rng = np.random.default_rng(2026)
    pairs = rng.integers(0, 256, size=(1000, 2))
    mine = [pmul(int(a), int(b)) for a, b in pairs]
    assert mine == [oracle.gf_mul(int(a), int(b)) for a, b in pairs]
    prod = [int(x) for x in _gf_dump(tmp_path, "mul", *pairs.reshape(-1).tolist()).split()]
    assert prod == mine
    for a in range(1, 256):
        assert pmul(a, pinv(a)) == 1
    # the crate's log/exp anchors quoted in SURVEY 8c: EXP[0..9], LOG[2..8]
    assert [ppow(2, i) for i in range(10)] == [1, 2, 4, 8, 16, 32, 64, 128, 29, 58]
    log = {ppow(2, i): i for i in range(255)}
    assert [log[v] for v in range(2, 9)] == [1, 25, 2, 50, 26, 198, 3]


def test_known_answer_vectors(oracle):
    # SURVEY 8c derived KAT: bincode("interesting_value") = 0x11 || bytes, RS(3,2), L = 6
    payload = bytes([17]) + b"interesting_value"
    sh = [list(payload[i * 6:(i + 1) * 6]) for i in range(3)]
    m = pcoding_matrix(3, 2)
    par = [[_xor_sum(pmul(m[3 + j][i], sh[i][b]) for i in range(3)) for b in range(6)] for j in range(2)]
    assert bytes(par[0]).hex() == "2b6c7b717e70" and bytes(par[1]).hex() == "2ffb9ccc5da8"
    shards = [np.array(s, dtype=np.uint8) for s in sh] + [np.zeros(6, dtype=np.uint8) for _ in range(2)]
    assert oracle.rs_encode(3, 2, shards) == 0
    assert shards[3].tobytes().hex() == "2b6c7b717e70" and shards[4].tobytes().hex() == "2ffb9ccc5da8"
    # decode matrix for present {1,3,4}: rows f5 69 9d / 01 00 00 / f4 68 9d
    sub = [m[1], m[3], m[4]]
    assert [" ".join(f"{x:02x}" for x in row) for row in pmat_inv(sub)] == ["f5 69 9d", "01 00 00", "f4 68 9d"]
    # upstream crate test_one_encode (5+5): data [0,1],[4,5],[2,3],[6,7],[8,9] -> parity [12,13],[10,11],[14,15],[90,91],[94,95]
    m55 = pcoding_matrix(5, 5)
    data = [[0, 1], [4, 5], [2, 3], [6, 7], [8, 9]]
    par = [[_xor_sum(pmul(m55[5 + j][i], data[i][b]) for i in range(5)) for b in range(2)] for j in range(5)]
    assert par == [[12, 13], [10, 11], [14, 15], [90, 91], [94, 95]]
