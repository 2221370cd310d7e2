"""The synthetic-workload generator and the mini-batch row mask are specified on Philox4x32-10 (oracle/synth_oracle.c header).
This file restates the spec a third time in pure Python, pins Philox to the published Random123 known answers, and checks the
C twin against it -- so the fixture the GPU generator is compared with (tests/test_gpu_parity.py) is itself pinned."""
import ctypes as C
import math

import numpy as np
import pytest

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c, k = list(ctr), list(key)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & MASK, p1 & MASK, ((p0 >> 32) ^ c[3] ^ k[1]) & MASK, p0 & MASK]
        k = [(k[0] + W0) & MASK, (k[1] + W1) & MASK]
    return c


def irwin_hall4(a, b):
    return (a & 0xFFFF) + (a >> 16) + (b & 0xFFFF) + (b >> 16) - 131070


SCALE64 = 1.7320508075688772 / 65536.0


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 with 10 rounds
    assert philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert philox4x32_10([MASK] * 4, [MASK] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


@pytest.mark.parametrize("seed,row0", [(42, 0), (0x1234567890ABCDEF, 5_000_000_000)])
def test_dense_generator_matches_spec(oracle, seed, row0):
    rows, d = 5, 13                         # odd d: the last column uses half a Philox draw
    X = oracle.synth_dense_f32(seed, row0, rows, d)
    key = [seed & MASK, seed >> 32]
    scale = np.float32(SCALE64)
    for r in range(rows):
        i = row0 + r
        for j in range(d):
            o = philox4x32_10([i & MASK, i >> 32, j >> 1, 1], key)
            t = irwin_hall4(o[0], o[1]) if j % 2 == 0 else irwin_hall4(o[2], o[3])
            assert X[r, j] == np.float32(t) * scale
    w = oracle.synth_wtrue(seed, d)
    for j in range(d):
        o = philox4x32_10([j, 0, 0, 2], key)
        assert w[j] == (irwin_hall4(o[0], o[1]) * SCALE64) / math.sqrt(d)


def test_labels_match_spec(oracle):
    seed, rows, d = 42, 40, 16
    key = [seed & MASK, seed >> 32]
    X = oracle.synth_dense_f32(seed, 0, rows, d)
    w = oracle.synth_wtrue(seed, d)
    for kind in ("logistic", "least_squares", "hinge"):
        y = oracle.synth_labels(seed, kind, 0, X, w)
        for i in range(rows):
            m = 0.0
            for j in range(d):
                m += float(X[i, j]) * w[j]
            o = philox4x32_10([i, 0, 0, 3], key)
            u = ((o[0] >> 5) * 67108864.0 + (o[1] >> 6) + 0.5) * 2.0 ** -53
            if kind == "logistic":
                want = 1.0 if m + math.log(u) - math.log(1.0 - u) > 0 else 0.0
            elif kind == "hinge":
                want = 1.0 if m > 0 else 0.0
                if u < 0.05:
                    want = 1.0 - want
            else:
                e = philox4x32_10([i, 0, 0, 4], key)
                want = m + 0.1 * (irwin_hall4(e[0], e[1]) * SCALE64)
            assert y[i] == want, (kind, i)


def test_csr_generator_matches_spec(oracle):
    seed, rows, d, k = 7, 6, 1000, 8
    key = [seed & MASK, seed >> 32]
    rowptr, idx, val = oracle.synth_csr_f32(seed, 3, rows, d, k)
    assert np.array_equal(rowptr, np.arange(rows + 1) * k)
    stride = d // k
    for r in range(rows):
        for t in range(k):
            o = philox4x32_10([3 + r, 0, t, 5], key)
            assert idx[r * k + t] == t * stride + o[0] % stride
            assert val[r * k + t] == np.float32(irwin_hall4(o[1], o[2])) * np.float32(SCALE64)


def test_row_mask_matches_spec_and_fraction(oracle):
    """data.sample(false, fraction, 42 + i) of runMiniBatchSGD is realised as: keep row g iff the 64-bit Philox draw keyed by the
    seed at counter (g_lo, g_hi, 0, 6) is below fraction * 2^64."""
    L = oracle.lib()
    L.oracle_row_selected.restype = C.c_int
    L.oracle_row_selected.argtypes = [C.c_uint64, C.c_uint64, C.c_int64]
    seed = 42 + 3
    thresh = int(math.ldexp(0.3, 64))
    for g in list(range(50)) + [2 ** 33 + 5]:
        o = philox4x32_10([g & MASK, g >> 32, 0, 6], [seed & MASK, seed >> 32])
        assert L.oracle_row_selected(seed, thresh, g) == int(((o[0] << 32) | o[1]) < thresh)
    kept = sum(L.oracle_row_selected(seed, thresh, g) for g in range(20000))
    assert abs(kept / 20000 - 0.3) < 0.015
    assert all(L.oracle_row_selected(seed, 0, g) == 1 for g in range(10))      # threshold 0 = no sampling
