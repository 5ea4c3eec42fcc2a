"""The windowed merge rounds of k_bpe_long (splintr_b200/csrc/spl_encode.cu, bpe_group) restated in Python and held
against the oracle's sequential loop (byte_pair_encode = bpe.rs:67-197):
tests/bpe_batch_sim.py is the round as an algorithm (m, theta, commit), tests/bpe_lane_model.py a lane-level model
with the per-lane bitmasks of the kernel.  CPU only; the kernel itself is covered by tests/test_gpu_parity.py."""
import os
import random
import sys

import pytest


from conftest import py_oracle                                   # noqa: E402
from oracle.py_oracle import byte_pair_encode                    # noqa: E402
import bpe_batch_sim                                             # noqa: E402
import bpe_lane_model                                            # noqa: E402


def _pieces(rng, n):
    out = []
    for _ in range(n):
        kind = rng.randrange(5)
        if kind == 0:
            out.append(bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(2, 300))))
        elif kind == 1:
            out.append(bytes([rng.choice(b"=-# \n")]) * rng.randint(2, 300))
        elif kind == 2:
            out.append(bytes(rng.choice(b"ab") for _ in range(rng.randint(2, 200))))
        elif kind == 3:
            out.append(bytes(rng.choice(b"aeiotnsr ETAOIN0123.,-_") for _ in range(rng.randint(2, 200))))
        else:
            out.append("".join(chr(rng.randint(0x4E00, 0x9FA5)) for _ in range(rng.randint(1, 40))).encode())
    return out


@pytest.mark.parametrize("name", ["cl100k_base", "o200k_base", "llama3"])
def test_windowed_rounds_equal_sequential_loop(name):
    enc = py_oracle(name).encoder
    rng = random.Random(2024)
    rounds = merges = 0
    for piece in _pieces(rng, 400):
        want = byte_pair_encode(piece, enc)
        got, r = bpe_batch_sim.batched_bpe(piece, enc, 0)
        assert got == want, piece
        rounds += r
        merges += max(len(piece) - len(want), 0)
    assert rounds * 4 < merges                                    # the point of the windows: far fewer rounds than merges


def test_lane_model_bitmaps_equal_sequential_loop():
    enc = py_oracle("cl100k_base").encoder
    dec = {v: k for k, v in enc.items()}
    rng = random.Random(99)
    for LG in range(6):
        for piece in _pieces(rng, 60):
            piece = piece[:32 << LG]
            if len(piece) < 2 or piece in enc:
                continue
            got, _ = bpe_lane_model.group_bpe(piece, enc, dec, LG)
            assert got == byte_pair_encode(piece, enc), (LG, piece)


def _m_by_definition(ranks):
    """m(i) = !(m(i-1) and key(i-1) < key(i)) and !(m(i+1) and key(i+1) < key(i)), pairs visited in key order."""
    NONE = 0x1FFFFF
    n = len(ranks)
    m = [False] * n
    for i in sorted((i for i in range(n) if ranks[i] != NONE), key=lambda i: (ranks[i], i)):
        m[i] = not (i > 0 and m[i - 1]) and not (i + 1 < n and m[i + 1])      # a neighbour with m set was visited before: its key is lower
    return m


def test_kernel_bit_tricks_equal_the_definition_of_m():
    """spl_bpe_bits.h (the code k_bpe_long runs: carry-ripple run masks, bit reversal, boundary bits between lanes)
    compiled for the host, against m evaluated pair by pair in key order."""
    import hostlib
    NONE = 0x1FFFFF
    rng = random.Random(5)
    worst = 0
    for trial in range(6000):
        G = 1 << rng.randrange(6)
        B = rng.randint(2, 32)
        n = rng.randint(1, G * B - 1)                               # pairs; parts = n + 1 <= G * B
        B = max(2, (n + 1 + G - 1) // G)                            # the kernel's block size for that many parts
        kind = rng.randrange(5)
        if kind == 0:
            ranks = [rng.randrange(300, 5000) for _ in range(n)]
        elif kind == 1:
            ranks = [777] * n                                       # one character repeated: a single slope over all lanes
        elif kind == 2:
            ranks = [rng.choice((400, 401)) for _ in range(n)]      # long ties
        elif kind == 3:
            ranks = sorted(rng.randrange(300, 900) for _ in range(n))
            if rng.random() < 0.5:
                ranks.reverse()                                     # monotone: one slope, to either side
        else:
            ranks = [rng.choice((NONE, 300, 301, 5000)) for _ in range(n)]
        got, passes = hostlib.bpe_window_m(ranks, G, B)
        assert got == _m_by_definition(ranks), (G, B, ranks)
        worst = max(worst, passes)
    assert worst <= 33
