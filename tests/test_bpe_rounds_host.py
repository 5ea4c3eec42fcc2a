"""The windowed merge rounds of k_bpe_long (splintr_b200/csrc/spl_encode.cu, bpe_group) restated in Python and held
against the oracle's sequential loop (byte_pair_encode = bpe.rs:67-197):
tools/bpe_batch_sim.py is the round as an algorithm (m, theta, commit), tools/bpe_lane_model.py a lane-level model
with row bitmaps and shuffled masks.  CPU only; the kernel itself is covered by tests/test_gpu_parity.py."""
import os
import random
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

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
