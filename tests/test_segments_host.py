"""Independent segments (splintr_b200/csrc/spl_segment.h + the filters spl_host.cpp builds): the device path cuts a
piece at every character boundary that no vocabulary key can cross and runs the merge loop of
/root/reference/src/core/bpe.rs:83-194 per segment.  Here the same __host__ __device__ code runs on the CPU
(tests/csrc/hosttest.cpp: ht_bpe_piece) against the plain merge loop over the whole piece and against the oracle."""
import random

import numpy as np
import pytest

import hostlib
from conftest import VOCABS, py_oracle
from splintr_b200 import presets as P

PID = {"cl100k_base": 0, "o200k_base": 1, "llama3": 1, "deepseek_v3": 1, "mistral_v3": 2}

CJK_COMMON = "的一是不了人我在有他这为之大来以个中上们到说国和地也子时道出而要于就下得可你年生自会那后能对着事其里所去行过家十用发天如然作方成者多日都三小军二无同么经法当起与好看学进种将还分此心前面又定见只主没公从"
CJK_RARE = "龘靐齉爨灪麤鱻饕鼗黻黼黽鼇鼈鼉鼊鼏鼐鼑鼒鼔鼕鼖鼗"
KANA = "あいうえおかきくけこさしすせそたちつてとなにぬねのはひふへほまみむめもやゆよらりるれろわをんアイウエオカキクケコサシスセソ"
HANGUL = "가나다라마바사아자차카타파하한국어대민공화조선인용이다는을를"
CYR = "абвгдежзийклмнопрстуфхцчшщъыьэюяАБВГДЕ"
HEB = "אבגדהוזחטיכלמנסעפצקרשת"
ARAB = "ابتثجحخدذرزسشصضطظعغفقكلمنهوي"
THAI = "กขคงจฉชซญดตถทนบปผพฟมยรลวสหอะาิีุู"
DEVA = "अआइईउऊएऐओऔकखगघचछजझटठडढणतथदधनपफबभमयरलवशषसह्ािीुू"
LATIN_X = "àáâãäåæçèéêëìíîïñòóôõöøùúûüýÿßœšžğışİ"
PUNCT = "，。！？：；“”‘’（）《》、…—·「」『』"
EMOJI = "😀😂🤣😊😍🥰😘🙏👍🔥✨🎉💯🚀🌟"
ASCII = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJ0123456789 .,!?-_'\"/()"
POOLS = [CJK_COMMON, CJK_COMMON, CJK_COMMON, CJK_RARE, KANA, HANGUL, CYR, HEB, ARAB, THAI, DEVA, LATIN_X, PUNCT, EMOJI, ASCII, ASCII]


def random_piece(rng: random.Random, max_chars: int) -> bytes:
    n = rng.randint(1, max_chars)
    mode = rng.random()
    if mode < 0.4:                       # one script
        pool = rng.choice(POOLS)
        s = "".join(rng.choice(pool) for _ in range(n))
    elif mode < 0.8:                     # two scripts, runs
        a, b = rng.choice(POOLS), rng.choice(POOLS)
        s = "".join(rng.choice(a if (i // rng.randint(1, 4)) % 2 == 0 else b) for i in range(n))
    else:                                # anything
        s = "".join(rng.choice(rng.choice(POOLS)) for _ in range(n))
    return s.encode()


@pytest.fixture(scope="module", params=VOCABS)
def tables(request):
    name = request.param
    p = P.PRESETS[name]
    return name, hostlib.HostTables(P.load_vocab_bytes(p.vocab_file), PID[name], p.byte_level, p.special_tokens)


def _raw_keys(o):
    """vocabulary keys as raw bytes (byte-level keys translated back, as the device tables hold them)"""
    if not o.byte_level:
        return list(o.encoder)
    from oracle.py_oracle import byte_level_decode_bytes
    return [r for r in (byte_level_decode_bytes(k) for k in o.encoder) if r]


def test_segmented_piece_equals_whole_merge_loop(tables):
    """60 000 random multi-script pieces per vocabulary: segment walker == merge loop over the whole piece."""
    name, t = tables
    rng = random.Random(sum(name.encode()))
    hostlib.seg_counters()
    for i in range(60000):
        piece = random_piece(rng, 14 if i % 4 else 60)
        if i % 16 == 5:
            hostlib.set_tile_limit(rng.randint(1, len(piece)))       # the piece leaves its tile here: no cuts beyond
        assert t.encode_piece(piece, True) == t.encode_piece(piece, False), (name, piece)
        hostlib.set_tile_limit()
    segs, single, beyond, safe = hostlib.seg_counters()
    assert safe > 0 and single > 0, "the filters never declared a boundary safe: the test exercised nothing"


def test_segmented_piece_equals_oracle(tables):
    """against byte_pair_encode of the oracle (whole-piece probe included) on pieces and on pre-tokenized text"""
    name, t = tables
    o = py_oracle(name)
    rng = random.Random(7)
    for _ in range(2500):
        piece = random_piece(rng, 12)
        assert t.encode_piece(piece, True) == o._encode_chunk(piece), (name, piece)
    for _ in range(600):
        s = " ".join(random_piece(rng, 10).decode() for _ in range(rng.randint(1, 6)))
        if "᠎" in s:
            continue
        assert t.encode(s.encode()) == o.encode(s), (name, s)


def test_keys_glued_together(tables):
    """Adversarial: pieces glued from vocabulary keys -- among them the keys that begin or end in the middle of a
    character (the irregular places of spl_segment.h) -- repaired to valid UTF-8 by dropping stray bytes."""
    name, t = tables
    o = py_oracle(name)
    keys = [k for k in _raw_keys(o) if any(b >= 0x80 for b in k)]
    assert len(keys) > 1000
    partial = []
    for k in keys:
        try:
            k.decode("utf-8")
        except UnicodeDecodeError:
            partial.append(k)
    assert partial, "no key with a cut character in this vocabulary?"
    rng = random.Random(11)
    n_checked = 0
    for i in range(40000):
        parts = [rng.choice(partial if rng.random() < 0.5 else keys) for _ in range(rng.randint(2, 5))]
        piece = b"".join(parts).decode("utf-8", "ignore").encode()
        if not piece:
            continue
        n_checked += 1
        assert t.encode_piece(piece, True) == t.encode_piece(piece, False), (name, piece)
        if i % 40 == 0:
            assert t.encode_piece(piece, True) == o._encode_chunk(piece), (name, piece)
    assert n_checked > 30000


def test_bytes_that_are_not_utf8(tables):
    """The C ABI takes bytes: malformed sequences must not be declared safe anywhere (same ids as the plain loop)."""
    name, t = tables
    rng = random.Random(3)
    for _ in range(20000):
        n = rng.randint(1, 40)
        piece = bytes(rng.choice([rng.randrange(256), rng.randrange(0x80, 0x100), 0xE4, 0xB8, 0xAD, 0x61]) for _ in range(n))
        assert t.encode_piece(piece, True) == t.encode_piece(piece, False), (name, piece)


def test_cjk_falls_apart(tables):
    """the point of the exercise: random CJK runs are mostly single-character segments"""
    name, t = tables
    rng = random.Random(5)
    hostlib.seg_counters()
    nchars = 0
    for _ in range(3000):
        k = rng.randint(4, 40)
        nchars += k
        t.encode_piece("".join(rng.choice(CJK_COMMON) for _ in range(k)).encode(), True)
    segs, single, beyond, safe = hostlib.seg_counters()
    assert beyond == 0
    assert single > 0.5 * nchars, (name, segs, single, nchars)
