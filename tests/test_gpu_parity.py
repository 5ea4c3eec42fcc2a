"""Parity tests proper: the CUDA path, called through the C-ABI (ctypes host in
splintr_b200/tokenizer.py), against the oracles on the same inputs.  Bit-exact ids and
offsets are required everywhere (integer work; no tolerance).

Reference behaviour under test: Tokenizer::encode / encode_with_special / encode_batch
(/root/reference/src/core/tokenizer.rs:729-808, 842-874, 932-942), byte_pair_encode
(src/core/bpe.rs:67-197), the split patterns (tokenizer.rs:39,42,64)."""
import random

import numpy as np
import pytest

import synth
from conftest import VOCABS, SP_VOCABS, py_oracle, c_oracle
from fuzz_alphabet import random_text

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def toks():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from splintr_b200 import Tokenizer
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = Tokenizer.from_pretrained(name, devices=[0])
        return cache[name]
    return get


def vocab_bytes(name):
    from splintr_b200 import presets as P
    return P.load_vocab_bytes(P.PRESETS[name].vocab_file)


@pytest.mark.parametrize("name", VOCABS)
def test_reference_golden_vectors(toks, name, ref_vectors):
    t = toks(name)
    for text, ids in ref_vectors[name]["encode"]:
        assert t.encode(text) == ids, (name, text)
        assert t.encode_rayon(text) == ids
    for text, ids in ref_vectors[name]["special"]:
        assert t.encode_with_special(text) == ids, (name, text)


@pytest.mark.parametrize("name", VOCABS)
def test_tiktoken_fixture(toks, name, xcheck_vectors):
    texts = [t for t, _ in xcheck_vectors[name]]
    assert toks(name).encode_batch(texts) == [ids for _, ids in xcheck_vectors[name]]


@pytest.mark.parametrize("name", VOCABS)
def test_fuzz_vs_oracles(toks, name):
    rng = random.Random(1234 + len(name))
    texts = [t for t in (random_text(rng, 80) for _ in range(6000)) if "᠎" not in t]
    texts += ["".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(20, 700))) for _ in range(60)]
    got = toks(name).encode_batch(texts)
    assert got == c_oracle(name).encode_batch(texts)
    po = py_oracle(name)
    for t, g in list(zip(texts, got))[:800]:
        assert g == po.encode(t), (name, t)


@pytest.mark.parametrize("name", VOCABS)
def test_edge_cases(toks, name):
    t, o = toks(name), c_oracle(name)
    assert t.encode_batch([]) == []
    assert t.encode("") == []
    assert t.encode_batch(["", "", ""]) == [[], [], []]
    cases = ["", "a", " ", "\n", "", "é", "\U0001f30d", "", "=" * 300, "-" * 77 + "\n", " " * 100 + "x", "#" * 33,
             "ab" * 100, "a" * 257, "a" * 4095, "a" * 4096, "a" * 4097, "x" * 5000, " " * 6000, "ab" * 4000,
             "\n" * 5000, "日本語のテキストをトークン化します。" * 9, "好" * 3000, "", "z"]
    assert t.encode_batch(cases) == o.encode_batch(cases)
    # ragged: documents whose boundaries fall on every offset of the 16-byte / 4096-byte tiling
    rag = ["x" * k for k in range(0, 40)] + ["word " * k for k in (818, 819, 820)] + ["é" * k for k in (2047, 2048, 2049)]
    assert t.encode_batch(rag) == o.encode_batch(rag)


@pytest.mark.parametrize("name", ["cl100k_base", "deepseek_v3"])
def test_tiles_dense_in_ids(toks, name, monkeypatch):
    """k_emit assembles a tile's ids in a 3 072-entry buffer: tiles with more ids than that (one id per byte: rare
    characters that fall apart into byte tokens, alternating one-byte pieces) take several phases, and a long piece
    may lie across the phase boundary.  Both output paths (SPL_EMIT_STAGE=0: per-thread stores) must agree."""
    rng = random.Random(99)
    rare = "龘靐齉爨灪麤鱻饕鼗黻黼黽鼇鼈鼉鼊"
    texts = ["".join(rng.choice(rare) for _ in range(6000)),                       # 18 000 bytes, mostly 1 id per byte
             "".join(rng.choice("!?;:") + rng.choice("\n\t") for _ in range(9000)),
             "".join(rng.choice(rare) for _ in range(900)) + "q" * 700 + "".join(rng.choice(rare) for _ in range(2000)),
             "".join(chr(rng.randrange(0x80, 0x250)) for _ in range(5000)),
             " ".join("".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(40, 400))) + rng.choice(rare) * 50 for _ in range(120))]
    want = c_oracle(name).encode_batch(texts)
    assert toks(name).encode_batch(texts) == want
    assert max(len(w) / len(t.encode()) for w, t in zip(want, texts)) > 0.8, "no tile of this test is dense in ids"
    from splintr_b200 import Tokenizer
    monkeypatch.setenv("SPL_EMIT_STAGE", "0")
    assert Tokenizer.from_pretrained(name, devices=[0]).encode_batch(texts) == want


@pytest.mark.parametrize("name", VOCABS)
def test_special_tokens(toks, name):
    t, po = toks(name), py_oracle(name)
    rng = random.Random(77)
    sp = list(po.special_tokens)
    texts = [s for s in sp[:12]] + [sp[0] + sp[1], "x" + sp[0], sp[0] + "y", sp[0][:-1], "<|", "a" + sp[2] + " b " + sp[3] + "\n"]
    texts += [random_text(rng, 30) + rng.choice(sp) + random_text(rng, 30) + rng.choice(sp) for _ in range(500)]
    texts = [x for x in texts if "᠎" not in x]
    got = t.encode_batch_with_special(texts)
    for x, g in zip(texts, got):
        assert g == po.encode_with_special(x), (name, x)
    # encode() never recognises specials (bindings.rs:246-256)
    assert t.encode(sp[0]) == po.encode(sp[0])
    assert t.decode(t.encode_with_special("a" + sp[0] + "b")) == "a" + sp[0] + "b"


def test_special_tokens_user_sets_that_overlap():
    """Tokenizer(vocab, pattern, special_tokens) takes ANY dict (bindings.rs:70-83; the Aho-Corasick automaton of
    tokenizer.rs:429-434 is built from it).  Strings that contain or overlap one another: the matches are those of
    MatchKind::Standard non-overlapping find_iter -- earliest end first, ties to the longest (k_resolve_specials,
    spl_special.h) -- checked against the oracle, through the batch call and in SentencePiece mode."""
    from splintr_b200 import Tokenizer, presets as P
    from oracle.py_oracle import OracleTokenizer
    rng = random.Random(5)
    sets = [{"<a>": 200001, "<a><b>": 200002, "<b>": 200003, "b><": 200004},
            {"<|end|>": 200001, "<|endoftext|>": 200002, "|><|": 200003, "<|": 200004},
            {"aa": 200001, "aaa": 200002, "ab": 200003}]
    for name in ("cl100k_base", "mistral_v2"):
        p = P.PRESETS[name]
        vb = P.load_vocab_bytes(p.vocab_file)
        for sp in sets:
            if p.sentencepiece:
                t = Tokenizer.from_bytes_sentencepiece(vb, p.pattern, sp)
            else:
                t = Tokenizer.from_bytes(vb, p.pattern, sp)
            o = OracleTokenizer.from_bytes(vb, p.pattern, sp, p.byte_level, p.sentencepiece)
            alpha = "".join(sp) + " xy\n"
            texts = ["", "<a><b>", "x<a><b><a>", "aaaaa aaaa aaa aa a"] + list(sp)
            texts += ["".join(rng.choice(alpha) for _ in range(rng.randint(0, 80))) for _ in range(400)]
            texts += [" ".join(rng.choice(list(sp) + ["hello", "world", "<", "|", ">"]) for _ in range(rng.randint(1, 30))) for _ in range(200)]
            got = t.encode_batch_with_special(texts)
            for x, g in zip(texts, got):
                assert g == o.encode_with_special(x), (name, sp, x)
            assert t.encode_batch(texts[:50]) == [o.encode(x) for x in texts[:50]]


@pytest.mark.parametrize("name", ["cl100k_base", "o200k_base"])
def test_batch_equals_individual_and_roundtrip(toks, name):
    """tests/cl100k.rs:191-214 and python/tests/test_cl100k.py:545-570 (700-text batch)."""
    t = toks(name)
    rng = random.Random(3)
    base = ["Hello, world!", "The quick brown fox jumps over the lazy dog.", "你好世界", "I'm sorry you're hurting—breakups suck.",
            "def f(x):\n    return x + 1\n", '{"a": [1, 2, 3]}', "   spaces   ", "MixedCASE and 12345 numbers"]
    texts = [rng.choice(base) + " " + str(i) for i in range(700)]
    batch = t.encode_batch(texts)
    assert len(batch) == 700
    for i in range(0, 700, 37):
        assert batch[i] == t.encode(texts[i])
    assert t.decode_batch(batch) == texts


def _check_packed(tok, orc, data, offsets):
    ids, off = tok.encode_packed(data, offsets)
    want_ids, want_off = orc.encode_packed(data, offsets)
    assert np.array_equal(off, want_off)
    assert np.array_equal(ids, want_ids)
    return ids, off


def test_cfg1_full(toks):
    d, o = synth.cfg1(vocab_bytes("cl100k_base"))
    _check_packed(toks("cl100k_base"), c_oracle("cl100k_base"), d, o)


def test_cfg2_full_size_bit_exact_and_roundtrip(toks):
    """BASELINE.json configs[1] at full size: 100 000 docs, ~100 MB."""
    d, o = synth.cfg2(vocab_bytes("cl100k_base"))
    tok = toks("cl100k_base")
    ids, off = _check_packed(tok, c_oracle("cl100k_base"), d, o)
    # size-independent property: decode(ids) reproduces the input bytes
    dec = tok._ensure_decoder()
    lut = [dec.get(i, b"") for i in range(max(dec) + 1)]
    sample = np.concatenate([np.arange(0, 2000), np.arange(50_000, 52_000), np.arange(98_000, 100_000)])
    raw = d.tobytes()
    for k in sample:
        got = b"".join(lut[i] for i in ids[int(off[k]):int(off[k + 1])].tolist())
        assert got == raw[int(o[k]):int(o[k + 1])]


def test_cfg3_mixed_o200k(toks):
    d, o = synth.cfg3(vocab_bytes("o200k_base"), 30_000)
    _check_packed(toks("o200k_base"), c_oracle("o200k_base"), d, o)


def test_cfg4_long_docs_llama3(toks):
    """BASELINE.json configs[3] on the 1 % document sample BASELINE.md allows: 100 docs of ~1 MB (100 MB)."""
    d, o = synth.cfg4(vocab_bytes("llama3"), 100, 1_000_000.0)
    _check_packed(toks("llama3"), c_oracle("llama3"), d, o)


def test_cfg5_cjk_deepseek_full_size(toks):
    """BASELINE.json configs[4] at full size: 100 000 CJK-heavy docs (~150 MB)."""
    d, o = synth.cfg5(b"", 100_000)
    _check_packed(toks("deepseek_v3"), c_oracle("deepseek_v3"), d, o)


@pytest.mark.parametrize("name", ["cl100k_base", "o200k_base", "llama3", "deepseek_v3"])
def test_merge_rounds_long_pieces(toks, name, monkeypatch):
    """k_bpe_long (windowed rounds, spl_encode.cu bpe_group) against the sequential loop of bpe.rs:119-167 in the C
    oracle: pieces of every length class; runs of one character (one slope across all lanes of a group), two-letter
    alphabets (long ties), random letters, CJK runs; lengths around the class and lane-block boundaries."""
    rng = random.Random(77)
    texts = []
    for ch in "=-#a \n.":
        texts += [ch * k for k in (33, 63, 64, 65, 66, 95, 127, 128, 129, 130, 191, 255, 256, 257, 300, 511, 512, 513, 700, 1023, 1024, 1025, 1500)]
    for alpha in ("ab", "abc", "etaoin", "abcdefghijklmnopqrstuvwxyz", "aeiou0123456789", "xyzXYZ_"):
        texts += ["".join(rng.choice(alpha) for _ in range(rng.randint(33, 1100))) for _ in range(150)]
    texts += ["".join(chr(rng.randint(0x4E00, 0x9FA5)) for _ in range(rng.randint(12, 330))) for _ in range(150)]
    texts += [" " + "".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(2 ** k - 1)) for k in range(5, 11)]
    rng.shuffle(texts)
    t, o = toks(name), c_oracle(name)
    assert t.encode_batch(texts) == o.encode_batch(texts)
    docs = ["\n".join(texts[i:i + 20]) for i in range(0, len(texts), 20)]     # many long pieces per warp task
    assert t.encode_batch(docs) == o.encode_batch(docs)


@pytest.mark.parametrize("name", VOCABS)
def test_segment_walker_multiscript(toks, name):
    """k_bpe, the segment walker (spl_segment.h): multi-script text -- CJK, kana, Hangul, Cyrillic, Hebrew, Arabic, Thai,
    Devanagari, emoji, accented Latin -- in pieces of every length class, pieces the walker has to leave to k_bpe_long
    (a segment beyond 32 bytes next to CJK), and keys glued together; against the C oracle."""
    from test_segments_host import random_piece, CJK_COMMON, CJK_RARE, HANGUL
    rng = random.Random(2024)
    texts = [" ".join(random_piece(rng, 14).decode() for _ in range(rng.randint(1, 12))) for _ in range(4000)]
    texts += ["".join(rng.choice(CJK_COMMON) for _ in range(k)) for k in (1, 2, 5, 10, 11, 12, 21, 22, 43, 85, 170, 171, 341, 342, 400, 1200, 5000)]
    texts += ["".join(rng.choice(CJK_RARE + HANGUL + CJK_COMMON) for _ in range(rng.randint(1, 200))) for _ in range(400)]
    # letters longer than a segment in the middle of a CJK run: one piece under O200K (Lo + Ll), the walker leaves it
    texts += ["".join(rng.choice(CJK_COMMON) for _ in range(rng.randint(1, 30))) + "".join(rng.choice("abcdefgh") for _ in range(rng.randint(30, 200))) +
              "".join(rng.choice(CJK_COMMON) for _ in range(rng.randint(0, 30))) for _ in range(300)]
    texts = [t for t in texts if "\u180e" not in t]
    t, o = toks(name), c_oracle(name)
    assert t.encode_batch(texts) == o.encode_batch(texts)
    doc = "\n".join(texts[:1500])
    assert t.encode(doc) == o.encode_batch([doc])[0]


def test_multi_id_characters_through_the_miss_list(monkeypatch):
    """Characters that are two or three ids are settled by k_probe's table: as SPL_PV_CHARREF values in pv, or -- in
    passes too large for that encoding, forced here with SPL_NO_CHARREF=1 -- as settled miss-list entries with their ids
    in the pool.  Both against the C oracle; SPL_NO_DEDUP=1 as well (every long piece through the merge loop)."""
    from splintr_b200 import Tokenizer
    from test_segments_host import CJK_COMMON, CJK_RARE, HANGUL, KANA
    rng = random.Random(9)
    texts = ["".join(rng.choice(CJK_RARE + HANGUL + KANA + CJK_COMMON + "，。 a") for _ in range(rng.randint(1, 4000))) for _ in range(60)]
    texts += ["".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(40, 600))) + " " for _ in range(40)] * 3
    for env in ("SPL_NO_CHARREF", "SPL_NO_DEDUP"):
        monkeypatch.setenv(env, "1")
        for name in ("deepseek_v3", "cl100k_base"):
            t = Tokenizer.from_pretrained(name, devices=[0])
            assert t.encode_batch(texts) == c_oracle(name).encode_batch(texts), (env, name)
            c = t.debug_counters()
            if env == "SPL_NO_CHARREF":
                assert c["settled_multi_id_chars"] > 0
            else:
                assert c["duplicates"] == 0
        monkeypatch.delenv(env)


def test_device_resident_entry_point(toks):
    import torch
    tok = toks("cl100k_base")
    d, o = synth.cfg2(vocab_bytes("cl100k_base"), 5000)
    n = len(d)
    buf = torch.zeros(n + ((-n) % 16), dtype=torch.uint8, device="cuda")
    buf[:n].copy_(torch.from_numpy(d))
    d_off = torch.from_numpy(o.astype(np.int64)).cuda()
    ids, out_off, n_tok = tok.encode_device(buf[:n], d_off)
    want_ids, want_off = c_oracle("cl100k_base").encode_packed(d, o)
    assert n_tok == len(want_ids)
    assert np.array_equal(ids[:n_tok].cpu().numpy().astype(np.uint32), want_ids)
    assert np.array_equal(out_off.cpu().numpy().astype(np.uint64), want_off)
    # idempotence: the same call again gives the same result (workspace is fully re-initialised)
    ids2, out2, n2 = tok.encode_device(buf[:n], d_off)
    assert n2 == n_tok and torch.equal(ids2[:n2], ids[:n_tok]) and torch.equal(out2, out_off)
    tok.set_profiling(True)
    tok.encode_device(buf[:n], d_off)
    kt = tok.last_kernel_times()
    tok.set_profiling(False)
    assert "k_probe" in kt and "k_emit" in kt and all(v >= 0 for v in kt.values())


def test_invalid_offsets_rejected(toks):
    tok = toks("cl100k_base")
    data = np.frombuffer(b"hello world", dtype=np.uint8)
    with pytest.raises(ValueError):
        tok.encode_packed(data, np.array([1, 11], dtype=np.uint64))
    with pytest.raises(ValueError):
        tok.encode_packed(data, np.array([0, 8, 4, 11], dtype=np.uint64))
    with pytest.raises(TypeError):
        tok.encode_batch(["ok", 5])


def test_multi_device_handle_matches_single(toks):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from splintr_b200 import Tokenizer
    d, o = synth.cfg2(vocab_bytes("cl100k_base"), 20_000)
    t2 = Tokenizer.from_pretrained("cl100k_base", devices=[0, 1])
    ids2, off2 = t2.encode_packed(d, o)
    ids1, off1 = toks("cl100k_base").encode_packed(d, o)
    assert np.array_equal(ids1, ids2) and np.array_equal(off1, off2)


def test_multi_device_handle_round_robin(toks, monkeypatch):
    """One handle over several devices: the chunks of the batch go to the devices round robin and come back in document
    order (spl_api.cu).  The same GPU listed three times drives exactly that code on a one-GPU box: default chunking,
    dozens of small chunks, empty documents on chunk boundaries, a document larger than a chunk, special tokens."""
    from splintr_b200 import Tokenizer
    o = c_oracle("cl100k_base")
    d, off = synth.cfg2(vocab_bytes("cl100k_base"), 6000)
    t3 = Tokenizer.from_pretrained("cl100k_base", devices=[0, 0, 0])
    ids, out = _check_packed(t3, o, d, off)
    ids1, out1 = toks("cl100k_base").encode_packed(d, off)
    assert np.array_equal(ids, ids1) and np.array_equal(out, out1)
    monkeypatch.setenv("SPL_CHUNK_BYTES", "30000")
    t2 = Tokenizer.from_pretrained("cl100k_base", devices=[0, 0])
    monkeypatch.delenv("SPL_CHUNK_BYTES")
    _check_packed(t2, o, d, off)
    texts = synth.unpack_texts(d, off)[:500]
    texts[0] = ""; texts[7] = ""; texts[8] = ""; texts[30] = "y" * 20_000; texts[31] = ""; texts[-1] = ""
    assert t2.encode_batch(texts) == o.encode_batch(texts)
    assert t2.encode_batch(["", "", ""]) == [[], [], []]
    assert t2.encode_batch([]) == []
    sp = "<|endoftext|>"
    texts2 = [tx[:40] + sp + tx[40:] for tx in texts[:200]]
    po = py_oracle("cl100k_base")
    assert t2.encode_batch_with_special(texts2) == [po.encode_with_special(x) for x in texts2]
    assert t2.decode_batch(t2.encode_batch(texts[100:120])) == texts[100:120]


def test_custom_vocab_with_unknown_bytes(toks):
    """bpe.rs:73-75,187-191: bytes that are not in the vocabulary are silently dropped."""
    import base64
    from splintr_b200 import Tokenizer, CL100K_BASE_PATTERN
    from oracle.py_oracle import OracleTokenizer
    toy = {b"a": 0, b"b": 1, b"c": 2, b"ab": 3, b"bc": 4, b"abc": 5, b" ": 6, b" a": 7}
    vb = b"".join(base64.b64encode(k) + b" " + str(v).encode() + b"\n" for k, v in toy.items())
    t = Tokenizer.from_bytes(vb, CL100K_BASE_PATTERN)
    o = OracleTokenizer.from_bytes(vb, CL100K_BASE_PATTERN)
    for s in ["a", "ab", "abc", "ac", "abcabc", "zab", "xyz", " a b", "cab abc", "aXbXc"]:
        assert t.encode(s) == o.encode(s), s


def test_huge_piece_scratch_grows_once_per_attempt(monkeypatch):
    """Pieces beyond 1 KiB that the segment walker leaves alone take global scratch (4 words per byte).  Several
    pipeline chunks that each outgrow the 64 MiB pool must make the call resize it ONCE and succeed (it used to
    quadruple per failing chunk and run out of memory); spl_encode_jsonl retries the same way."""
    import base64, json
    from splintr_b200 import Tokenizer, CL100K_BASE_PATTERN
    vb = b"".join(base64.b64encode(bytes([b])) + b" " + str(b).encode() + b"\n" for b in range(256))   # bytes only: no merges
    monkeypatch.setenv("SPL_CHUNK_BYTES", "6000000")
    t = Tokenizer.from_bytes(vb, CL100K_BASE_PATTERN)
    monkeypatch.delenv("SPL_CHUNK_BYTES")
    piece = "#" * 2000 + " "
    docs = [piece * 300 for _ in range(40)]                      # 24 MB, ~12 000 pieces of 2 000 bytes, ~4 chunks
    got = t.encode_batch(docs)
    want = list((piece * 300).encode())
    assert all(g == want for g in got)
    jl = "".join(json.dumps({"text": d}) + "\n" for d in docs).encode()
    ids, off = t.encode_jsonl(jl)
    assert len(off) == len(docs) + 1 and ids.tolist() == want * len(docs)


def test_threads_share_a_tokenizer(toks):
    """The reference's binding holds the GIL for a whole call, so threads may share a Tokenizer; here a per-object lock
    serialises them (one call in flight per handle)."""
    import threading
    tok, o = toks("cl100k_base"), c_oracle("cl100k_base")
    d, off = synth.cfg2(vocab_bytes("cl100k_base"), 2000)
    texts = synth.unpack_texts(d, off)
    want = o.encode_batch(texts)
    out, errs = {}, []

    def work(k):
        try:
            for r in range(6):
                lo = (k * 37 + r * 101) % 1500
                out[(k, r)] = (lo, tok.encode_batch(texts[lo:lo + 300]), tok.decode_batch(want[lo:lo + 20]))
        except Exception as e:                                   # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=work, args=(k,)) for k in range(6)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errs, errs
    for (k, r), (lo, ids, dec) in out.items():
        assert ids == want[lo:lo + 300] and dec == texts[lo:lo + 20]


def test_device_entry_async_status(toks):
    """sync=False: nothing is checked in the call; spl_device_status delivers the flags, and rejected offsets are never
    used as indices by the kernels behind k_mark_docs."""
    import torch
    tok = toks("cl100k_base")
    d, o = synth.cfg2(vocab_bytes("cl100k_base"), 300)
    n = len(d)
    buf = torch.zeros(n + ((-n) % 16), dtype=torch.uint8, device="cuda")
    buf[:n].copy_(torch.from_numpy(d))
    good = torch.from_numpy(o.astype(np.int64)).cuda()
    ids, out, _ = tok.encode_device(buf[:n], good, sync=False)
    assert tok.device_status() == 0
    want_ids, want_off = c_oracle("cl100k_base").encode_packed(d, o)
    assert np.array_equal(out.cpu().numpy().astype(np.uint64), want_off)
    bad = good.clone()
    bad[5] = 10 ** 12                                            # far outside the text
    bad[9] = bad[8] - 3                                          # decreasing
    tok.encode_device(buf[:n], bad, sync=False)
    assert tok.device_status() & 1
    with pytest.raises(ValueError):
        tok.encode_device(buf[:n], bad, sync=True)
    ids2, out2, nt = tok.encode_device(buf[:n], good)            # the handle is still usable
    assert nt == len(want_ids) and tok.device_status() == 0


def test_pipelined_host_call_many_chunks(toks, monkeypatch):
    """spl_encode_batch cuts a shard into pipeline chunks (H2D / kernels / D2H overlapped); force tiny chunks so
    one call runs through dozens of them, with empty and large documents on the chunk boundaries."""
    from splintr_b200 import Tokenizer
    monkeypatch.setenv("SPL_CHUNK_BYTES", "20000")
    t = Tokenizer.from_pretrained("cl100k_base", devices=[0])
    monkeypatch.delenv("SPL_CHUNK_BYTES")
    o = c_oracle("cl100k_base")
    d, off = synth.cfg2(vocab_bytes("cl100k_base"), 3000)
    _check_packed(t, o, d, off)
    texts = synth.unpack_texts(d, off)[:400]
    texts[3] = ""; texts[4] = ""; texts[17] = "x" * 70_000; texts[18] = ""; texts[-1] = ""
    assert t.encode_batch(texts) == o.encode_batch(texts)
    assert t.encode_batch(["", ""]) == [[], []]
    sp = "<|endoftext|>"
    texts2 = [tx[:50] + sp + tx[50:] for tx in texts[:200]]
    po = py_oracle("cl100k_base")
    assert t.encode_batch_with_special(texts2) == [po.encode_with_special(x) for x in texts2]


def test_large_shard_beyond_2gib_tiling_property(toks):
    """One device pass over more than 2^31 bytes (positions, tile indices and list indices are 32-bit inside the
    kernels; the limit is 4 GiB): the batch is cfg2 (100 MB, checked bit-exact above) repeated 22 times, so the ids
    must be the 100 MB result repeated and the offsets the same ramp shifted -- a size-independent property."""
    import torch
    tok = toks("cl100k_base")
    d, o = synth.cfg2(vocab_bytes("cl100k_base"))
    n1, nd1, reps = len(d), len(o) - 1, 22
    base = torch.from_numpy(d).cuda()
    n = n1 * reps
    assert n > (1 << 31)
    buf = torch.empty(n + ((-n) % 16), dtype=torch.uint8, device="cuda")
    buf[:n].view(reps, n1).copy_(base.unsqueeze(0).expand(reps, n1))
    o1 = torch.from_numpy(o.astype(np.int64)).cuda()
    d_off = torch.cat([o1[:-1] + r * n1 for r in range(reps)] + [torch.tensor([n], dtype=torch.int64, device="cuda")])
    b1 = torch.zeros(n1 + ((-n1) % 16), dtype=torch.uint8, device="cuda")
    b1[:n1].copy_(base)
    ids1, off1, nt1 = tok.encode_device(b1[:n1], o1)
    ids1 = ids1[:nt1].clone(); off1 = off1.clone()
    ids, off, nt = tok.encode_device(buf[:n], d_off)
    assert nt == nt1 * reps
    assert torch.equal(ids[:nt].view(reps, nt1), ids1.unsqueeze(0).expand(reps, nt1))
    want_off = torch.cat([off1[:-1] + r * nt1 for r in range(reps)] + [torch.tensor([nt], dtype=torch.int64, device="cuda")])
    assert torch.equal(off, want_off)


def test_cfg3_large_host_call_equals_parts(toks):
    """cfg3 (o200k_base, mixed prose / code / JSON) at 150 000 documents (~300 MB) through the pipelined host call:
    bit-exact against the C oracle, and identical to encoding the batch in three separate calls."""
    tok = toks("o200k_base")
    d, o = synth.cfg3(vocab_bytes("o200k_base"), 150_000)
    ids, off = _check_packed(tok, c_oracle("o200k_base"), d, o)
    cuts = [0, 50_000, 100_000, 150_000]
    parts_ids, parts_off = [], [np.zeros(1, dtype=np.uint64)]
    for a, b in zip(cuts[:-1], cuts[1:]):
        pi, po = tok.encode_packed(d[int(o[a]):int(o[b])], o[a:b + 1] - o[a])
        parts_ids.append(pi)
        parts_off.append(po[1:] + parts_off[-1][-1])
    assert np.array_equal(np.concatenate(parts_ids), ids)
    assert np.array_equal(np.concatenate(parts_off), off)


# ---- decode on the device ---------------------------------------------------------
@pytest.mark.parametrize("name", VOCABS + SP_VOCABS)
def test_decode_batch_matches_host_decode(toks, name):
    """spl_decode_batch against the ORACLE's decode_bytes (oracle/py_oracle.py, restating tokenizer.rs:877-897) on
    ordinary ids, special ids, ids outside the vocabulary (skipped) and, for the byte-level vocabulary, the
    untranslatable keys (ids 0-2 of deepseek_v3)."""
    t = toks(name)
    rng = random.Random(99)
    vs = t.vocab_size
    sp = list(t._special_tokens.values())
    lists = [[], [0], [1, 2, 3], [vs - 1], [vs + 5, 7, 2 ** 31 + 3], sp[:5], [sp[0], 11, sp[1]]]
    lists += [[rng.randrange(vs) for _ in range(rng.randint(0, 300))] for _ in range(400)]
    lists += [[rng.randrange(vs) for _ in range(5000)]]
    off = np.cumsum([0] + [len(x) for x in lists]).astype(np.uint64)
    ids = np.array([x for l in lists for x in l], dtype=np.uint32)
    data, boff = t.decode_packed(ids, off)
    raw = data.tobytes()
    o = py_oracle(name)                                          # decode_bytes of the reference, tokenizer.rs:877-897
    for i, l in enumerate(lists):
        assert raw[int(boff[i]):int(boff[i + 1])] == o.decode_bytes(l), (name, i)
        assert t.decode_bytes(l) == o.decode_bytes(l)            # the host-side single-list decode, same oracle
    assert t.decode_batch_lossy(lists[:50]) == [o.decode_lossy(l) for l in lists[:50]]


def test_decode_roundtrip_cfg2_full_and_device_entry(toks):
    """encode -> decode over the whole cfg2 batch reproduces the input bytes and document boundaries exactly
    (size-independent property), through the host call and through the device-resident entry point."""
    import torch
    tok = toks("cl100k_base")
    d, o = synth.cfg2(vocab_bytes("cl100k_base"))
    ids, off = tok.encode_packed(d, o)
    data, boff = tok.decode_packed(ids, off)
    assert np.array_equal(boff, o) and np.array_equal(data, d)
    d_ids = torch.from_numpy(ids.astype(np.int64)).to(torch.int32).cuda()
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    out, out_off = tok.decode_device(d_ids, d_off)
    assert torch.equal(out.cpu(), torch.from_numpy(d)) and np.array_equal(out_off.cpu().numpy().astype(np.uint64), o)
    bad = next(i for i in range(1000) if tok.decode_bytes([i]).decode("utf-8", errors="replace") != tok.decode_bytes([i]).decode("latin-1")
               and b"\xef\xbf\xbd" not in tok.decode_bytes([i]))
    with pytest.raises(ValueError):                            # a lone byte >= 0x80 is not valid UTF-8 (bindings.rs:300-304)
        tok.decode_batch([[bad]])


def test_decode_batch_strings_roundtrip(toks):
    t = toks("deepseek_v3")
    texts = ["Hello 你好 World 世界!", "", " hello world ", "日本語のテキスト" * 50, "émoji 🌍 ok"]
    assert t.decode_batch(t.encode_batch(texts)) == texts


# ---- SentencePiece mode on the device (mistral_v1 / mistral_v2; tokenizer.rs:737-795) -----------
def _sp_texts(seed, n, maxlen):
    rng = random.Random(seed)
    ws = [" ", " ", " ", "  ", "\n", "\t", "\x0b", "\x0c", "\r", "\xa0", "　", " ", "\x85", "\x1c"]
    out = []
    for _ in range(n):
        if rng.random() < 0.5:
            t = random_text(rng, maxlen)
        else:
            t = "".join(rng.choice(ws) if rng.random() < 0.45 else rng.choice("ab,Zé中🙂") for _ in range(rng.randint(0, maxlen)))
        if "᠎" not in t:
            out.append(t)
    return out


@pytest.mark.parametrize("name", SP_VOCABS)
def test_sentencepiece_golden_and_edges(toks, name, ref_vectors):
    t, o = toks(name), c_oracle(name)
    for text, ids in ref_vectors[name]["encode"]:
        assert t.encode(text) == ids, (name, text)
    for text, ids in ref_vectors[name]["special"]:
        assert t.encode_with_special(text) == ids, (name, text)
    assert t.vocab_size == {"mistral_v1": 32054, "mistral_v2": 32822}[name]        # tests/mistral_v2.rs:116
    assert t.encode_batch([]) == [] and t.encode_batch(["", ""]) == [[], []]
    cases = ["", " ", "  ", "a", "a ", " a", "a b", "a  b", "\n", " \n ", "\n\n  x", "a\tb", "a \t b", "　 a", "a 　 b",
             "x\x0b y", "\x0b", " \x0b ", "a \x85 b", "\xa0 x", "       x", "x       ", "a\r\n b", "Hello world", " world!",
             "Hello 🌍 World!", "def f():\n    return 1\n\n\nx = 2  # c\n", " " * 100, "\n" * 50, " " * 9000 + "y",
             "x\x0b" + " " * 9000 + "y", "a" * 4095 + " b", "a" * 4096 + " b", ("word " * 820), "é " * 2049, "好 " * 3000,
             "a" * 300, "ab " * 4000, ""]
    assert t.encode_batch(cases) == o.encode_batch(cases)
    # ids can outnumber the input bytes here: each of the 7 spaces is three byte tokens (E2 96 81)
    assert len(t.encode("       x")) == 22
    rag = [" " * k + "x" * (40 - k) for k in range(0, 41)] + ["x " * k for k in (2047, 2048, 2049)]
    assert t.encode_batch(rag) == o.encode_batch(rag)


@pytest.mark.parametrize("name", SP_VOCABS)
def test_sentencepiece_fuzz_and_special(toks, name):
    t, o, po = toks(name), c_oracle(name), py_oracle(name)
    texts = _sp_texts(31, 6000, 90)
    got = t.encode_batch(texts)
    assert got == o.encode_batch(texts)
    for x, g in list(zip(texts, got))[:600]:
        assert g == po.encode(x), (name, x)
    rng = random.Random(5)
    sp = list(po.special_tokens)
    st = [s for s in sp[:10]] + [sp[0] + sp[1], " " + sp[0] + " ", "a  " + sp[2] + "  b", sp[0][:-1]]
    st += [x + rng.choice(sp) + y + rng.choice(sp) + " " for x, y in zip(_sp_texts(7, 400, 30), _sp_texts(8, 400, 30))]
    gs = t.encode_batch_with_special(st)
    for x, g in zip(st, gs):
        assert g == po.encode_with_special(x), (name, x)


@pytest.mark.parametrize("name", SP_VOCABS)
def test_sentencepiece_roundtrip_and_batch(toks, name):
    """python/tests/test_mistral_v1.py:68-109,176-190: decode(encode(x)) == x (U+2581 -> space, tokenizer.rs:923-930),
    batch == individual; plus a 20 MB English batch against the C oracle, through the host call (many pipeline chunks,
    one mid-pipeline size read-back each) and through the device entry point."""
    import torch
    t, o = toks(name), c_oracle(name)
    texts = ["Hello, world!", "The quick brown fox jumps over the lazy dog.", "1234567890", " world!",
             "Unicode: こんにちは 世界 🦀", "Mixed: Hello 你好 🌍 World!", "Multi-line\ntext\nwith\nnewlines",
             'def hello_world():\n    print("Hello, World!")\n\nif __name__ == "__main__":\n    hello_world()\n']
    batch = t.encode_batch(texts)
    assert batch == [t.encode(x) for x in texts]
    assert t.decode_batch(batch) == texts and [t.decode(b) for b in batch] == texts
    d, off = synth.cfg2(vocab_bytes("cl100k_base"), 20_000)
    ids, out_off = _check_packed(t, o, d, off)
    n = len(d)
    buf = torch.zeros(n + ((-n) % 16), dtype=torch.uint8, device="cuda")
    buf[:n].copy_(torch.from_numpy(d))
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    dids, dout, n_tok = t.encode_device(buf[:n], d_off)
    assert n_tok == len(ids)
    assert np.array_equal(dids[:n_tok].cpu().numpy().astype(np.uint32), ids)
    assert np.array_equal(dout.cpu().numpy().astype(np.uint64), out_off)
    with pytest.raises(ValueError):                     # the id buffer must cover the transformed text
        t.encode_device(buf[:n], d_off, ids_out=torch.empty(n // 2, dtype=torch.int32, device="cuda"))


def test_sentencepiece_pipelined_small_chunks(toks, monkeypatch):
    from splintr_b200 import Tokenizer
    monkeypatch.setenv("SPL_CHUNK_BYTES", "20000")
    t = Tokenizer.from_pretrained("mistral_v2", devices=[0])
    monkeypatch.delenv("SPL_CHUNK_BYTES")
    o = c_oracle("mistral_v2")
    texts = _sp_texts(41, 3000, 200)
    texts[3] = ""; texts[17] = " " * 50_000; texts[18] = ""; texts[-1] = ""
    assert t.encode_batch(texts) == o.encode_batch(texts)
    assert t.pcre2().encode("a  b") == o.encode("a  b")           # clones keep the mode


# ---- JSON Lines ingestion on the device ---------------------------------------------
def test_ingest_jsonl_matches_json_module(toks):
    """spl_ingest_jsonl_device against Python's json module: documents = [json.loads(l)["text"] for non-blank l]
    (missing / non-string members -> empty documents), then encode_jsonl == encode_batch of those documents."""
    import json
    import torch
    from jsonl_cases import make_lines, join_lines
    tok = toks("cl100k_base")
    for seed, n, final_nl in ((1, 400, True), (2, 3000, False), (3, 1, True), (4, 20000, True)):
        lines, want = make_lines(seed, n)
        data = join_lines(lines, final_nl)
        buf = torch.zeros(len(data) + ((-len(data)) % 16) + 16, dtype=torch.uint8, device="cuda")
        buf[:len(data)].copy_(torch.frombuffer(bytearray(data), dtype=torch.uint8))
        text, offs, st = tok.ingest_jsonl_device(buf[:len(data)])
        raw = text.cpu().numpy().tobytes()
        o = offs.cpu().tolist()
        assert st["n_docs"] == len(want) and st["n_bad"] == 0
        assert [raw[o[i]:o[i + 1]].decode("utf-8") for i in range(len(want))] == want, seed
        assert st["n_missing"] == sum(1 for l in lines if l.strip(" \t\r") and not isinstance(json.loads(l).get("text"), str))
        ids, ioff = tok.encode_jsonl(data)
        exp = tok.encode_batch(want)
        flat = ids.tolist()
        io = ioff.tolist()
        assert [flat[io[i]:io[i + 1]] for i in range(len(want))] == exp
    # empty input, blank-only input, another member name, capacity query
    assert tok.encode_jsonl(b"")[1].tolist() == [0]
    assert tok.encode_jsonl(b"\n \n\t\n")[1].tolist() == [0]
    ids, ioff = tok.encode_jsonl(b'{"body":"Hello world","text":"no"}\n', field="body")
    assert ids.tolist() == [9906, 1917]
    with pytest.raises(ValueError):
        tok.ingest_jsonl_device(torch.zeros(16, dtype=torch.uint8, device="cuda")[:3], field="")


def test_encode_jsonl_many_chunks(toks, monkeypatch):
    """spl_encode_jsonl cuts the file at line ends into pipeline chunks; force small chunks so one call runs through
    dozens of them (blank lines, empty documents and a huge line on the boundaries), also in SentencePiece mode."""
    import json
    from splintr_b200 import Tokenizer
    from jsonl_cases import make_lines, join_lines
    monkeypatch.setenv("SPL_CHUNK_BYTES", "30000")
    t1 = Tokenizer.from_pretrained("cl100k_base", devices=[0])
    t2 = Tokenizer.from_pretrained("mistral_v2", devices=[0])
    monkeypatch.delenv("SPL_CHUNK_BYTES")
    lines, want = make_lines(77, 6000)
    lines[10] = json.dumps({"text": "x " * 40000}); want_big = "x " * 40000
    want = []
    for l in lines:
        if l.strip(" \t\r"):
            v = json.loads(l).get("text")
            want.append(v if isinstance(v, str) else "")
    assert want_big in want
    data = join_lines(lines)
    for t in (t1, t2):
        ids, off, st = t.encode_jsonl(data, return_stats=True)
        assert st["n_docs"] == len(want) and st["n_bad"] == 0
        flat, o = ids.tolist(), off.tolist()
        assert [flat[o[i]:o[i + 1]] for i in range(len(want))] == t.encode_batch(want)
    ids_s, off_s = t1.encode_jsonl(data, with_special=True)
    flat, o = ids_s.tolist(), off_s.tolist()
    assert [flat[o[i]:o[i + 1]] for i in range(len(want))] == t1.encode_batch_with_special(want)


def test_ingest_jsonl_cfg2_roundtrip(toks):
    """cfg2 (100 000 documents) written as JSON Lines with json.dumps and ingested on the device reproduces the packed
    batch byte for byte, and its ids equal the ids of the packed batch."""
    import json
    import torch
    tok = toks("cl100k_base")
    d, o = synth.cfg2(vocab_bytes("cl100k_base"), 100_000)
    texts = synth.unpack_texts(d, o)
    data = ("\n".join(json.dumps({"id": i, "text": t, "meta": {"text": "x"}}) for i, t in enumerate(texts)) + "\n").encode()
    n = len(data)
    buf = torch.zeros(n + ((-n) % 16) + 16, dtype=torch.uint8, device="cuda")
    buf[:n].copy_(torch.frombuffer(bytearray(data), dtype=torch.uint8))
    text, offs, st = tok.ingest_jsonl_device(buf[:n])
    assert st["n_docs"] == 100_000 and st["n_missing"] == 0 and st["n_bad"] == 0
    assert np.array_equal(text.cpu().numpy(), d) and np.array_equal(offs.cpu().numpy().astype(np.uint64), o)
    ids, ioff = tok.encode_jsonl(data)
    want_ids, want_off = tok.encode_packed(d, o)
    assert np.array_equal(ids, want_ids) and np.array_equal(ioff, want_off)


# ---- row N4, Parquet / Arrow: the column's pages are decompressed and decoded on the device ----------------------------

def _pq_bytes(table, **kw):
    import io
    import pyarrow.parquet as pq
    buf = io.BytesIO()
    pq.write_table(table, buf, **kw)
    return buf.getvalue()


def test_ingest_parquet_matches_pyarrow(toks):
    """Every page kind the reader takes (tests/test_parquet_host.py runs the same decoder on the CPU): device output
    == pyarrow's column, a null row an empty document; encode_parquet == encode_batch of that column."""
    import pyarrow as pa
    from test_parquet_host import OPTS, make_texts, expect
    tok = toks("cl100k_base")
    for i, opt in enumerate(OPTS):
        rng = random.Random(40 + i)
        for n, nulls, repeat in ((0, 0.0, False), (1, 0.0, False), (4000, 0.15, i % 2 == 0), (9000, 0.0, i % 2 == 1)):
            texts = make_texts(rng, n, nulls, repeat)
            data = _pq_bytes(pa.table({"id": pa.array(range(n)), "text": pa.array(texts, pa.string())}), **opt)
            want = expect(texts)
            text, offs, st = tok.ingest_parquet(data, "text")
            raw, o = text.cpu().numpy().tobytes(), offs.cpu().tolist()
            assert st["n_docs"] == n and o[0] == 0 and len(o) == n + 1
            assert [raw[o[k]:o[k + 1]] for k in range(n)] == want, (opt, n, nulls, repeat)
            ids, ioff = tok.encode_parquet(data, "text")
            flat, io_ = ids.tolist(), ioff.tolist()
            assert [flat[io_[k]:io_[k + 1]] for k in range(n)] == tok.encode_batch([w.decode() for w in want]), (opt, n)
    # a string inside nullable structs, by dotted path
    rng = random.Random(3)
    vals = [None if rng.random() < 0.1 else {"body": None if rng.random() < 0.2 else "doc %d é" % k} for k in range(3000)]
    t = pa.table({"meta": pa.array(vals, pa.struct([("body", pa.string())]))})
    text, offs, _ = tok.ingest_parquet(_pq_bytes(t), "meta.body")
    raw, o = text.cpu().numpy().tobytes(), offs.cpu().tolist()
    assert [raw[o[k]:o[k + 1]] for k in range(3000)] == [b"" if (v is None or v["body"] is None) else v["body"].encode() for v in vals]


def test_encode_parquet_cfg2_row_groups_and_batches(toks, monkeypatch):
    """cfg2 (100 000 documents) as a snappy Parquet file of 13 row groups: ingested text == the packed batch byte for
    byte, ids == the ids of the packed batch; the same through many small batches (SPL_CHUNK_BYTES), in SentencePiece
    mode, and with special tokens."""
    import pyarrow as pa
    from splintr_b200 import Tokenizer
    tok = toks("cl100k_base")
    d, o = synth.cfg2(vocab_bytes("cl100k_base"), 100_000)
    texts = synth.unpack_texts(d, o)
    table = pa.table({"id": pa.array(range(len(texts))), "text": pa.array(texts, pa.string())})
    want_ids, want_off = tok.encode_packed(d, o)
    for kw in (dict(compression="snappy", row_group_size=8000), dict(compression="none", use_dictionary=False, data_page_version="2.0")):
        data = _pq_bytes(table, **kw)
        text, offs, st = tok.ingest_parquet(data, "text")
        assert np.array_equal(text.cpu().numpy(), d) and np.array_equal(offs.cpu().numpy().astype(np.uint64), o)
        ids, ioff, st = tok.encode_parquet(data, "text", return_stats=True)
        assert np.array_equal(ids, want_ids) and np.array_equal(ioff, want_off)
        assert st["n_docs"] == 100_000 and st["h2d_bytes"] < len(data)
    data = _pq_bytes(table.slice(0, 30_000), compression="snappy", row_group_size=1500)
    sub = texts[:30_000]
    sub[7] = "<|endoftext|> and <|fim_prefix|>" + sub[7]
    data_s = _pq_bytes(pa.table({"text": pa.array(sub)}), compression="snappy", row_group_size=1500)
    monkeypatch.setenv("SPL_CHUNK_BYTES", "2000000")
    t1 = Tokenizer.from_pretrained("cl100k_base", devices=[0])
    t2 = Tokenizer.from_pretrained("mistral_v2", devices=[0])
    monkeypatch.delenv("SPL_CHUNK_BYTES")
    for t in (t1, t2):
        ids, ioff = t.encode_parquet(data, "text")
        w_ids, w_off = t.encode_batch_packed(texts[:30_000])
        assert np.array_equal(ids, w_ids) and np.array_equal(ioff, w_off)
    ids, ioff = t1.encode_parquet(data_s, "text", with_special=True)
    w_ids, w_off = t1.encode_batch_packed(sub, with_special=True)
    assert np.array_equal(ids, w_ids) and np.array_equal(ioff, w_off)
    assert 100257 in ids.tolist()[:int(ioff[8])]


def test_parquet_refusals_and_damaged_pages(toks):
    """What the reader does not take is refused with a message (ValueError), never guessed at; a page whose bytes do not
    decode is reported by the device (error bits -> ValueError), never read out of bounds."""
    import io
    import pyarrow as pa
    import pyarrow.parquet as pq
    tok = toks("cl100k_base")
    table = pa.table({"text": pa.array(["a", "b"]), "n": pa.array([1, 2]), "l": pa.array([["x"], ["y", "z"]])})
    for column, kw, pat in (("body", {}, "no column named"), ("n", {}, "BYTE_ARRAY"), ("l.list.element", {}, "repeated"),
                            ("text", dict(compression="zstd"), "snappy"),
                            ("text", dict(use_dictionary=False, column_encoding={"text": "DELTA_BYTE_ARRAY"}), "DELTA")):
        with pytest.raises(ValueError, match=pat):
            tok.encode_parquet(_pq_bytes(table, **kw), column)
    with pytest.raises(ValueError, match="not a Parquet file"):
        tok.encode_parquet(b"hello world, this is not parquet", "text")
    rng = random.Random(8)
    texts = ["".join(rng.choice("abcdefgh ") for _ in range(rng.randint(0, 300))) for _ in range(3000)]
    good = _pq_bytes(pa.table({"text": pa.array(texts)}), compression="snappy", use_dictionary=False, data_page_size=20000)
    col = pq.ParquetFile(io.BytesIO(good)).metadata.row_group(0).column(0)
    lo, hi = col.data_page_offset + 40, col.data_page_offset + col.total_compressed_size
    outcomes = {"ok": 0, "err": 0}
    for k in range(12):
        b = bytearray(good)
        for _ in range(3):
            b[rng.randrange(lo, hi)] ^= 1 << rng.randrange(8)
        try:
            ids, off = tok.encode_parquet(bytes(b), "text")
            assert len(off) == 3001
            outcomes["ok"] += 1
        except ValueError as e:
            assert "parquet" in str(e)
            outcomes["err"] += 1
    assert outcomes["err"] > 0
    ids, off = tok.encode_parquet(good, "text")                      # the handle is fine afterwards
    assert [ids[off[k]:off[k + 1]].tolist() for k in range(50)] == tok.encode_batch(texts[:50])


def test_encode_arrow_string_columns(toks):
    """An Arrow string array is the packed layout already: data buffer + offsets go in as they are (nulls, slices,
    chunked arrays, 64-bit offsets)."""
    import pyarrow as pa
    tok = toks("o200k_base")
    rng = random.Random(21)
    texts = [None if rng.random() < 0.1 else random_text(rng, 60) for _ in range(5000)]
    texts = [t if (t is None or "᠎" not in t) else "x" for t in texts]
    want = tok.encode_batch([t or "" for t in texts])

    def check(col, exp):
        ids, off = tok.encode_arrow(col)
        flat, o = ids.tolist(), off.tolist()
        assert len(o) == len(exp) + 1 and o[0] == 0
        assert [flat[o[k]:o[k + 1]] for k in range(len(exp))] == exp

    a = pa.array(texts, pa.string())
    check(a, want)
    check(a.slice(1234, 2000), want[1234:3234])
    check(pa.array(texts, pa.large_string()), want)
    check(pa.chunked_array([a.slice(0, 100), a.slice(100, 0), a.slice(100, 4900)]), want)
    check(pa.array([], pa.string()), [])
    # a null slot that holds bytes (legal in Arrow): must still be an empty document
    import numpy as _np
    data = pa.py_buffer(b"helloJUNKworld")
    offs = pa.py_buffer(_np.array([0, 5, 9, 14], dtype=_np.int32).tobytes())
    valid = pa.py_buffer(bytes([0b101]))
    odd = pa.Array.from_buffers(pa.string(), 3, [valid, offs, data])
    assert odd.to_pylist() == ["hello", None, "world"]
    check(odd, tok.encode_batch(["hello", "", "world"]))
    with pytest.raises(TypeError):
        tok.encode_arrow(pa.array([1, 2, 3]))
