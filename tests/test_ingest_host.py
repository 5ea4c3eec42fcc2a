"""JSON Lines ingestion, the per-line parser of splintr_b200/csrc/spl_ingest.h run on the CPU
(tests/csrc/hosttest.cpp) against Python's json module: member lookup (last duplicate wins, nested look-alikes
ignored, escaped names), every escape incl. surrogate pairs, blank lines, missing / non-string members."""
import json

import hostlib
from jsonl_cases import make_lines, join_lines


def test_parser_matches_json_module():
    for seed in range(40):
        lines, want = make_lines(seed, 150, maxlen=120 if seed % 4 else 700)    # long values: the 8-bytes-at-a-time paths
        docs, missing, bad = hostlib.jsonl(join_lines(lines, final_newline=seed % 2 == 0))
        assert bad == 0
        assert [d.decode("utf-8") for d in docs] == want, seed
        assert missing == sum(1 for l in lines if l.strip(" \t\r") and not isinstance(json.loads(l).get("text"), str))


def test_escapes_and_edges():
    U = "\\" + "u"                                             # backslash-u, spelled so that no tool rewrites it
    cases = [
        (r'{"text":"a\"b\\c\/d\b\f\n\r\t"}', 'a"b\\c/d\b\f\n\r\t'),
        ('{"text":"' + U + '00e9' + U + '4e2d' + U + 'd83d' + U + 'de42 ' + U + '0078"}', chr(0xe9) + chr(0x4e2d) + chr(0x1f642) + " x"),
        ('{"text":"A' + U + '0000z"}', "A" + chr(0) + "z"),
        ('{"text":"raw ' + chr(0xe9) + chr(0x4e2d) + chr(0x1f642) + '"}', "raw " + chr(0xe9) + chr(0x4e2d) + chr(0x1f642)),
        (r'{"a":{"text":"inner"},"text":"outer"}', "outer"),
        (r'{"a":["text",{"text":"x"}],"b":"text","text":"y"}', "y"),
        (r'{"text":"first","text":"second"}', "second"),
        ('{"' + U + '0074e' + U + '0078t":"escaped name"}', "escaped name"),
        (r'{"text" : "spaced" , "z":1}', "spaced"),
        (r'{"texts":"no","Text":"no","tex":"no"}', ""),
        (r'{"text":5}', ""), (r'{"text":null}', ""), (r'{}', ""), (r'{"text":""}', ""),
        (r'{"s":"}{\"text\":\"trap\"","text":"ok"}', "ok"),
    ]
    for line, want in cases:                                   # the table itself agrees with the json module
        got = json.loads(line).get("text")
        assert (got if isinstance(got, str) else "") == want
    data = "\n".join(l for l, _ in cases).encode("utf-8")
    docs, missing, bad = hostlib.jsonl(data)
    assert [d.decode("utf-8") for d in docs] == [w for _, w in cases]
    assert bad == 0 and missing == 4
    # lone surrogates become U+FFFD; lines that are not objects are counted and left empty
    blob = ('{"text":"a' + U + 'd800b' + U + 'dc00c"}').encode() + b"\n[1,2]\nnot json\n" + b'{"text":"unterminated' + b"\n\n  \n" + b'{"text":"last"}'
    docs, missing, bad = hostlib.jsonl(blob)
    assert [d.decode("utf-8") for d in docs] == ["a" + chr(0xfffd) + "b" + chr(0xfffd) + "c", "", "", "", "last"] and bad == 3 and missing == 0
    assert hostlib.jsonl(b"") == ([], 0, 0) and hostlib.jsonl(b"\n\n") == ([], 0, 0)
    assert hostlib.jsonl(b'{"body":"x"}', "body")[0] == [b"x"]
