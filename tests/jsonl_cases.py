"""JSON Lines test material shared by the CPU test of the line parser and the GPU parity test (row N4)."""
import json
import random

ALPHA = ["a", "b", " ", "Z", "9", '"', "\\", "/", "\n", "\t", "\r", "\b", "\f", "\x01", "é", "中", "\U0001f642",
         " ", "{", "}", "[", "]", ":", ",", "'"]


def rand_str(rng, n):
    return "".join(rng.choice(ALPHA) for _ in range(rng.randint(0, n)))


def rand_value(rng, depth=0):
    k = rng.randrange(8 if depth < 2 else 5)
    if k == 0:
        return rand_str(rng, 20)
    if k == 1:
        return rng.randint(-1000, 1000)
    if k == 2:
        return rng.random() * 1e6
    if k == 3:
        return rng.choice([True, False, None])
    if k == 4:
        return rand_str(rng, 5)
    if k == 5:
        return [rand_value(rng, depth + 1) for _ in range(rng.randint(0, 3))]
    return {rng.choice(["text", "t", "x", rand_str(rng, 4)]): rand_value(rng, depth + 1) for _ in range(rng.randint(0, 3))}


def make_lines(seed, n, field="text", maxlen=120):
    """(raw line strings, expected document strings): valid JSON objects with the member present / absent / not a
    string / duplicated, random member order, nested look-alikes, both ensure_ascii settings, blank lines."""
    rng = random.Random(seed)
    lines, want = [], []
    for _ in range(n):
        r = rng.random()
        if r < 0.05:
            lines.append(rng.choice(["", " ", "\t \r", "  "]))            # blank: skipped
            continue
        members = []
        for _k in range(rng.randint(0, 4)):
            members.append((rng.choice(["id", "meta", "t", "texts", "Text", rand_str(rng, 5)]), rand_value(rng)))
        if r < 0.80:
            members.insert(rng.randint(0, len(members)), (field, rand_str(rng, maxlen)))
            if rng.random() < 0.1:                                         # duplicate name: the last one wins
                members.append((field, rand_str(rng, 10) if rng.random() < 0.7 else 7))
        elif r < 0.88:
            members.insert(rng.randint(0, len(members)), (field, rng.choice([1, None, True, [1, "text"], {"text": "inner"}])))
        ea = rng.random() < 0.5
        parts = [json.dumps(k, ensure_ascii=ea) + rng.choice([":", ": ", " : "]) + json.dumps(v, ensure_ascii=ea) for k, v in members]
        line = rng.choice(["", " "]) + "{" + rng.choice([",", ", ", " , "]).join(parts) + "}" + rng.choice(["", " ", "\r"])
        if rng.random() < 0.05:                                            # the member name itself escaped
            line = line.replace('"' + field + '"', '"' + "".join("\\u%04x" % ord(c) for c in field) + '"', 1)
        got = json.loads(line).get(field)                                  # ground truth: the json module
        lines.append(line)
        want.append(got if isinstance(got, str) else "")
    return lines, want


def join_lines(lines, final_newline=True):
    return ("\n".join(lines) + ("\n" if final_newline else "")).encode("utf-8")
