"""Emulation of the windowed multi-merge BPE round used by k_bpe ("merge windows") against the
sequential leftmost-min loop of the oracle (bpe.rs:119-167).  Counts rounds per piece.

A round, on the current part list with pair keys key[e] = (rank(e, next e), e):
  1. m[e] (pair e merges if no new pair interferes): local minima of key are true; a pair with a lower neighbour
     that merges is false; fixpoint of  m[e] = !(m[l] and key[l] < key[e]) and !(m[r] and key[r] < key[e]).
  2. every m-true pair looks up the ranks of the pairs its merge can create: (left part, T), (T, right part) and,
     when the pair two to the right is m-true as well, (T, T').  theta = min of those ranks.
  3. commit the m-true pairs with rank < theta (the global minimum always).  Re-rank around them.
The sequential loop performs exactly these merges before any other (no pair created in the window ranks below theta).

Usage: python tests/bpe_batch_sim.py [vocab] [n_pieces]
"""
import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.py_oracle import byte_pair_encode, load_tiktoken_bpe
from splintr_b200.presets import get_preset, load_vocab_bytes

NONE = 0xFFFFFFFF


def batched_bpe(piece: bytes, enc, refine: int = 1):
    n = len(piece)
    if n == 1:
        r = enc.get(piece)
        return ([] if r is None else [r]), 0
    r = enc.get(piece)
    if r is not None:
        return [r], 0
    parts = [piece[i:i + 1] for i in range(n)]          # live parts, in order (the device keeps a linked list)
    rounds = 0

    def rank(a, b):
        return enc.get(a + b, NONE)

    while True:
        L = len(parts)
        rk = [rank(parts[i], parts[i + 1]) for i in range(L - 1)]
        if not rk or min(rk) == NONE:
            break
        rounds += 1
        key = [(rk[i], i) for i in range(L - 1)]
        live = [rk[i] != NONE for i in range(L - 1)]

        def before(a, b):      # pair a merges before pair b in the static order
            return 0 <= a < L - 1 and live[a] and key[a] < key[b]
        m = [live[i] and not before(i - 1, i) and not before(i + 1, i) for i in range(L - 1)]   # local minima
        while True:
            nm = [live[i] and not (before(i - 1, i) and m[i - 1]) and not (before(i + 1, i) and m[i + 1]) for i in range(L - 1)]
            if nm == m:
                break
            m = nm
        cand = []                                        # (rank, new rank lower bound)
        for i in range(L - 1):
            if not m[i]:
                continue
            T = parts[i] + parts[i + 1]
            nr = NONE
            if i > 0:
                nr = min(nr, rank(parts[i - 1], T))
            if i + 2 < L:
                nr = min(nr, rank(T, parts[i + 2]))
            if i + 3 < L and i + 2 < L - 1 and m[i + 2]:
                nr = min(nr, rank(T, parts[i + 2] + parts[i + 3]))
            cand.append((rk[i], i, nr))
        theta = min(c[2] for c in cand)
        for _ in range(refine):                          # T1 <= T3 <= ... <= exact window end
            t2 = min([c[2] for c in cand if c[0] < theta] + [NONE])
            theta = min([c[2] for c in cand if c[0] < t2] + [NONE])
        gmin = min(c[:2] for c in cand)
        commit = [c[1] for c in cand if c[0] < theta or c[:2] == gmin]
        out, i, cs = [], 0, set(commit)
        while i < L:
            if i in cs:
                out.append(parts[i] + parts[i + 1]); i += 2
            else:
                out.append(parts[i]); i += 1
        parts = out
    ids = []
    for p in parts:
        r = enc.get(p)
        if r is not None:
            ids.append(r)
        else:
            ids.extend(enc[bytes([b])] for b in p if bytes([b]) in enc)
    return ids, rounds


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "llama3"
    npieces = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    refine = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    p = get_preset(name)
    enc = load_tiktoken_bpe(load_vocab_bytes(p.vocab_file))
    rng = random.Random(7)
    tot_r = tot_seq = 0
    kinds = {"rand": lambda: bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz") for _ in range(int(2 ** rng.uniform(5, 9)))),
             "punct": lambda: bytes([rng.choice(b"=-#")]) * rng.randint(2, 256),
             "mixed": lambda: bytes(rng.choice(b"aeiotnsr ETAOIN0123.,-_") for _ in range(rng.randint(2, 200))),
             "utf8": lambda: "".join(chr(rng.choice([rng.randint(0x4E00, 0x9FA5), rng.randint(0x3041, 0x3096), rng.randint(0xAC00, 0xD7A3)])) for _ in range(rng.randint(1, 24))).encode()}
    for kind, gen in kinds.items():
        tr = ts = 0
        for _ in range(npieces):
            piece = gen()
            want = byte_pair_encode(piece, enc)
            got, rounds = batched_bpe(piece, enc, refine)
            assert got == want, (kind, piece, got, want)
            tr += rounds
            ts += len(piece) - len(want) if len(want) > 1 or len(piece) == 1 else 0
        print(f"{name} {kind}: {npieces} pieces exact; rounds batched {tr}, sequential merges {ts}, ratio {ts / max(tr, 1):.2f}")


if __name__ == "__main__":
    main()
