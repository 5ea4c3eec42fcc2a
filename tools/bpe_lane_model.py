"""Lane-level model of one k_bpe group (spl_encode.cu, bpe_group<LG>): the same row bitmaps, ballots, carries and
shuffled masks as the kernel, one Python loop iteration per lane, checked against the oracle's sequential loop.
Usage: python tools/bpe_lane_model.py [vocab] [pieces per kind]"""
import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.py_oracle import byte_pair_encode, load_tiktoken_bpe
from splintr_b200.presets import get_preset, load_vocab_bytes

NONE = 0x1FFFFF
M32 = 0xFFFFFFFF


def clz(x):
    return 32 - x.bit_length()


def ffs(x):
    return (x & -x).bit_length()


def group_bpe(piece, enc, dec, LG):
    """enc: bytes -> rank, dec: rank -> bytes; symbols are ranks (single bytes too)."""
    G = 1 << LG
    GM = M32 if G == 32 else (1 << G) - 1
    n = len(piece)
    L = n
    S = {}
    K = {}
    X = {}

    def lookup(a, b):
        return enc.get(dec[a] + dec[b], NONE)
    for i in range(n):
        S[i] = enc[piece[i:i + 1]]
    for i in range(n):
        K[i] = lookup(S[i], S[i + 1]) if i + 1 < L else NONE
    rounds = 0

    def shfl(vals, src):
        return [vals[src[g]] for g in range(G)]

    def next_mask(vals, d):
        return [(vals[(g + d) & (G - 1)] >> ((g + d) >> LG)) & M32 for g in range(G)]

    def prev_mask(vals):
        return [(vals[(g + G - 1) & (G - 1)] << (0 if g else 1)) & M32 for g in range(G)]

    while True:
        rowsW = (L + G - 1) >> LG
        assert rowsW <= 32
        fV = [0] * G; fDL = [0] * G; fDR = [0] * G; fPK = [0] * G
        gmin = M32
        for g in range(G):
            for r in range(rowsW):
                i = (r << LG) + g
                if i + 1 < L:
                    kc = K[i]
                    if kc != NONE:
                        kl = K[i - 1] if i else NONE
                        kr = K[i + 1] if i + 2 < L else NONE
                        lt, rt = kl <= kc, kr < kc
                        bit = 1 << r
                        if lt:
                            if rt: fPK[g] |= bit
                            else: fDL[g] |= bit
                        else:
                            if rt: fDR[g] |= bit
                            else: fV[g] |= bit
                        gmin = min(gmin, (kc << 11) | i)
        if gmin == M32:
            break
        rounds += 1
        m = list(fV)
        if any(fDL):
            carry = 0
            for r in range(rowsW):
                bit = 1 << r
                bDL = sum(1 << g for g in range(G) if fDL[g] & bit) & GM
                for g in range(G):
                    if fDL[g] & bit:
                        below = ~bDL & ((1 << g) - 1) & M32
                        mv = (((g - (31 - clz(below))) & 1) ^ 1) if below else carry ^ ((g + 1) & 1)
                        if mv: m[g] |= bit
                carry = (m[G - 1] >> r) & 1
        if any(fDR):
            carry = 0
            for r in range(rowsW - 1, -1, -1):
                bit = 1 << r
                bDR = sum(1 << g for g in range(G) if fDR[g] & bit) & GM
                for g in range(G):
                    if fDR[g] & bit:
                        above = ~bDR & GM & ~(((2 << g) & M32) - 1) & M32
                        mv = (((ffs(above) - 1 - g) & 1) ^ 1) if above else carry ^ ((G - g) & 1)
                        if mv: m[g] |= bit
                carry = (m[0] >> r) & 1
        mL, mR = prev_mask(m), next_mask(m, 1)
        m = [m[g] | (fPK[g] & ~mL[g] & ~mR[g] & M32) for g in range(G)]
        m2 = next_mask(m, 2)
        theta = NONE
        for g in range(G):
            mm = m[g]
            while mm:
                r = ffs(mm) - 1; i = (r << LG) + g
                mm &= mm - 1
                tm = K[i]
                hasL, hasR = i > 0, i + 2 < L
                hasC = hasR and ((m2[g] >> r) & 1)
                ra = lookup(S[i - 1], tm) if hasL else NONE
                rb = lookup(tm, S[i + 2]) if hasR else NONE
                rc = lookup(tm, K[i + 2]) if hasC else NONE
                theta = min(theta, ra, rb, rc)
                X[i] = (ra | (rc << 21)) & M32
                X[i + 1] = (rb | ((rc >> 11) << 21)) & M32
        cm = [0] * G
        for g in range(G):
            mm = m[g]
            while mm:
                r = ffs(mm) - 1
                mm &= mm - 1
                kc = K[(r << LG) + g]
                if kc < theta or ((kc << 11) | ((r << LG) + g)) == gmin:
                    cm[g] |= 1 << r
        cmL, cmR1, cmR2 = prev_mask(cm), next_mask(cm, 1), next_mask(cm, 2)
        base = 0
        for r in range(rowsW):
            new = []
            b = 0
            for g in range(G):
                i = (r << LG) + g
                surv = i < L and not ((cmL[g] >> r) & 1)
                if surv:
                    if (cm[g] >> r) & 1:
                        x0, x1 = X[i], X[i + 1]
                        s_new = K[i]
                        k_new = ((x0 >> 21) | ((x1 >> 21) << 11)) if ((cmR2[g] >> r) & 1) else (x1 & NONE)
                    else:
                        s_new = S[i]
                        k_new = (X[i + 1] & NONE) if ((cmR1[g] >> r) & 1) else (K[i] if i + 1 < L else NONE)
                    b |= 1 << g
                    new.append((g, s_new, k_new))
            for g, s_new, k_new in new:
                pos = base + bin(b & ((1 << g) - 1)).count("1")
                assert pos <= (r << LG) + g
                S[pos] = s_new; K[pos] = k_new
            base += bin(b).count("1")
        L = base
    return [S[i] for i in range(L)], rounds


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cl100k_base"
    npieces = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    p = get_preset(name)
    enc = load_tiktoken_bpe(load_vocab_bytes(p.vocab_file))
    dec = {v: k for k, v in enc.items()}
    rng = random.Random(11)
    kinds = {"rand": lambda hi: bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(2, hi))),
             "punct": lambda hi: bytes([rng.choice(b"=-# \n")]) * rng.randint(2, hi),
             "two": lambda hi: bytes(rng.choice(b"ab") for _ in range(rng.randint(2, hi))),
             "mixed": lambda hi: bytes(rng.choice(b"aeiotnsr ETAOIN0123.,-_") for _ in range(rng.randint(2, hi))),
             "utf8": lambda hi: "".join(chr(rng.choice([rng.randint(0x4E00, 0x9FA5), rng.randint(0x3041, 0x3096), rng.randint(0xAC00, 0xD7A3)])) for _ in range(rng.randint(1, hi // 3))).encode()}
    for LG in range(6):
        hi = 32 << LG
        for kind, gen in kinds.items():
            for _ in range(npieces):
                piece = gen(hi)
                if piece in enc:
                    continue
                want = byte_pair_encode(piece, enc)
                got, _ = group_bpe(piece, enc, dec, LG)
                assert got == want, (LG, kind, piece, got, want)
        print(f"{name} LG={LG}: lane model exact on {npieces} pieces x {len(kinds)} kinds (lengths <= {hi})")


if __name__ == "__main__":
    main()
