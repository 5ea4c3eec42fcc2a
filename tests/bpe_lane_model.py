"""Lane-level model of one k_bpe_long group (spl_encode.cu, bpe_group<LG>): the same per-lane bitmasks over consecutive
parts, carry-ripple run masks (runs_above / slope_m, bit reversal for the right-hand slopes), boundary bits between
lanes, packed probe results and compaction as the kernel, checked against the oracle's sequential loop.
Usage: python tests/bpe_lane_model.py [vocab] [pieces per kind]"""
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


def brev(x):
    return int(format(x & M32, "032b")[::-1], 2)


def runs_above(D, st):
    """spl_encode.cu runs_above: the runs of D that start right above the bits of st (the addition ripples through them)."""
    return D & ~((D + ((st << 1) & M32)) & M32) & M32


def slope_m(D, V, cin, first):
    """spl_encode.cu slope_m: m along upward slopes; a run that starts at bit `first` continues a foreign slope (m = cin below it)."""
    E = 0x55555555
    run0 = D & ~((D + first) & M32) & M32
    P = (~E & M32) if (first & E) else E
    return (runs_above(D, V & E) & E) | (runs_above(D, V & ~E & M32) & ~E & M32) | (run0 & (P if cin else (~P & M32)))


def group_bpe(piece, enc, dec, LG):
    """One group of G = 2^LG lanes of bpe_group<LG> (spl_encode.cu): lane g owns the parts [g * B, g * B + B), its flags
    are bitmasks over them.  enc: bytes -> rank, dec: rank -> bytes; symbols are ranks (single bytes too)."""
    G = 1 << LG
    L = len(piece)
    S = {}
    K = {}
    X = {}

    def lookup(a, b):
        return enc.get(dec[a] + dec[b], NONE)
    for i in range(L):
        S[i] = enc[piece[i:i + 1]]
    for i in range(L):
        K[i] = lookup(S[i], S[i + 1]) if i + 1 < L else NONE
    rounds = 0
    while True:
        B = max(2, (L + G - 1) >> LG)
        assert B <= 32
        e0 = [g * B for g in range(G)]
        nv = [min(B, L - e0[g]) if e0[g] < L else 0 for g in range(G)]
        LV = [0] * G; LT = [0] * G; RT = [0] * G
        gmin = M32
        for g in range(G):
            prev = K[e0[g] - 1] if (e0[g] and e0[g] < L) else NONE
            cur = K[e0[g]] if e0[g] < L else NONE
            for j in range(nv[g]):
                e = e0[g] + j
                nxt = K.get(e + 1, 12345)                     # beyond L - 1: never used (K[L - 1] is NONE)
                if cur != NONE:
                    bit = 1 << j
                    LV[g] |= bit
                    if prev <= cur: LT[g] |= bit
                    if nxt < cur: RT[g] |= bit
                    gmin = min(gmin, (cur << 11) | e)
                prev, cur = cur, nxt
        if gmin == M32:
            break
        rounds += 1
        fV = [LV[g] & ~LT[g] & ~RT[g] & M32 for g in range(G)]
        fDL = [LV[g] & LT[g] & ~RT[g] & M32 for g in range(G)]
        fDR = [LV[g] & RT[g] & ~LT[g] & M32 for g in range(G)]
        fPK = [LV[g] & LT[g] & RT[g] for g in range(G)]
        cin = [0] * G; cin2 = [0] * G
        firstR = 1 << (32 - B)
        while True:
            m = [fV[g] | slope_m(fDL[g], fV[g], cin[g], 1) | brev(slope_m(brev(fDR[g]), brev(fV[g]), cin2[g], firstR)) for g in range(G)]
            ncin = [((m[g - 1] >> (B - 1)) & 1) if g else 0 for g in range(G)]
            ncin2 = [(m[g + 1] & 1) if g + 1 < G else 0 for g in range(G)]
            ch = any((ncin[g] != cin[g] and (fDL[g] & 1)) or (ncin2[g] != cin2[g] and ((fDR[g] >> (B - 1)) & 1)) for g in range(G))
            cin, cin2 = ncin, ncin2
            if not ch:
                break
        m = [m[g] | (fPK[g] & ~((m[g] << 1) | cin[g]) & ~((m[g] >> 1) | (cin2[g] << (B - 1))) & M32) for g in range(G)]
        theta = NONE
        for g in range(G):
            dn = m[g + 1] if g + 1 < G else 0
            m2 = ((m[g] | (dn << B)) >> 2) & M32
            mm = m[g]
            while mm:
                j = ffs(mm) - 1
                mm &= mm - 1
                e = e0[g] + j
                tm = K[e]
                hasL, hasR, hasC = e > 0, e + 2 < L, (m2 >> j) & 1
                ra = lookup(S[e - 1], tm) if hasL else NONE
                rb = lookup(tm, S[e + 2]) if hasR else NONE
                rc = lookup(tm, K[e + 2]) if hasC else NONE
                theta = min(theta, ra, rb, rc)
                X[e] = (ra | (rc << 21)) & M32
                X[e + 1] = (rb | ((rc >> 11) << 21)) & M32
        cm = [0] * G
        for g in range(G):
            mm = m[g]
            while mm:
                j = ffs(mm) - 1
                mm &= mm - 1
                kc = K[e0[g] + j]
                if kc < theta or ((kc << 11) | (e0[g] + j)) == gmin:
                    cm[g] |= 1 << j
        surv = [0] * G
        for g in range(G):
            up = cm[g - 1] if g else 0
            dn = cm[g + 1] if g + 1 < G else 0
            cw = cm[g] | (dn << B)
            cm1, cm2 = (cw >> 1) & M32, (cw >> 2) & M32
            ex = M32 if nv[g] >= 32 else (1 << nv[g]) - 1
            surv[g] = ex & ~((cm[g] << 1) | (((up >> (B - 1)) & 1) if g else 0)) & M32
            mm = surv[g] & (cm[g] | cm1)
            while mm:
                j = ffs(mm) - 1
                mm &= mm - 1
                e = e0[g] + j
                x1 = X[e + 1]
                if (cm[g] >> j) & 1:
                    x0 = X[e]
                    S[e] = K[e]
                    K[e] = ((x0 >> 21) | ((x1 >> 21) << 11)) if ((cm2 >> j) & 1) else (x1 & NONE)
                else:
                    K[e] = x1 & NONE
        newS, newK, d = {}, {}, 0
        for g in range(G):                                    # lane bases = prefix of popc(surv) over the group
            mm = surv[g]
            while mm:
                j = ffs(mm) - 1
                mm &= mm - 1
                newS[d] = S[e0[g] + j]; newK[d] = K[e0[g] + j]
                d += 1
        S, K, X, L = newS, newK, {}, d
        assert K[L - 1] == NONE
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
