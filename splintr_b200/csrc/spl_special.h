// Special-token matches of one document for ARBITRARY special-token sets (host + device).
//
// The reference scans with aho-corasick 1.1, default MatchKind::Standard, non-overlapping find_iter
// (/root/reference/src/core/tokenizer.rs:429-434 builds the automaton from whatever strings the user passes, :851
// iterates).  Its published semantics: from the cursor on, the automaton reports the first position at which some
// pattern ENDS; of the patterns ending there it reports the state's own one first, i.e. the longest (smallest start);
// the search then restarts at that end with the automaton in its start state.  So a match is
//     the occurrence with the earliest end among all occurrences that start at or after the cursor;
//     ties on the end go to the smallest start.
// When no special string contains or overlaps another one (all seven bundled sets) every occurrence is a match and
// k_mark_specials marks them independently.  Otherwise occurrences compete, the choice depends on the cursor, and the
// document is walked sequentially -- but only from candidate start to candidate start (the `cand` bitmap a parallel
// pass has filled: positions where some special string occurs inside the document), which are rare.
#pragma once
#include "spl_common.h"

struct SplSpecialSet {
    const uint8_t* bytes;        // concatenated strings
    const uint32_t* off;         // [n + 1]
    uint32_t n;
};

template <class Text>
SPL_HD bool spl_special_match(const Text& t, const SplSpecialSet& S, uint32_t k, uint32_t i, uint32_t limit) {
    const uint32_t o = S.off[k], len = S.off[k + 1] - o;
    if (len > limit - i) return false;
    for (uint32_t j = 0; j < len; ++j)
        if (t.byte(i + j) != S.bytes[o + j]) return false;
    return true;
}

// Document [d0, d1).  cand.next(from, lim): the first position p with from <= p < lim at which some special string
// occurs (entirely inside the document), or lim.  emit(start, end, k) is called for every match, in order.
template <class Text, class Cand, class Emit>
SPL_HD void spl_special_walk(const Text& t, const Cand& cand, const SplSpecialSet& S, uint32_t d0, uint32_t d1, Emit emit) {
    uint32_t pos = d0;
    while (pos < d1) {
        const uint32_t i0 = cand.next(pos, d1);
        if (i0 >= d1) return;
        // the earliest end of any occurrence that starts at or after i0: only starts before that end can beat it
        uint32_t E = d1 + 1u;
        for (uint32_t i = i0; i < E && i < d1; i = cand.next(i + 1u, d1)) {
            for (uint32_t k = 0; k < S.n; ++k) {
                const uint32_t len = S.off[k + 1] - S.off[k];
                if (i + len < E && spl_special_match(t, S, k, i, d1)) E = i + len;
            }
        }
        if (E > d1) return;                                    // (cannot happen: i0 is a candidate)
        // of the occurrences that end there, the one with the smallest start
        bool found = false;
        for (uint32_t i = i0; i < E && !found; i = cand.next(i + 1u, d1)) {
            for (uint32_t k = 0; k < S.n; ++k) {
                if (S.off[k + 1] - S.off[k] == E - i && spl_special_match(t, S, k, i, d1)) { emit(i, E, k); found = true; break; }
            }
        }
        pos = E;
    }
}
