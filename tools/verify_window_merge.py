#!/usr/bin/env python3
"""Empirical check of the windowed-minimum parallel merge rule used for very long pre-tokens.

Sequential reference order (src/bpe.rs:104-153): repeatedly merge the globally lowest-rank pair, leftmost
on ties, one at a time.  Parallel rule: in one round merge EVERY pair whose (rank, position) is the strict
minimum among all pairs within D symbols on either side, D = length of the longest token (in initial
symbols).  Claim: same final ids for ANY merge table (monotone or not)."""
import random, sys

def seq(toks, ranks, new):
    toks = list(toks)
    while True:
        best = None
        for i in range(len(toks) - 1):
            r = ranks.get((toks[i], toks[i + 1]))
            if r is not None and (best is None or r < best[0]):
                best = (r, i)
        if best is None:
            return toks
        i = best[1]
        toks[i:i + 2] = [new[(toks[i], toks[i + 1])]]

def par(toks, ranks, new, D):
    toks = list(toks)
    rounds = 0
    while True:
        n = len(toks)
        rk = [ranks.get((toks[i], toks[i + 1])) for i in range(n - 1)]
        sel = []
        for i, r in enumerate(rk):
            if r is None:
                continue
            ok = True
            for j in range(max(0, i - D), min(n - 1, i + D + 1)):
                if j != i and rk[j] is not None and (rk[j], j) < (r, i):
                    ok = False
                    break
            if ok:
                sel.append(i)
        if not sel:
            return toks, rounds
        rounds += 1
        out, i, s = [], 0, set(sel)
        while i < n:
            if i in s:
                out.append(new[(toks[i], toks[i + 1])]); i += 2
            else:
                out.append(toks[i]); i += 1
        toks = out

def trial(rng, nsym, lmax, nmerge, textlen, d_slack=0):
    # vocab: strings over nsym letters; a merge (x, y) is legal iff x, y, x+y are all tokens
    base = [chr(97 + k) for k in range(nsym)]
    vocab = set(base)
    for _ in range(nmerge * 3):
        L = rng.randint(2, lmax)
        vocab.add(''.join(rng.choice(base) for _ in range(L)))
    pairs = []
    for t in vocab:
        for c in range(1, len(t)):
            if t[:c] in vocab and t[c:] in vocab:
                pairs.append((t[:c], t[c:]))
    rng.shuffle(pairs)
    pairs = pairs[:nmerge]
    ranks = {p: r for r, p in enumerate(pairs)}        # arbitrary order: NOT monotone in general
    new = {p: p[0] + p[1] for p in pairs}
    D = max(len(t) for t in vocab) + d_slack
    bad = 0
    for _ in range(20):
        text = [rng.choice(base) for _ in range(rng.randint(1, textlen))]
        a = seq(text, ranks, new)
        b, rounds = par(text, ranks, new, D)
        if a != b:
            bad += 1
    return bad

if __name__ == '__main__':
    rng = random.Random(7)
    tot = bad = 0
    for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3000):
        nsym = rng.choice([2, 2, 3, 4]); lmax = rng.choice([2, 3, 4, 5]); nm = rng.choice([3, 6, 12, 30])
        bad += trial(rng, nsym, lmax, nm, 120); tot += 20
    print('window = longest token:      mismatches %d / %d' % (bad, tot))
    # sanity: the rule must break with a window that is too small (window 1 = "local minima")
    rng = random.Random(7); bad1 = 0
    for it in range(600):
        nsym = rng.choice([2, 2, 3, 4]); lmax = rng.choice([3, 4, 5]); nm = rng.choice([6, 12, 30])
        bad1 += trial(rng, nsym, lmax, nm, 120, d_slack=-lmax - 10 + 1 + 10 - lmax if False else -(lmax - 1))
    print('window = 1 (local minima):   mismatches %d / %d (expected > 0)' % (bad1, 600 * 20))
