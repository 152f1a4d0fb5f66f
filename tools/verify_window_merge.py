#!/usr/bin/env python3
"""Empirical check of the round-parallel merge rule used for very long pre-tokens (csrc/encode_xlong.cu).

Sequential reference order (src/bpe.rs:104-153): repeatedly merge the globally lowest-rank pair, leftmost
on ties, one at a time.

Round-parallel rule.  Pair j = (sym[j], sym[j+1]) with rank r[j].  W = length of the longest token in
initial symbols.  In one round, simultaneously merge every pair j that is SELECTED:
    blocked(j)   some pair within W symbols of j (either side) has a strictly lower rank
    in a run     r[j-1] == r[j] (then sym[j-1] == sym[j] == sym[j+1]); s = first pair of the run
    selected(j)  not in a run:  not blocked(j)
                 in a run:      (j - s) even and no pair of s..j is blocked     (greedy left-to-right pairing)
Claim: identical final ids for every MONOTONE merge table (every pair that contains a merged token has a
higher rank than every merge producing that token -- true of any table a BPE trainer emits without duplicate
products; the loader checks it and keeps the sequential path otherwise).  The proof sketch is in DESIGN.md;
this tool fuzzes it, and shows that the rule does break for non-monotone tables and for a window of 1."""
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

INF = 1 << 60

def par(toks, ranks, new, W):
    toks = list(toks)
    rounds = 0
    while True:
        n = len(toks)
        rk = [ranks.get((toks[i], toks[i + 1]), INF) for i in range(n - 1)]
        blocked = [False] * (n - 1)
        for i, r in enumerate(rk):
            lo, hi = max(0, i - W), min(n - 1, i + W + 1)
            blocked[i] = min(rk[lo:hi]) < r
        sel = set()
        s = 0; bad = False
        for i, r in enumerate(rk):
            if i == 0 or rk[i - 1] != r:
                s = i; bad = False
            bad = bad or blocked[i]
            if r != INF and not bad and (i - s) % 2 == 0:
                sel.add(i)
        if not sel:
            return toks, rounds
        rounds += 1
        out, i = [], 0
        while i < n:
            if i in sel:
                out.append(new[(toks[i], toks[i + 1])]); i += 2
            else:
                out.append(toks[i]); i += 1
        toks = out

def reaches(vocab_tokens):
    """WL[t] = how far (in initial symbols) a token that ENDS with t can extend to the left of t,
    WR[t] = how far a token that STARTS with t can extend to the right of t (csrc/loader.cpp: token_reach)."""
    toks = set(vocab_tokens)
    WL = {t: 0 for t in toks}; WR = {t: 0 for t in toks}
    for t in toks:
        for c in range(1, len(t)):
            p_, s_ = t[:c], t[c:]
            if p_ in toks: WR[p_] = max(WR[p_], len(t) - len(p_))
            if s_ in toks: WL[s_] = max(WL[s_], len(t) - len(s_))
    return WL, WR

def par_reach(toks, ranks, new, WL, WR, in_symbols=False):
    """Same rounds, but every pair only looks as far as tokens around its own two symbols can reach:
    left of x = sym[j] up to WL[x] initial symbols, right of y = sym[j+1] up to WR[y]."""
    toks = list(toks)
    rounds = 0
    while True:
        n = len(toks)
        pos = [0] * (n + 1)
        for i, t in enumerate(toks): pos[i + 1] = pos[i] + (1 if in_symbols else len(t))   # in_symbols: the grid-wide kernels' conservative unit
        rk = [ranks.get((toks[i], toks[i + 1]), INF) for i in range(n - 1)]
        blocked = [False] * (n - 1)
        for j, r in enumerate(rk):
            if r == INF: continue
            x, y = toks[j], toks[j + 1]
            k = j - 1
            while k >= 0 and pos[j] - pos[k] <= WL[x]:
                if rk[k] < r: blocked[j] = True; break
                k -= 1
            k = j + 1
            while not blocked[j] and k <= n - 2 and pos[k + 2] - pos[j + 2] <= WR[y]:
                if rk[k] < r: blocked[j] = True; break
                k += 1
        sel = set(); s0 = 0; bad = False
        for i, r in enumerate(rk):
            if i == 0 or rk[i - 1] != r:
                s0 = i; bad = False
            bad = bad or blocked[i]
            if r != INF and not bad and (i - s0) % 2 == 0:
                sel.add(i)
        if not sel:
            return toks, rounds
        rounds += 1
        out, i = [], 0
        while i < n:
            if i in sel:
                out.append(new[(toks[i], toks[i + 1])]); i += 2
            else:
                out.append(toks[i]); i += 1
        toks = out

def make_table(rng, nsym, lmax, nmerge, monotone):
    base = [chr(97 + k) for k in range(nsym)]
    vocab = set(base)
    for _ in range(nmerge * 3):
        L = rng.randint(2, lmax)
        vocab.add(''.join(rng.choice(base) for _ in range(L)))
    if monotone and rng.random() < 0.5:                     # closed vocab: more deep merge chains
        for t in list(vocab):
            for c in range(2, len(t)):
                vocab.add(t[:c])
    pairs = []
    for t in vocab:
        for c in range(1, len(t)):
            if t[:c] in vocab and t[c:] in vocab:
                pairs.append((t[:c], t[c:]))
    rng.shuffle(pairs)
    pairs = pairs[:nmerge]
    if monotone:
        # keep one producer per token, order so that components are produced before they are used
        seen, uniq = set(), []
        for p in pairs:
            if p[0] + p[1] not in seen:
                seen.add(p[0] + p[1]); uniq.append(p)
        prod = {p[0] + p[1] for p in uniq}
        order, done, pending = [], set(base), list(uniq)
        while pending:
            ready = [p for p in pending if (p[0] in done or p[0] not in prod) and (p[1] in done or p[1] not in prod)]
            if not ready:
                break
            p = rng.choice(ready)
            order.append(p); done.add(p[0] + p[1]); pending.remove(p)
        # merges whose components are never produced can never fire; dropping them keeps the table monotone
        pairs = [p for p in order if all(len(x) == 1 or x in done for x in p)]
    ranks = {p: r for r, p in enumerate(pairs)}
    newt = {p: p[0] + p[1] for p in pairs}
    W = max(len(t) for t in vocab)
    make_table.last_vocab = set(base) | set(newt.values()) | {x for p in pairs for x in p}
    return base, ranks, newt, W

def trial(rng, nsym, lmax, nmerge, textlen, monotone, W_override=None, reach=False, in_symbols=False):
    base, ranks, newt, W = make_table(rng, nsym, lmax, nmerge, monotone)
    if reach:
        WL, WR = reaches(make_table.last_vocab)
    bad = 0; rmax = 0
    for _ in range(20):
        if rng.random() < 0.3:                              # runs of one symbol: the parity rule
            text = []
            while len(text) < textlen:
                text += [rng.choice(base)] * rng.randint(1, 12)
        else:
            text = [rng.choice(base) for _ in range(rng.randint(1, textlen))]
        a = seq(text, ranks, newt)
        b, rounds = par_reach(text, ranks, newt, WL, WR, in_symbols) if reach else par(text, ranks, newt, W_override or W)
        rmax = max(rmax, rounds)
        if a != b:
            bad += 1
    return bad, rmax

if __name__ == '__main__':
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    for label, mono, wo, reach in (('monotone table, window = longest token', True, None, False),
                                   ('monotone table, per-symbol reach windows', True, None, True),
                                   ('monotone table, reach counted in symbols', True, None, 'sym'),
                                   ('NON-monotone table (expected > 0)', False, None, False),
                                   ('monotone table, window = 1 (expected > 0)', True, 1, False)):
        rng = random.Random(7); tot = bad = 0; rmax = 0
        for it in range(N):
            nsym = rng.choice([1, 2, 2, 3, 4]); lmax = rng.choice([2, 3, 4, 5, 6]); nm = rng.choice([3, 6, 12, 30, 60])
            b, r = trial(rng, nsym, lmax, nm, 120, mono, wo, bool(reach), reach == 'sym'); bad += b; tot += 20; rmax = max(rmax, r)
        print('%-46s mismatches %d / %d   (most rounds %d)' % (label, bad, tot, rmax))
