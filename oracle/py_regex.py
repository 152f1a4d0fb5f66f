"""TEST INFRASTRUCTURE (oracle): the `Split` pre-tokenizer of the reference, restated for the CPU.

The reference compiles the pattern with the third-party crate `regex` (range "1.10" in Cargo.toml; no Cargo.lock, so no pinned
patch version) and walks `find_iter` (src/pretokenizers.rs:298-433).  The crate is not under /root/reference; its published
semantics are restated here for the subset of its syntax this project supports:

  * leftmost-first matching (Perl-like: the first alternative / the greedy choice that leads to a match wins), which for a
    backtracking matcher is simply "the first success in priority order";
  * Unicode classes: \\p{General_Category}, \\s = White_Space, \\d = Nd, \\w = Alphabetic | M | Nd | Pc | Join_Control,
    a few scripts; `.` = any char but \\n; bracket classes with ranges, negation, nesting and POSIX names;
  * `find_iter`: successive non-overlapping matches, each search starting where the previous match ended.

Outside the subset (-> Unsupported, which the loader mirrors with CTK_ERR_UNSUPPORTED): flags, anchors and word boundaries,
class set operations, patterns that can match the empty string, repetition of a nullable operand.  Look-around and
back-references make Regex::new FAIL in the crate: the reference then passes the text through (pretokenizers.rs:299-302).

The matcher is a backtracking one over the AST -- deliberately a different algorithm from the product's (ordered-subset DFA
built on the host, csrc/regex_dfa.cpp), so that agreement between the two means something.  tests/test_split_oracle.py pins
it against Python's `regex` module on patterns of the subset.
"""
import bisect
import sys
import unicodedata as ud

import unicode_props_gen as UP

sys.setrecursionlimit(max(sys.getrecursionlimit(), 20000))

MAXCP = 0x10FFFF
WHITE_SPACE = [(0x09, 0x0D), (0x20, 0x20), (0x85, 0x85), (0xA0, 0xA0), (0x1680, 0x1680), (0x2000, 0x200A), (0x2028, 0x2029),
               (0x202F, 0x202F), (0x205F, 0x205F), (0x3000, 0x3000)]
GC_GROUPS = {'L': 'Lu Ll Lt Lm Lo', 'M': 'Mn Mc Me', 'N': 'Nd Nl No', 'P': 'Pc Pd Ps Pe Pi Pf Po', 'S': 'Sm Sc Sk So',
             'Z': 'Zs Zl Zp', 'C': 'Cc Cf Cs Co Cn', 'LC': 'Lu Ll Lt'}
GC_LONG = {'Letter': 'L', 'Mark': 'M', 'Number': 'N', 'Punctuation': 'P', 'Symbol': 'S', 'Separator': 'Z', 'Other': 'C',
           'Uppercase_Letter': 'Lu', 'Lowercase_Letter': 'Ll', 'Titlecase_Letter': 'Lt', 'Modifier_Letter': 'Lm', 'Other_Letter': 'Lo',
           'Decimal_Number': 'Nd', 'Letter_Number': 'Nl', 'Other_Number': 'No', 'Nonspacing_Mark': 'Mn', 'Spacing_Mark': 'Mc',
           'Enclosing_Mark': 'Me', 'Cased_Letter': 'LC'}
GC_ALL = set('Lu Ll Lt Lm Lo Mn Mc Me Nd Nl No Pc Pd Ps Pe Pi Pf Po Sm Sc Sk So Zs Zl Zp Cc Cf Cs Co Cn'.split())
POSIX = {'alnum': '0-9A-Za-z', 'alpha': 'A-Za-z', 'ascii': '\x00-\x7f', 'blank': '\t ', 'cntrl': '\x00-\x1f\x7f', 'digit': '0-9',
         'graph': '!-~', 'lower': 'a-z', 'print': ' -~', 'punct': '!-/:-@[-`{-~', 'space': '\t\n\x0b\x0c\r ', 'upper': 'A-Z',
         'word': '0-9A-Za-z_', 'xdigit': '0-9A-Fa-f'}


class Unsupported(Exception):
    """the pattern is (or may be) valid for the crate but outside the supported subset"""


class CharSet:
    """a set of code points: union of ranges, General_Category values and named properties, optionally negated"""

    def __init__(self):
        self.ranges = []            # (lo, hi)
        self.cats = set()           # General_Category values
        self.props = []             # sorted range lists (Alphabetic, scripts)
        self.subs = []              # nested CharSets (union)
        self.neg = False
        self._memo = {}

    def has(self, cp):
        r = self._memo.get(cp)
        if r is None:
            r = self._has(cp) != self.neg
            self._memo[cp] = r
        return r

    def _has(self, cp):
        for lo, hi in self.ranges:
            if lo <= cp <= hi:
                return True
        if self.cats and ud.category(chr(cp)) in self.cats:
            return True
        for p in self.props:
            i = bisect.bisect_right(p, (cp, MAXCP + 1)) - 1
            if i >= 0 and p[i][0] <= cp <= p[i][1]:
                return True
        return any(s.has(cp) for s in self.subs)


def _prop_set(name, negate):
    """\\p{name}"""
    cs = CharSet()
    cs.neg = negate
    if name.startswith('^'):
        cs.neg = not cs.neg
        name = name[1:]
    for pre in ('gc=', 'General_Category=', 'sc=', 'Script=', 'scx='):
        if name.startswith(pre):
            if pre == 'scx=':
                raise Unsupported('Script_Extensions')
            name = name[len(pre):]
            break
    name = GC_LONG.get(name, name)
    if name in GC_GROUPS:
        cs.cats = set(GC_GROUPS[name].split())
    elif name in GC_ALL:
        cs.cats = {name}
    elif name in ('White_Space', 'space', 'Whitespace'):
        cs.ranges = list(WHITE_SPACE)
    elif name == 'Alphabetic':
        cs.props = [UP.ALPHABETIC]
    elif name in UP.SCRIPTS:
        cs.props = [UP.SCRIPTS[name]]
    elif name == 'Any':
        cs.ranges = [(0, MAXCP)]
    else:
        raise Unsupported('unicode property ' + name)
    return cs


def _perl_set(c):
    cs = CharSet()
    cs.neg = c.isupper()
    k = c.lower()
    if k == 's':
        cs.ranges = list(WHITE_SPACE)
    elif k == 'd':
        cs.cats = {'Nd'}
    else:                                                    # \w
        cs.props = [UP.ALPHABETIC]
        cs.cats = {'Mn', 'Mc', 'Me', 'Nd', 'Pc'}
        cs.ranges = [(0x200C, 0x200D)]
    return cs


# ---- AST: ('set', CharSet) | ('cat', [nodes]) | ('alt', [nodes]) | ('rep', node, min, max|None, greedy)

class _Parser:
    def __init__(self, pat):
        self.p = pat
        self.i = 0
        self.n_nodes = 0

    def peek(self, k=0):
        j = self.i + k
        return self.p[j] if j < len(self.p) else ''

    def parse(self):
        node = self.alternation()
        if self.i != len(self.p):
            raise Unsupported('unbalanced )')
        return node

    def alternation(self):
        branches = [self.concat()]
        while self.peek() == '|':
            self.i += 1
            branches.append(self.concat())
        return branches[0] if len(branches) == 1 else ('alt', branches)

    def concat(self):
        items = []
        while self.i < len(self.p) and self.peek() not in '|)':
            items.append(self.repeat())
        return ('cat', items)

    def repeat(self):
        atom = self.atom()
        while True:
            c = self.peek()
            if c in ('*', '+', '?') and c:
                self.i += 1
                lo, hi = {'*': (0, None), '+': (1, None), '?': (0, 1)}[c]
            elif c == '{':
                j = self.p.find('}', self.i)
                body = self.p[self.i + 1:j] if j > 0 else ''
                parts = body.split(',')
                if j < 0 or not 1 <= len(parts) <= 2 or not parts[0].strip().isdigit() or (len(parts) == 2 and parts[1].strip() and not parts[1].strip().isdigit()) \
                        or not body.isascii():
                    raise Unsupported('counted repetition')
                lo = int(parts[0])
                hi = lo if len(parts) == 1 else (int(parts[1]) if parts[1].strip() else None)
                if hi is not None and hi < lo:
                    raise Unsupported('counted repetition bounds')
                if lo > 1000 or (hi or 0) > 1000:
                    raise Unsupported('counted repetition too large')
                self.i = j + 1
            else:
                return atom
            greedy = True
            if self.peek() == '?':
                self.i += 1
                greedy = False
            if nullable(atom) and (hi is None or hi > 1):
                raise Unsupported('repetition of an operand that can match the empty string')
            atom = ('rep', atom, lo, hi, greedy)

    def atom(self):
        c = self.peek()
        self.n_nodes += 1
        if c == '(':
            self.i += 1
            if self.peek() == '?':
                if self.peek(1) == ':':
                    self.i += 2
                elif self.peek(1) == 'P' and self.peek(2) == '<' or (self.peek(1) == '<' and self.peek(2) not in '=!'):
                    j = self.p.find('>', self.i)
                    if j < 0:
                        raise Unsupported('group name')
                    self.i = j + 1
                else:
                    raise Unsupported('flags or look-around')
            node = self.alternation()
            if self.peek() != ')':
                raise Unsupported('unclosed group')
            self.i += 1
            return node
        if c == '[':
            return ('set', self.bracket())
        if c == '.':
            self.i += 1
            cs = CharSet()
            cs.ranges = [(0x0A, 0x0A)]
            cs.neg = True
            return ('set', cs)
        if c == '\\':
            return ('set', self.escape(in_class=False))
        if c in '*+?{' or c in '^$' or c == '':
            raise Unsupported('operator without operand, or an anchor')
        self.i += 1
        cs = CharSet()
        cs.ranges = [(ord(c), ord(c))]
        return ('set', cs)

    def escape(self, in_class):
        """after a backslash: a CharSet (single code point or a class)"""
        self.i += 1
        c = self.peek()
        self.i += 1
        cs = CharSet()
        if c in 'dswDSW' and c:
            return _perl_set(c)
        if c in 'pP' and c:
            if self.peek() == '{':
                j = self.p.find('}', self.i)
                if j < 0:
                    raise Unsupported('unclosed \\p{')
                name = self.p[self.i + 1:j]
                self.i = j + 1
            else:
                name = self.peek()
                self.i += 1
            return _prop_set(name, c == 'P')
        simple = {'n': 10, 'r': 13, 't': 9, 'f': 12, 'v': 11, 'a': 7, '0': None}
        if c in simple and c != '0':
            cs.ranges = [(simple[c],) * 2]
            return cs
        if c in 'xuU' and c:
            if self.peek() == '{':
                j = self.p.find('}', self.i)
                hexs = self.p[self.i + 1:j] if j > 0 else ''
                self.i = j + 1
            else:
                w = {'x': 2, 'u': 4, 'U': 8}[c]
                hexs = self.p[self.i:self.i + w]
                self.i += w
                if len(hexs) != w:
                    raise Unsupported('hex escape')
            try:
                v = int(hexs, 16)
            except ValueError:
                raise Unsupported('hex escape')
            if not hexs.isascii() or v > MAXCP or 0xD800 <= v <= 0xDFFF:
                raise Unsupported('hex escape')
            cs.ranges = [(v, v)]
            return cs
        if c and c.isascii() and not c.isalnum() and c not in '<>':       # escaped punctuation is itself (\< \> are word boundaries)
            cs.ranges = [(ord(c), ord(c))]
            return cs
        raise Unsupported('escape \\' + c)

    def bracket(self):
        """[...] with ranges, negation, nesting, POSIX names; no set operations"""
        self.i += 1
        cs = CharSet()
        if self.peek() == '^':
            cs.neg = True
            self.i += 1
        first = True
        while True:
            c = self.peek()
            if c == '':
                raise Unsupported('unclosed class')
            if c == ']' and not first:
                self.i += 1
                return cs
            first = False
            if c == '[':
                if self.peek(1) == ':':
                    j = self.p.find(':]', self.i)
                    name = self.p[self.i + 2:j] if j > 0 else ''
                    neg = name.startswith('^')
                    if name.lstrip('^') not in POSIX:
                        raise Unsupported('POSIX class')
                    sub = CharSet()
                    sub.neg = neg
                    spec = POSIX[name.lstrip('^')]
                    k = 0
                    while k < len(spec):
                        if k + 2 < len(spec) and spec[k + 1] == '-':
                            sub.ranges.append((ord(spec[k]), ord(spec[k + 2])))
                            k += 3
                        else:
                            sub.ranges.append((ord(spec[k]), ord(spec[k])))
                            k += 1
                    cs.subs.append(sub)
                    self.i = j + 2
                    continue
                cs.subs.append(self.bracket())
                continue
            if c in '&-~' and self.peek(1) == c:
                raise Unsupported('class set operation')
            if c == '\\':
                item = self.escape(in_class=True)
                single = len(item.ranges) == 1 and item.ranges[0][0] == item.ranges[0][1] and not item.cats and not item.props and not item.neg
                if not single:
                    cs.subs.append(item)
                    continue
                lo = item.ranges[0][0]
            else:
                lo = ord(c)
                self.i += 1
            if self.peek() == '-' and self.peek(1) not in (']', ''):
                if self.peek(1) == '-':
                    raise Unsupported('class set operation')
                self.i += 1
                if self.peek() == '\\':
                    item = self.escape(in_class=True)
                    if not (len(item.ranges) == 1 and item.ranges[0][0] == item.ranges[0][1] and not item.cats and not item.props and not item.neg):
                        raise Unsupported('class range end')
                    hi = item.ranges[0][0]
                elif self.peek() == '[':
                    raise Unsupported('class range end')
                else:
                    hi = ord(self.peek())
                    self.i += 1
                if hi < lo:
                    raise Unsupported('class range out of order')
                cs.ranges.append((lo, hi))
            else:
                cs.ranges.append((lo, lo))


def nullable(node):
    k = node[0]
    if k == 'set':
        return False
    if k == 'cat':
        return all(nullable(x) for x in node[1])
    if k == 'alt':
        return any(nullable(x) for x in node[1])
    return node[2] == 0 or nullable(node[1])


def compile_pattern(pat):
    """-> AST of a supported pattern; raises Unsupported"""
    ps = _Parser(pat)
    node = ps.parse()
    if nullable(node):
        raise Unsupported('pattern can match the empty string')
    return node


def _match(node, s, i, cont):
    """end of the highest-priority match of node at s[i:] whose continuation succeeds, else None (s: list of code points)"""
    k = node[0]
    if k == 'set':
        if i < len(s) and node[1].has(s[i]):
            return cont(i + 1)
        return None
    if k == 'cat':
        items = node[1]

        def step(idx, j):
            if idx == len(items):
                return cont(j)
            return _match(items[idx], s, j, lambda e: step(idx + 1, e))
        return step(0, i)
    if k == 'alt':
        for b in node[1]:
            r = _match(b, s, i, cont)
            if r is not None:
                return r
        return None
    _, sub, lo, hi, greedy = node
    if sub[0] == 'set':                                       # the common case, without recursion per character
        cs = sub[1]
        m = 0
        lim = len(s) - i if hi is None else min(hi, len(s) - i)
        while m < lim and cs.has(s[i + m]):
            m += 1
        if m < lo:
            return None
        order = range(m, lo - 1, -1) if greedy else range(lo, m + 1)
        for c in order:
            r = cont(i + c)
            if r is not None:
                return r
        return None

    def rep(count, j):
        more = hi is None or count < hi

        def again():
            return _match(sub, s, j, lambda e: rep(count + 1, e) if e > j else None) if more else None
        if greedy:
            r = again()
            if r is not None:
                return r
            return cont(j) if count >= lo else None
        if count >= lo:
            r = cont(j)
            if r is not None:
                return r
        return again()
    return rep(0, i)


def find_iter(node, text):
    """regex::Regex::find_iter for a pattern that cannot match the empty string: [(start, end)] in characters"""
    s = [ord(c) for c in text]
    out = []
    pos = 0
    while pos < len(s):
        hit = None
        for st in range(pos, len(s)):
            e = _match(node, s, st, lambda j: j)
            if e is not None:
                hit = (st, e)
                break
        if hit is None:
            break
        out.append(hit)
        pos = hit[1]
    return out


BEHAVIORS = ('Removed', 'Isolated', 'MergedWithPrevious', 'MergedWithNext', 'Contiguous')


def split_with_behavior(node, text, behavior, invert):
    """regex_split_with_behavior, pretokenizers.rs:298-433, line by line (node = compiled pattern)"""
    matches = [(a, b) for a, b in find_iter(node, text)]
    if not matches:
        return [text]                                        # :305-307 (not filtered)
    result = []
    last_end = 0
    if behavior == 'Removed':                                # :313-331
        for a, b in matches:
            if invert:
                if a > last_end:
                    result.append(text[last_end:a])
            else:
                result.append(text[a:b])
            last_end = b
        if invert and last_end < len(text):
            result.append(text[last_end:])
    elif behavior == 'Isolated':                             # :332-347
        for a, b in matches:
            if a > last_end:
                result.append(text[last_end:a])
            result.append(text[a:b])
            last_end = b
        if last_end < len(text):
            result.append(text[last_end:])
    elif behavior == 'MergedWithPrevious':                   # :348-375
        for a, b in matches:
            if a > last_end:
                result.append(text[last_end:a] + text[a:b])
            elif result:
                result.append(result.pop() + text[a:b])
            else:
                result.append(text[a:b])
            last_end = b
        if last_end < len(text):
            result.append(text[last_end:])
    elif behavior == 'MergedWithNext':                       # :376-404
        pending = None
        for a, b in matches:
            if a > last_end:
                before = text[last_end:a]
                if pending is not None:
                    result.append(pending + before)
                else:
                    result.append(before)
            elif pending is not None:
                result.append(pending)
            pending = text[a:b]
            last_end = b
        if last_end < len(text):
            remaining = text[last_end:]
            result.append(pending + remaining if pending is not None else remaining)
        elif pending is not None:
            result.append(pending)
    else:                                                    # Contiguous :405-428
        cur = ''
        for a, b in matches:
            if a > last_end:
                if cur:
                    result.append(cur)
                    cur = ''
                result.append(text[last_end:a])
            cur += text[a:b]
            last_end = b
        if cur:
            result.append(cur)
        if last_end < len(text):
            result.append(text[last_end:])
    return [x for x in result if x]                          # :432
