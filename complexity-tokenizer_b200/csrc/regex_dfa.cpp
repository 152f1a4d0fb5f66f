// Host-side compiler for Split pre-tokenizer patterns: see regex_dfa.hpp.
#include "regex_dfa.hpp"

#include <algorithm>
#include <map>
#include <stdexcept>
#include <unordered_map>

#include "unicode_props_gen.h"

namespace ctk {
namespace {

constexpr uint32_t MAXCP = 0x10FFFF;
constexpr size_t MAX_INSTS = 20000, MAX_STATES = 8192, MAX_TABLE = 1u << 20;

struct Unsupported : std::runtime_error { using std::runtime_error::runtime_error; };

// ---- sets of code points: sorted, disjoint, non-adjacent ranges
typedef std::vector<std::pair<uint32_t, uint32_t>> RSet;

void normalise(RSet& s) {
    std::sort(s.begin(), s.end());
    RSet o;
    for (auto& r : s) {
        if (!o.empty() && r.first <= o.back().second + 1) o.back().second = std::max(o.back().second, r.second);
        else o.push_back(r);
    }
    s.swap(o);
}
RSet negate(RSet s) {
    normalise(s);
    RSet o;
    uint32_t next = 0;
    bool open = true;
    for (auto& r : s) {
        if (r.first > next) o.emplace_back(next, r.first - 1);
        if (r.second >= MAXCP) { open = false; break; }
        next = r.second + 1;
    }
    if (open) o.emplace_back(next, MAXCP);
    return o;
}
void unite(RSet& a, const RSet& b) { a.insert(a.end(), b.begin(), b.end()); }

const std::pair<uint32_t, uint32_t> kWhiteSpace[] = {{0x09, 0x0D}, {0x20, 0x20}, {0x85, 0x85}, {0xA0, 0xA0}, {0x1680, 0x1680}, {0x2000, 0x200A},
                                                     {0x2028, 0x2029}, {0x202F, 0x202F}, {0x205F, 0x205F}, {0x3000, 0x3000}};

RSet gc_set(const std::vector<std::string>& names) {          // General_Category values; "Cn" = everything unassigned
    RSet s;
    bool want_cn = false;
    std::vector<int> idx;
    for (auto& n : names) {
        if (n == "Cn") { want_cn = true; continue; }
        for (int k = 0; k < (int)(sizeof(UP_GC_NAMES) / sizeof(UP_GC_NAMES[0])); ++k) if (n == UP_GC_NAMES[k]) idx.push_back(k);
    }
    RSet assigned;
    for (int i = 0; i < UP_N_GC_RANGES; ++i) {
        if (want_cn) assigned.emplace_back(UP_GC_RANGES[i][0], UP_GC_RANGES[i][1]);
        for (int k : idx) if ((int)UP_GC_RANGES[i][2] == k) s.emplace_back(UP_GC_RANGES[i][0], UP_GC_RANGES[i][1]);
    }
    if (want_cn) unite(s, negate(assigned));
    normalise(s);
    return s;
}

std::vector<std::string> split_ws(const char* s) {
    std::vector<std::string> o;
    std::string cur;
    for (; *s; ++s) { if (*s == ' ') { if (!cur.empty()) o.push_back(cur); cur.clear(); } else cur += *s; }
    if (!cur.empty()) o.push_back(cur);
    return o;
}

RSet property_set(std::string name, bool neg) {                  // \p{name} / \P{name}
    if (!name.empty() && name[0] == '^') { neg = !neg; name.erase(0, 1); }
    for (const char* pre : {"gc=", "General_Category=", "sc=", "Script="})
        if (name.rfind(pre, 0) == 0) { name.erase(0, std::string(pre).size()); break; }
    if (name.rfind("scx=", 0) == 0) throw Unsupported("Script_Extensions");
    static const std::map<std::string, std::string> longs = {
        {"Letter", "L"}, {"Mark", "M"}, {"Number", "N"}, {"Punctuation", "P"}, {"Symbol", "S"}, {"Separator", "Z"}, {"Other", "C"},
        {"Uppercase_Letter", "Lu"}, {"Lowercase_Letter", "Ll"}, {"Titlecase_Letter", "Lt"}, {"Modifier_Letter", "Lm"}, {"Other_Letter", "Lo"},
        {"Decimal_Number", "Nd"}, {"Letter_Number", "Nl"}, {"Other_Number", "No"}, {"Nonspacing_Mark", "Mn"}, {"Spacing_Mark", "Mc"},
        {"Enclosing_Mark", "Me"}, {"Cased_Letter", "LC"}};
    static const std::map<std::string, const char*> groups = {
        {"L", "Lu Ll Lt Lm Lo"}, {"M", "Mn Mc Me"}, {"N", "Nd Nl No"}, {"P", "Pc Pd Ps Pe Pi Pf Po"}, {"S", "Sm Sc Sk So"},
        {"Z", "Zs Zl Zp"}, {"C", "Cc Cf Cs Co Cn"}, {"LC", "Lu Ll Lt"}};
    auto lg = longs.find(name);
    if (lg != longs.end()) name = lg->second;
    RSet s;
    auto g = groups.find(name);
    bool found = false;
    if (g != groups.end()) { s = gc_set(split_ws(g->second)); found = true; }
    if (!found) for (const char* n : UP_GC_NAMES) if (name == n) { s = gc_set({name}); found = true; break; }
    if (!found && (name == "White_Space" || name == "space" || name == "Whitespace")) { s.assign(std::begin(kWhiteSpace), std::end(kWhiteSpace)); found = true; }
    if (!found && name == "Alphabetic") { for (int i = 0; i < UP_N_ALPHA_RANGES; ++i) s.emplace_back(UP_ALPHA_RANGES[i][0], UP_ALPHA_RANGES[i][1]); found = true; }
    if (!found && name == "Any") { s.emplace_back(0, MAXCP); found = true; }
    if (!found) {
        const int ns = (int)(sizeof(UP_SCRIPT_NAMES) / sizeof(UP_SCRIPT_NAMES[0]));
        for (int k = 0; k < ns && !found; ++k)
            if (name == UP_SCRIPT_NAMES[k]) {
                for (int i = 0; i < UP_N_SCRIPT_RANGES; ++i) if ((int)UP_SCRIPT_RANGES[i][2] == k) s.emplace_back(UP_SCRIPT_RANGES[i][0], UP_SCRIPT_RANGES[i][1]);
                found = true;
            }
    }
    if (!found) throw Unsupported("unicode property '" + name + "'");
    normalise(s);
    return neg ? negate(s) : s;
}

RSet perl_set(uint32_t c) {                                      // \d \s \w and their negations
    RSet s;
    const uint32_t k = c | 0x20u;
    if (k == 's') s.assign(std::begin(kWhiteSpace), std::end(kWhiteSpace));
    else if (k == 'd') s = gc_set({"Nd"});
    else {                                                       // \w = Alphabetic | M | Nd | Pc | Join_Control
        s = gc_set({"Mn", "Mc", "Me", "Nd", "Pc"});
        for (int i = 0; i < UP_N_ALPHA_RANGES; ++i) s.emplace_back(UP_ALPHA_RANGES[i][0], UP_ALPHA_RANGES[i][1]);
        s.emplace_back(0x200C, 0x200D);
    }
    normalise(s);
    return (c & 0x20u) ? s : negate(s);
}

// ---- AST
enum { N_SET, N_CAT, N_ALT, N_REP };
struct Node { int kind = N_SET; int set = -1; std::vector<int> kids; int lo = 0, hi = 0; bool greedy = true; };   // hi < 0: unbounded

struct Parser {
    std::vector<uint32_t> p;
    size_t i = 0;
    std::vector<Node> nodes;
    std::vector<RSet> sets;

    uint32_t peek(size_t k = 0) const { return i + k < p.size() ? p[i + k] : 0xFFFFFFFFu; }
    bool eof() const { return i >= p.size(); }
    int new_node(const Node& n) { nodes.push_back(n); return (int)nodes.size() - 1; }
    int set_node(RSet s) { normalise(s); sets.push_back(std::move(s)); Node n; n.kind = N_SET; n.set = (int)sets.size() - 1; return new_node(n); }

    bool nullable(int id) const {
        const Node& n = nodes[id];
        switch (n.kind) {
            case N_SET: return false;
            case N_CAT: for (int k : n.kids) if (!nullable(k)) return false; return true;
            case N_ALT: for (int k : n.kids) if (nullable(k)) return true; return false;
            default: return n.lo == 0 || nullable(n.kids[0]);
        }
    }

    int alternation() {
        std::vector<int> br{concat()};
        while (peek() == '|') { ++i; br.push_back(concat()); }
        if (br.size() == 1) return br[0];
        Node n; n.kind = N_ALT; n.kids = br;
        return new_node(n);
    }
    int concat() {
        Node n; n.kind = N_CAT;
        while (!eof() && peek() != '|' && peek() != ')') n.kids.push_back(repeat());
        return new_node(n);
    }
    int repeat() {
        int atom_id = atom();
        for (;;) {
            const uint32_t c = peek();
            int lo, hi;
            if (c == '*') { lo = 0; hi = -1; ++i; }
            else if (c == '+') { lo = 1; hi = -1; ++i; }
            else if (c == '?') { lo = 0; hi = 1; ++i; }
            else if (c == '{') {
                size_t j = i + 1;
                auto number = [&](int& v) -> bool {
                    while (j < p.size() && p[j] == ' ') ++j;
                    size_t s = j;
                    long long x = 0;
                    while (j < p.size() && p[j] >= '0' && p[j] <= '9') { x = x * 10 + (p[j] - '0'); if (x > 1000) throw Unsupported("counted repetition too large"); ++j; }
                    if (j == s) return false;
                    while (j < p.size() && p[j] == ' ') ++j;
                    v = (int)x;
                    return true;
                };
                if (!number(lo)) throw Unsupported("counted repetition");
                if (j < p.size() && p[j] == '}') hi = lo;
                else if (j < p.size() && p[j] == ',') {
                    ++j;
                    while (j < p.size() && p[j] == ' ') ++j;
                    if (j < p.size() && p[j] == '}') hi = -1;
                    else if (!number(hi)) throw Unsupported("counted repetition");
                } else throw Unsupported("counted repetition");
                if (j >= p.size() || p[j] != '}') throw Unsupported("counted repetition");
                if (hi >= 0 && hi < lo) throw Unsupported("counted repetition bounds");
                i = j + 1;
            } else return atom_id;
            bool greedy = true;
            if (peek() == '?') { ++i; greedy = false; }
            if (nullable(atom_id) && (hi < 0 || hi > 1)) throw Unsupported("repetition of an operand that can match the empty string");
            Node n; n.kind = N_REP; n.kids = {atom_id}; n.lo = lo; n.hi = hi; n.greedy = greedy;
            atom_id = new_node(n);
        }
    }
    int atom() {
        const uint32_t c = peek();
        if (c == '(') {
            ++i;
            if (peek() == '?') {
                if (peek(1) == ':') i += 2;
                else if ((peek(1) == 'P' && peek(2) == '<') || (peek(1) == '<' && peek(2) != '=' && peek(2) != '!')) {
                    while (!eof() && peek() != '>') ++i;
                    if (eof()) throw Unsupported("group name");
                    ++i;
                } else throw Unsupported("flags or look-around");
            }
            int n = alternation();
            if (peek() != ')') throw Unsupported("unclosed group");
            ++i;
            return n;
        }
        if (c == '[') return set_node(bracket());
        if (c == '.') { ++i; return set_node(negate(RSet{{0x0A, 0x0A}})); }
        if (c == '\\') return set_node(escape());
        if (c == '*' || c == '+' || c == '?' || c == '{' || c == '^' || c == '$' || c == 0xFFFFFFFFu) throw Unsupported("operator without operand, or an anchor");
        ++i;
        return set_node(RSet{{c, c}});
    }
    static int hexval(uint32_t c) {
        if (c >= '0' && c <= '9') return (int)(c - '0');
        if ((c | 0x20u) >= 'a' && (c | 0x20u) <= 'f') return (int)((c | 0x20u) - 'a' + 10);
        return -1;
    }
    // after a backslash; `single` tells whether the result is one literal code point (usable as a range end)
    RSet escape(bool* single = nullptr) {
        ++i;
        const uint32_t c = peek();
        ++i;
        if (single) *single = false;
        if (c == 'd' || c == 's' || c == 'w' || c == 'D' || c == 'S' || c == 'W') return perl_set(c);
        if (c == 'p' || c == 'P') {
            std::string name;
            if (peek() == '{') {
                ++i;
                while (!eof() && peek() != '}') { if (peek() > 0x7E) throw Unsupported("unicode property name"); name += (char)peek(); ++i; }
                if (eof()) throw Unsupported("unclosed \\p{");
                ++i;
            } else {
                if (eof() || peek() > 0x7E) throw Unsupported("unicode property name");
                name += (char)peek();
                ++i;
            }
            return property_set(name, c == 'P');
        }
        if (single) *single = true;
        switch (c) {
            case 'n': return RSet{{10, 10}};
            case 'r': return RSet{{13, 13}};
            case 't': return RSet{{9, 9}};
            case 'f': return RSet{{12, 12}};
            case 'v': return RSet{{11, 11}};
            case 'a': return RSet{{7, 7}};
            default: break;
        }
        if (c == 'x' || c == 'u' || c == 'U') {
            uint64_t v = 0;
            int digits = 0;
            if (peek() == '{') {
                ++i;
                while (!eof() && peek() != '}') { int h = hexval(peek()); if (h < 0 || ++digits > 8) throw Unsupported("hex escape"); v = v * 16 + (uint64_t)h; ++i; }
                if (eof() || digits == 0) throw Unsupported("hex escape");
                ++i;
            } else {
                const int w = c == 'x' ? 2 : (c == 'u' ? 4 : 8);
                for (int k = 0; k < w; ++k) { int h = eof() ? -1 : hexval(peek()); if (h < 0) throw Unsupported("hex escape"); v = v * 16 + (uint64_t)h; ++i; }
            }
            if (v > MAXCP || (v >= 0xD800 && v <= 0xDFFF)) throw Unsupported("hex escape");
            return RSet{{(uint32_t)v, (uint32_t)v}};
        }
        const bool alnum = (c >= '0' && c <= '9') || ((c | 0x20u) >= 'a' && (c | 0x20u) <= 'z');
        if (c < 0x80 && !alnum && c != '<' && c != '>') return RSet{{c, c}};      // escaped punctuation is itself (\< \> are word boundaries)
        throw Unsupported("escape sequence");
    }
    RSet bracket() {
        ++i;
        RSet s;
        bool neg = false;
        if (peek() == '^') { neg = true; ++i; }
        bool first = true;
        for (;;) {
            const uint32_t c = peek();
            if (eof()) throw Unsupported("unclosed class");
            if (c == ']' && !first) { ++i; break; }
            first = false;
            if (c == '[') {
                if (peek(1) == ':') {
                    size_t j = i + 2;
                    std::string name;
                    while (j + 1 < p.size() && !(p[j] == ':' && p[j + 1] == ']')) { if (p[j] > 0x7E) throw Unsupported("POSIX class"); name += (char)p[j]; ++j; }
                    if (j + 1 >= p.size()) throw Unsupported("POSIX class");
                    bool pneg = false;
                    if (!name.empty() && name[0] == '^') { pneg = true; name.erase(0, 1); }
                    static const std::map<std::string, const char*> posix = {
                        {"alnum", "0-9A-Za-z"}, {"alpha", "A-Za-z"}, {"ascii", "\x01-\x7f"}, {"blank", "\t "}, {"cntrl", "\x01-\x1f\x7f"}, {"digit", "0-9"},
                        {"graph", "!-~"}, {"lower", "a-z"}, {"print", " -~"}, {"punct", "!-/:-@[-`{-~"}, {"space", "\t\n\x0b\x0c\r "}, {"upper", "A-Z"},
                        {"word", "0-9A-Za-z_"}, {"xdigit", "0-9A-Fa-f"}};
                    auto f = posix.find(name);
                    if (f == posix.end()) throw Unsupported("POSIX class");
                    RSet ps;
                    const std::string spec = f->second;
                    for (size_t k = 0; k < spec.size();) {
                        if (k + 2 < spec.size() && spec[k + 1] == '-') { ps.emplace_back((uint8_t)spec[k], (uint8_t)spec[k + 2]); k += 3; }
                        else { ps.emplace_back((uint8_t)spec[k], (uint8_t)spec[k]); k += 1; }
                    }
                    if (name == "ascii" || name == "cntrl") ps.emplace_back(0, 0);     // (NUL cannot sit inside the C string above)
                    normalise(ps);
                    unite(s, pneg ? negate(ps) : ps);
                    i = j + 2;
                    continue;
                }
                unite(s, bracket());
                continue;
            }
            if ((c == '&' || c == '-' || c == '~') && peek(1) == c) throw Unsupported("class set operation");
            uint32_t lo;
            if (c == '\\') {
                bool single;
                RSet item = escape(&single);
                if (!single) { unite(s, item); continue; }
                lo = item[0].first;
            } else { lo = c; ++i; }
            if (peek() == '-' && peek(1) != ']' && !(i + 1 >= p.size())) {
                if (peek(1) == '-') throw Unsupported("class set operation");
                ++i;
                uint32_t hi;
                if (peek() == '\\') {
                    bool single;
                    RSet item = escape(&single);
                    if (!single) throw Unsupported("class range end");
                    hi = item[0].first;
                } else if (peek() == '[') throw Unsupported("class range end");
                else { hi = peek(); ++i; }
                if (hi < lo) throw Unsupported("class range out of order");
                s.emplace_back(lo, hi);
            } else s.emplace_back(lo, lo);
        }
        normalise(s);
        return neg ? negate(s) : s;
    }
};

// ---- NFA (Thompson), built back to front: compile(node, next) returns the entry of `node` continuing at `next`
enum { I_CHAR, I_SPLIT, I_MATCH };
struct Inst { int op, set, x, y; };                              // CHAR: set, x = next; SPLIT: x preferred over y

struct Nfa {
    const Parser& ps;
    std::vector<Inst> prog;
    explicit Nfa(const Parser& p) : ps(p) {}
    int emit(Inst in) {
        if (prog.size() >= MAX_INSTS) throw Unsupported("pattern too large");
        prog.push_back(in);
        return (int)prog.size() - 1;
    }
    int compile(int id, int next) {
        const Node& n = ps.nodes[id];
        switch (n.kind) {
            case N_SET: return emit({I_CHAR, n.set, next, -1});
            case N_CAT: { int cur = next; for (size_t k = n.kids.size(); k-- > 0;) cur = compile(n.kids[k], cur); return cur; }
            case N_ALT: {
                int cur = compile(n.kids.back(), next);
                for (size_t k = n.kids.size() - 1; k-- > 0;) { int s = compile(n.kids[k], next); cur = emit({I_SPLIT, -1, s, cur}); }
                return cur;
            }
            default: {
                const int sub = n.kids[0];
                int cur;
                if (n.hi < 0) {                                  // x{lo,}: lo copies, then a loop
                    const int loop = emit({I_SPLIT, -1, -1, -1});
                    const int body = compile(sub, loop);
                    prog[loop].x = n.greedy ? body : next;
                    prog[loop].y = n.greedy ? next : body;
                    cur = loop;
                } else {                                         // x{lo,hi}: hi - lo nested optional copies
                    cur = next;
                    for (int k = n.lo; k < n.hi; ++k) {
                        const int body = compile(sub, cur);
                        cur = n.greedy ? emit({I_SPLIT, -1, body, next}) : emit({I_SPLIT, -1, next, body});
                    }
                }
                for (int k = 0; k < n.lo; ++k) cur = compile(sub, cur);
                return cur;
            }
        }
    }
};

void add_thread(const std::vector<Inst>& prog, int pc, std::vector<int>& list, std::vector<uint8_t>& seen) {
    std::vector<int> stack{pc};                                  // depth-first, preferred branch first = priority order
    while (!stack.empty()) {
        const int q = stack.back();
        stack.pop_back();
        if (seen[q]) continue;
        seen[q] = 1;
        if (prog[q].op == I_SPLIT) { stack.push_back(prog[q].y); stack.push_back(prog[q].x); }
        else list.push_back(q);
    }
}
void cut_after_match(const std::vector<Inst>& prog, std::vector<int>& list) {
    for (size_t k = 0; k < list.size(); ++k) if (prog[list[k]].op == I_MATCH) { list.resize(k + 1); return; }
}

}  // namespace

int compile_split_regex(const std::string& pattern, SplitDfa& out, std::string& err) {
    try {
        Parser ps;
        for (size_t k = 0; k < pattern.size();) {                // the pattern is a JSON string: valid UTF-8
            const unsigned char c = (unsigned char)pattern[k];
            uint32_t cp;
            int len = c < 0x80 ? 1 : c < 0xE0 ? 2 : c < 0xF0 ? 3 : 4;
            if (k + len > pattern.size()) throw Unsupported("pattern is not UTF-8");
            if (len == 1) cp = c;
            else if (len == 2) cp = ((c & 0x1Fu) << 6) | ((unsigned char)pattern[k + 1] & 63u);
            else if (len == 3) cp = ((c & 0x0Fu) << 12) | (((unsigned char)pattern[k + 1] & 63u) << 6) | ((unsigned char)pattern[k + 2] & 63u);
            else cp = ((c & 7u) << 18) | (((unsigned char)pattern[k + 1] & 63u) << 12) | (((unsigned char)pattern[k + 2] & 63u) << 6) | ((unsigned char)pattern[k + 3] & 63u);
            ps.p.push_back(cp);
            k += len;
        }
        const int root = ps.alternation();
        if (!ps.eof()) throw Unsupported("unbalanced )");
        if (ps.nullable(root)) throw Unsupported("pattern can match the empty string");
        Nfa nfa(ps);
        const int match_pc = nfa.emit({I_MATCH, -1, -1, -1});
        const int start_pc = nfa.compile(root, match_pc);
        const std::vector<Inst>& prog = nfa.prog;

        // ---- equivalence classes of code points: intervals between all range ends, merged by membership signature
        std::vector<uint32_t> cuts{0, MAXCP + 1};
        for (auto& s : ps.sets) for (auto& r : s) { cuts.push_back(r.first); cuts.push_back(r.second + 1); }
        std::sort(cuts.begin(), cuts.end());
        cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
        const size_t n_iv = cuts.size() - 1, n_sets = ps.sets.size();
        std::vector<std::vector<uint8_t>> sig(n_iv, std::vector<uint8_t>(n_sets, 0));
        for (size_t s = 0; s < n_sets; ++s)
            for (auto& r : ps.sets[s]) {
                size_t a = std::lower_bound(cuts.begin(), cuts.end(), r.first) - cuts.begin();
                size_t b = std::lower_bound(cuts.begin(), cuts.end(), r.second + 1) - cuts.begin();
                for (size_t k = a; k < b; ++k) sig[k][s] = 1;
            }
        std::map<std::vector<uint8_t>, int> class_of_sig;
        std::vector<int> iv_class(n_iv);
        std::vector<std::vector<uint8_t>> class_sig;
        for (size_t k = 0; k < n_iv; ++k) {
            auto f = class_of_sig.find(sig[k]);
            if (f == class_of_sig.end()) { f = class_of_sig.emplace(sig[k], (int)class_sig.size()).first; class_sig.push_back(sig[k]); }
            iv_class[k] = f->second;
        }
        const size_t n_cls = class_sig.size();
        if (n_cls > 256) throw Unsupported("pattern distinguishes too many classes of characters");

        // ---- ordered-subset construction
        std::map<std::vector<int>, int> id_of;
        std::vector<std::vector<int>> states;
        auto intern = [&](std::vector<int>& l) -> int {
            cut_after_match(prog, l);
            auto f = id_of.find(l);
            if (f != id_of.end()) return f->second;
            if (states.size() >= MAX_STATES) throw Unsupported("pattern needs too many automaton states");
            id_of.emplace(l, (int)states.size());
            states.push_back(l);
            return (int)states.size() - 1;
        };
        std::vector<int> dead;
        intern(dead);                                            // state 0
        std::vector<uint8_t> seen(prog.size());
        std::vector<int> l0;
        add_thread(prog, start_pc, l0, seen);
        const int start = intern(l0);
        std::vector<uint16_t> trans;
        for (size_t s = 0; s < states.size(); ++s) {
            if ((s + 1) * n_cls > MAX_TABLE) throw Unsupported("pattern needs too large a transition table");
            trans.resize((s + 1) * n_cls, 0);
            for (size_t c = 0; c < n_cls; ++c) {
                std::vector<int> nl;
                std::fill(seen.begin(), seen.end(), 0);
                const std::vector<int> cur = states[s];          // (copy: `states` may grow)
                for (int pc : cur) if (prog[pc].op == I_CHAR && class_sig[c][prog[pc].set]) add_thread(prog, prog[pc].x, nl, seen);
                const int t = intern(nl);
                bool acc = false;
                for (int pc : states[t]) if (prog[pc].op == I_MATCH) acc = true;
                trans[s * n_cls + c] = (uint16_t)(t | (acc ? 0x8000 : 0));
            }
        }
        out.n_states = (uint32_t)states.size();
        out.n_classes = (uint32_t)n_cls;
        out.start = (uint32_t)start;
        out.trans.swap(trans);
        // ---- code point -> class: ASCII directly, the rest through a two-stage table with shared blocks
        std::vector<uint8_t> flat(MAXCP + 1);
        for (size_t k = 0; k < n_iv; ++k) std::fill(flat.begin() + cuts[k], flat.begin() + cuts[k + 1], (uint8_t)iv_class[k]);
        out.ascii_class.assign(flat.begin(), flat.begin() + 128);
        out.neutral[0] = out.neutral[1] = out.neutral[2] = out.neutral[3] = 0;
        for (uint32_t c = 0; c < 128; ++c) {                       // neutral: in no set at all (the all-zero membership signature)
            bool in_any = false;
            for (uint8_t v : class_sig[out.ascii_class[c]]) in_any = in_any || v;
            if (!in_any) out.neutral[c >> 5] |= 1u << (c & 31);
        }
        {   // class pairs that can be adjacent inside a match: some live state steps on c1 to a live state that steps on c2 to a live state
            std::vector<uint8_t> possible(n_cls * n_cls, 0);
            for (uint32_t st = 1; st < out.n_states; ++st)
                for (size_t c1 = 0; c1 < n_cls; ++c1) {
                    const uint32_t t = out.trans[st * n_cls + c1] & 0x7FFFu;
                    if (!t) continue;
                    for (size_t c2 = 0; c2 < n_cls; ++c2) if (out.trans[t * n_cls + c2] & 0x7FFFu) possible[c1 * n_cls + c2] = 1;
                }
            out.pair_impossible.assign(512, 0);
            for (uint32_t b1 = 0; b1 < 128; ++b1)
                for (uint32_t b2 = 0; b2 < 128; ++b2)
                    if (!possible[out.ascii_class[b1] * n_cls + out.ascii_class[b2]]) out.pair_impossible[b1 * 4 + (b2 >> 5)] |= 1u << (b2 & 31);
        }
        out.trans_ascii.clear();
        if (out.n_states <= 512) {
            out.trans_ascii.resize((size_t)out.n_states * 128);
            for (uint32_t st = 0; st < out.n_states; ++st)
                for (uint32_t c = 0; c < 128; ++c) out.trans_ascii[st * 128 + c] = out.trans[st * n_cls + out.ascii_class[c]];
        }
        out.stage1.assign(0x1100, 0);
        out.blocks.clear();
        std::map<std::vector<uint8_t>, uint16_t> block_id;
        for (uint32_t b = 0; b < 0x1100; ++b) {
            std::vector<uint8_t> blk(flat.begin() + b * 256, flat.begin() + b * 256 + 256);
            auto f = block_id.find(blk);
            if (f == block_id.end()) {
                f = block_id.emplace(blk, (uint16_t)(out.blocks.size() / 256)).first;
                out.blocks.insert(out.blocks.end(), blk.begin(), blk.end());
            }
            out.stage1[b] = f->second;
        }
        return 0;
    } catch (const Unsupported& e) {
        err = std::string("Split pattern outside the supported subset: ") + e.what();
        return 1;
    }
}

}  // namespace ctk
