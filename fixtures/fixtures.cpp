// Synthetic fixtures for tests and bench (NOT product code, NOT oracle).
//
// The reference ships no corpora, no tokenizer.json and there is no network,
// so every fixture is synthetic and generated here, deterministically from a seed:
//   fx_gen_corpus      config-1/2 English-like text and config-3 mixed French/CJK/emoji text
//                      
//   fx_train_bpe       incremental byte-level BPE trainer that produces the merges table for
//                      the synthetic tokenizer.json files.  It plays the role of the reference's
//                      bpe_trainer (src/bpe_trainer.rs:100-228) -- same objective (most frequent
//                      adjacent pair, min_frequency), deterministic tie-break (count, then pair)
//                      where the reference's is hash-iteration order (bpe_trainer.rs:152-155).
// Build: g++ -O2 -shared -fPIC -pthread fixtures/fixtures.cpp -o fixtures/_build/libfixtures.so
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <queue>
#include <set>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

struct Rng {
    uint64_t s[4];
    static uint64_t splitmix(uint64_t& x) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    explicit Rng(uint64_t seed) { for (auto& v : s) v = splitmix(seed); }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
    bool chance(double p) { return uni() < p; }
    double normal() {
        double u1 = uni(), u2 = uni();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};

// English letter frequencies (per mille, a..z)
const int LETTER_W[26] = {82, 15, 28, 43, 127, 22, 20, 61, 70, 2, 8, 40, 24,
                          67, 75, 19, 1, 60, 63, 91, 28, 10, 24, 2, 20, 1};

struct Lexicon {
    std::vector<std::string> words;
    std::vector<double> cdf;
    void build(uint64_t seed, int n, double zipf_s, const std::vector<std::string>& alphabet,
               const std::vector<int>& weights, int minlen, int maxlen) {
        Rng r(seed);
        int wsum = 0;
        for (int w : weights) wsum += w;
        std::unordered_set<std::string> seen;
        words.reserve(n);
        for (int rank = 0; rank < n; ++rank) {
            for (int attempt = 0;; ++attempt) {
                double mu = 1.5 + 0.75 * std::log2((double)rank + 2.0) * 0.62 + attempt * 0.5;
                int len = (int)std::lround(mu + r.normal() * 1.6);
                len = std::max(minlen, std::min(maxlen, len));
                std::string w;
                for (int k = 0; k < len; ++k) {
                    int x = (int)r.below(wsum), j = 0;
                    while (x >= weights[j]) x -= weights[j++];
                    w += alphabet[j];
                }
                if (seen.insert(w).second) { words.push_back(w); break; }
            }
        }
        cdf.resize(n);
        double acc = 0;
        for (int i = 0; i < n; ++i) { acc += 1.0 / std::pow((double)i + 1.0, zipf_s); cdf[i] = acc; }
        for (auto& c : cdf) c /= acc;
    }
    const std::string& sample(Rng& r) const {
        double u = r.uni();
        size_t i = std::lower_bound(cdf.begin(), cdf.end(), u) - cdf.begin();
        if (i >= words.size()) i = words.size() - 1;
        return words[i];
    }
};

void put_utf8(std::string& s, uint32_t cp) {
    if (cp < 0x80) s += (char)cp;
    else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 63)); }
    else if (cp < 0x10000) { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 63)); s += (char)(0x80 | (cp & 63)); }
    else { s += (char)(0xF0 | (cp >> 18)); s += (char)(0x80 | ((cp >> 12) & 63)); s += (char)(0x80 | ((cp >> 6) & 63)); s += (char)(0x80 | (cp & 63)); }
}
std::string utf8(uint32_t cp) { std::string s; put_utf8(s, cp); return s; }

struct Gen {
    Lexicon en, fr;
    std::vector<uint32_t> han;
    bool ascii_only;
};

const char* CONTR[] = {"'s", "n't", "'re", "'ve", "'ll", "'d", "'m"};

// Western (English / French) prose.  `decomp_p` = probability that an accented letter is written
// decomposed (base + combining mark) to exercise NFC.
void gen_western(Rng& r, const Lexicon& lx, std::string& out, size_t target, double decomp_p, bool tabs) {
    int sent_in_par = 0, par_len = 3 + r.below(6);
    if (tabs && r.chance(0.3)) out += '\t';
    while (out.size() < target) {
        int nw = 5 + r.below(21);
        bool open_q = false, open_p = false;
        for (int w = 0; w < nw; ++w) {
            if (w) out += r.chance(0.01) ? "  " : " ";
            if (!open_q && r.chance(0.02)) { out += '"'; open_q = true; }
            if (!open_p && r.chance(0.01)) { out += '('; open_p = true; }
            if (r.chance(0.03)) {                       // digit group
                int nd = 1 + r.below(4);
                for (int k = 0; k < nd; ++k) out += (char)('0' + r.below(10));
                if (r.chance(0.15)) { out += r.chance(0.5) ? '.' : ','; for (int k = 0; k < 2 + (int)r.below(2); ++k) out += (char)('0' + r.below(10)); }
            } else {
                std::string wd = lx.sample(r);
                if (w == 0 || r.chance(0.03)) {
                    if (wd[0] >= 'a' && wd[0] <= 'z') wd[0] = (char)(wd[0] - 32);
                }
                if (decomp_p > 0) {                     // rewrite some precomposed accents as base+mark
                    std::string t;
                    for (size_t i = 0; i < wd.size();) {
                        unsigned char c = (unsigned char)wd[i];
                        if (c == 0xC3 && i + 1 < wd.size() && r.chance(decomp_p)) {
                            unsigned char d = (unsigned char)wd[i + 1];
                            uint32_t cp = 0xC0 + (d - 0x80);
                            char base = 0; uint32_t mark = 0;
                            switch (cp) {
                                case 0xE9: base = 'e'; mark = 0x301; break; case 0xE8: base = 'e'; mark = 0x300; break;
                                case 0xEA: base = 'e'; mark = 0x302; break; case 0xEB: base = 'e'; mark = 0x308; break;
                                case 0xE0: base = 'a'; mark = 0x300; break; case 0xE2: base = 'a'; mark = 0x302; break;
                                case 0xE7: base = 'c'; mark = 0x327; break; case 0xF9: base = 'u'; mark = 0x300; break;
                                case 0xFB: base = 'u'; mark = 0x302; break; case 0xF4: base = 'o'; mark = 0x302; break;
                                case 0xEE: base = 'i'; mark = 0x302; break; case 0xEF: base = 'i'; mark = 0x308; break;
                                default: break;
                            }
                            if (base) { t += base; put_utf8(t, mark); i += 2; continue; }
                        }
                        t += wd[i++];
                    }
                    wd.swap(t);
                }
                out += wd;
                if (r.chance(0.02)) out += CONTR[r.below(7)];
            }
            if (open_p && r.chance(0.3)) { out += ')'; open_p = false; }
            if (open_q && r.chance(0.25)) { out += '"'; open_q = false; }
            if (w + 1 < nw) {
                double u = r.uni();
                if (u < 0.08) out += ',';
                else if (u < 0.09) out += ';';
                else if (u < 0.10) out += ':';
                else if (u < 0.105) out += " -";
            }
        }
        if (open_p) out += ')';
        if (open_q) out += '"';
        double u = r.uni();
        out += u < 0.85 ? '.' : (u < 0.93 ? '?' : '!');
        if (++sent_in_par >= par_len) { out += "\n\n"; sent_in_par = 0; par_len = 3 + r.below(6); if (tabs && r.chance(0.2)) out += '\t'; }
        else out += ' ';
    }
}

void gen_cjk(Rng& r, const Gen& g, std::string& out, size_t target) {
    static const uint32_t punct[] = {0x3002, 0x3001, 0xFF01, 0xFF1F, 0x300C, 0x300D, 0xFF0C};
    while (out.size() < target) {
        int n = 4 + r.below(40);
        for (int k = 0; k < n; ++k) {
            double u = r.uni();
            if (u < 0.70) { double z = r.uni(); put_utf8(out, g.han[(size_t)(z * z * g.han.size())]); }
            else if (u < 0.88) put_utf8(out, 0x3041 + r.below(0x3094 - 0x3041));   // Hiragana
            else put_utf8(out, 0x30A1 + r.below(0x30F7 - 0x30A1));                   // Katakana
        }
        double u = r.uni();
        if (u < 0.75) put_utf8(out, punct[r.below(7)]);
        else if (u < 0.85) put_utf8(out, 0x3000);           // ideographic space: \s
        else if (u < 0.92) out += ' ';
        else if (u < 0.96) { int nd = 1 + r.below(4); for (int k = 0; k < nd; ++k) put_utf8(out, r.chance(0.5) ? ('0' + r.below(10)) : (0xFF10 + r.below(10))); }
        else out += "\n";
    }
}

void gen_emoji(Rng& r, const Gen& g, std::string& out, size_t target) {
    while (out.size() < target) {
        double u = r.uni();
        if (u < 0.45) { out += g.en.sample(r); out += ' '; }
        else if (u < 0.80) { int n = 1 + r.below(3); for (int k = 0; k < n; ++k) put_utf8(out, 0x1F600 + r.below(0x50)); if (r.chance(0.5)) out += ' '; }
        else if (u < 0.90) { put_utf8(out, 0x1F44D); put_utf8(out, 0x1F3FB + r.below(5)); out += ' '; }
        else if (u < 0.96) { put_utf8(out, 0x1F468); put_utf8(out, 0x200D); put_utf8(out, 0x1F469); put_utf8(out, 0x200D); put_utf8(out, 0x1F467); }
        else out += r.chance(0.5) ? "!!! " : "\n";
    }
}

// cut `s` to at most `len` bytes on a UTF-8 boundary, then pad with 'x' / ' ' to exactly len
void fit(std::string& s, size_t len) {
    if (s.size() > len) {
        size_t k = len;
        while (k > 0 && ((unsigned char)s[k] & 0xC0) == 0x80) --k;
        s.resize(k);
        // never end a doc on a bare combining mark cut: harmless either way (still valid UTF-8)
    }
    while (s.size() < len) s += (s.size() + 1 < len) ? 'x' : '.';
}

}  // namespace

extern "C" {

// kind: 0 = English (config 1), 1 = ASCII English with tabs (config 2), 2 = mixed French/CJK/emoji (config 3)
// Fills `out` with exactly sum(doc lengths) bytes and offsets[0..n_docs].  Returns n_docs (or -1 if
// capacity is insufficient).  Total size is the first prefix of doc lengths reaching target_bytes.
long fx_gen_corpus(int kind, uint64_t seed, uint64_t target_bytes, double doc_median, uint32_t doc_min,
                   uint32_t doc_max, int lexicon_size, uint8_t* out, uint64_t cap, uint64_t* offsets,
                   uint64_t max_docs, int n_threads) {
    Gen g;
    g.ascii_only = kind != 2;
    std::vector<std::string> az;
    std::vector<int> w(LETTER_W, LETTER_W + 26);
    for (int i = 0; i < 26; ++i) az.push_back(std::string(1, (char)('a' + i)));
    g.en.build(0xC0FFEE ^ 0x1111, lexicon_size, 1.07, az, w, 1, 14);     // lexicon independent of corpus seed
    if (kind == 2) {
        std::vector<std::string> fa = az;
        std::vector<int> fw = w;
        const uint32_t acc[] = {0xE9, 0xE8, 0xEA, 0xE0, 0xE7, 0xF9, 0xF4, 0xEE, 0xEB, 0xE2, 0xFB, 0xEF, 0x153};
        const int aw[] = {25, 8, 4, 7, 4, 2, 2, 2, 1, 2, 1, 1, 1};
        for (int i = 0; i < 13; ++i) { fa.push_back(utf8(acc[i])); fw.push_back(aw[i]); }
        g.fr.build(0xC0FFEE ^ 0x2222, lexicon_size, 1.07, fa, fw, 1, 14);
        for (int i = 0; i < 3500; ++i) g.han.push_back(0x4E00 + (uint32_t)i * 5 + (i % 3));
    }
    // phase 1: document lengths (sequential)
    Rng r(seed);
    std::vector<uint64_t> off{0};
    double sigma = 0.8, mu = std::log(doc_median);
    while (off.back() < target_bytes) {
        double L = std::exp(mu + sigma * r.normal());
        uint64_t len = (uint64_t)std::max<double>(doc_min, std::min<double>(doc_max, L));
        if (off.back() + len > target_bytes) len = target_bytes - off.back();
        off.push_back(off.back() + len);
    }
    uint64_t nd = off.size() - 1;
    if (nd > max_docs || off.back() > cap) return -1;
    for (uint64_t i = 0; i <= nd; ++i) offsets[i] = off[i];
    // phase 2: fill each doc (parallel; per-doc seed => independent of thread count)
    if (n_threads < 1) n_threads = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
        th.emplace_back([&, t]() {
            std::string s;
            for (uint64_t d = t; d < nd; d += n_threads) {
                size_t len = (size_t)(off[d + 1] - off[d]);
                Rng dr(seed * 0x9E3779B97F4A7C15ull + d * 0xD1B54A32D192ED03ull + 7);
                s.clear();
                if (kind == 0) gen_western(dr, g.en, s, len, 0, false);
                else if (kind == 1) gen_western(dr, g.en, s, len, 0, true);
                else {
                    double u = dr.uni();
                    if (dr.chance(0.001)) {
                        s += "<s> </s><pad> ";
                        put_utf8(s, 0x212B); s += ' '; put_utf8(s, 0xF900); s += ' ';
                    }
                    if (u < 0.4) gen_western(dr, g.fr, s, len, 0.01, false);
                    else if (u < 0.8) gen_cjk(dr, g, s, len);
                    else if (u < 0.9) gen_emoji(dr, g, s, len);
                    else gen_western(dr, g.en, s, len, 0, false);
                }
                fit(s, len);
                std::memcpy(out + off[d], s.data(), len);
            }
        });
    for (auto& x : th) x.join();
    return (long)nd;
}

// ---------------------------------------------------------------------------------------------
// BPE trainer over a byte-level word histogram.
// text/offsets: training sample.  Words are split with a simplified byte-class rule (letters incl.
// all bytes >= 0x80, digits, whitespace, other; one leading space glued) -- it only has to produce
// a plausible table, not to match the reference's regex.
// Output: merges as (left_symbol, right_symbol) in symbol space: 0..255 = bytes, 256+k = k-th merge.
// Returns number of merges produced (<= n_merges).
long fx_train_bpe(const uint8_t* text, uint64_t n, long n_merges, long min_frequency, uint32_t* out_pairs) {
    auto klass = [](uint8_t c) -> int {
        if ((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c >= 0x80) return 1;
        if (c >= '0' && c <= '9') return 2;
        if (c == ' ' || (c >= 9 && c <= 13)) return 3;
        return 0;
    };
    std::unordered_map<std::string, long> hist;
    for (uint64_t i = 0; i < n;) {
        uint64_t s = i;
        int k = klass(text[i]);
        if (text[i] == ' ' && i + 1 < n && klass(text[i + 1]) != 3) { ++i; k = klass(text[i]); }
        while (i < n && klass(text[i]) == k) ++i;
        if (k == 1 && i - s > 24) {                  // long CJK runs: cut into pieces so training stays local
            for (uint64_t p = s; p < i; p += 12) hist[std::string((const char*)text + p, std::min<uint64_t>(12, i - p))]++;
        } else {
            hist[std::string((const char*)text + s, i - s)]++;
        }
    }
    struct Word { std::vector<uint32_t> sym; long cnt; };
    std::vector<Word> words;
    {
        std::vector<std::pair<std::string, long>> hv(hist.begin(), hist.end());
        std::sort(hv.begin(), hv.end());
        for (auto& kv : hv) {
            Word w; w.cnt = kv.second;
            for (unsigned char c : kv.first) w.sym.push_back(c);
            words.push_back(std::move(w));
        }
    }
    typedef uint64_t Pair;
    auto mk = [](uint32_t a, uint32_t b) { return ((uint64_t)a << 32) | b; };
    std::unordered_map<Pair, long> cnt;
    std::unordered_map<Pair, std::vector<uint32_t>> where;
    for (uint32_t wi = 0; wi < words.size(); ++wi) {
        auto& s = words[wi].sym;
        for (size_t i = 0; i + 1 < s.size(); ++i) {
            Pair p = mk(s[i], s[i + 1]);
            cnt[p] += words[wi].cnt;
            auto& v = where[p];
            if (v.empty() || v.back() != wi) v.push_back(wi);
        }
    }
    // max-heap on (count, smaller pair first)
    typedef std::pair<long, Pair> HE;
    auto cmp = [](const HE& a, const HE& b) { return a.first != b.first ? a.first < b.first : a.second > b.second; };
    std::priority_queue<HE, std::vector<HE>, decltype(cmp)> heap(cmp);
    for (auto& kv : cnt) heap.push({kv.second, kv.first});
    long made = 0;
    while (made < n_merges && !heap.empty()) {
        HE top = heap.top(); heap.pop();
        auto it = cnt.find(top.second);
        if (it == cnt.end() || it->second != top.first) continue;     // stale
        if (top.first < min_frequency) break;
        uint32_t a = (uint32_t)(top.second >> 32), b = (uint32_t)top.second, nid = 256 + (uint32_t)made;
        out_pairs[2 * made] = a; out_pairs[2 * made + 1] = b; ++made;
        std::vector<uint32_t> ws; ws.swap(where[top.second]);
        cnt.erase(it);
        std::set<Pair> touched;
        for (uint32_t wi : ws) {
            auto& s = words[wi].sym; long c = words[wi].cnt;
            std::vector<uint32_t> ns; ns.reserve(s.size());
            bool any = false;
            for (size_t i = 0; i < s.size();) {
                if (i + 1 < s.size() && s[i] == a && s[i + 1] == b) { ns.push_back(nid); i += 2; any = true; }
                else ns.push_back(s[i++]);
            }
            if (!any) continue;
            for (size_t i = 0; i + 1 < s.size(); ++i) { Pair p = mk(s[i], s[i + 1]); if (p != top.second) { cnt[p] -= c; touched.insert(p); } }
            for (size_t i = 0; i + 1 < ns.size(); ++i) {
                Pair p = mk(ns[i], ns[i + 1]); cnt[p] += c; touched.insert(p);
                auto& v = where[p];
                if (v.empty() || v.back() != wi) v.push_back(wi);
            }
            s.swap(ns);
        }
        for (Pair p : touched) {
            auto f = cnt.find(p);
            if (f == cnt.end()) continue;
            if (f->second <= 0) cnt.erase(f); else heap.push({f->second, p});
        }
    }
    return made;
}

}  // extern "C"
