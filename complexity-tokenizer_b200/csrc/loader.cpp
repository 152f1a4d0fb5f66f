// tokenizer.json -> HostModel, following the reference's load rules:
//   schema / merges forms        src/huggingface/mod.rs:31-116
//   from_tokenizer_json_*        src/huggingface/mod.rs:247-334
//   BpeTokenizer::new            src/bpe.rs:52-79   (rank = original index, merges vector compacted;
//                                                     new_id read back as merges[rank] at bpe.rs:141)
//   Vocab::new                   src/vocab.rs:47-51
//   component defaults           src/huggingface/parsing.rs:10-90, 93-190, 272-364
#include "model.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "../../include/ctk.h"
#include "json.hpp"

namespace ctk {

void byte_map(uint32_t byte_to_cp[256]) {
    int n = 0;
    for (int b = 0; b < 256; ++b) {
        bool keep = (b >= '!' && b <= '~') || (b >= 0xA1 && b <= 0xAC) || (b >= 0xAE);
        byte_to_cp[b] = keep ? (uint32_t)b : (uint32_t)(256 + n++);
    }
}

namespace {

void put_utf8(std::string& s, uint32_t cp) {
    if (cp < 0x80) s += (char)cp;
    else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 63)); }
    else if (cp < 0x10000) { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 63)); s += (char)(0x80 | (cp & 63)); }
    else { s += (char)(0xF0 | (cp >> 18)); s += (char)(0x80 | ((cp >> 12) & 63)); s += (char)(0x80 | ((cp >> 6) & 63)); s += (char)(0x80 | (cp & 63)); }
}

// decode one code point of a valid UTF-8 std::string
uint32_t next_cp(const std::string& s, size_t& i) {
    unsigned char c = (unsigned char)s[i];
    if (c < 0x80) { ++i; return c; }
    if (c < 0xE0) { uint32_t r = ((c & 0x1Fu) << 6) | ((unsigned char)s[i + 1] & 63u); i += 2; return r; }
    if (c < 0xF0) { uint32_t r = ((c & 0x0Fu) << 12) | (((unsigned char)s[i + 1] & 63u) << 6) | ((unsigned char)s[i + 2] & 63u); i += 3; return r; }
    uint32_t r = ((c & 7u) << 18) | (((unsigned char)s[i + 1] & 63u) << 12) | (((unsigned char)s[i + 2] & 63u) << 6) | ((unsigned char)s[i + 3] & 63u);
    i += 4;
    return r;
}

// Does Rust `regex` certainly REJECT this pattern?  It has no look-around and no back-references: Regex::new then fails
// and regex_split_with_behavior returns the text unchanged (pretokenizers.rs:299-302).  The scan knows escapes and
// character classes: an escaped `\(\?=` or a `(?=` inside `[...]` is no look-around, and `\1` inside a class is no
// back-reference the decision may rest on.  Anything not certainly rejected is treated as compiling, i.e. as a Split
// this library would have to apply -- and reports UNSUPPORTED rather than guessing (never a silent pass-through).
bool rust_regex_certainly_rejected(const std::string& p) {
    int cls = 0;                                                // depth of [...] nesting (Rust allows nested classes)
    for (size_t i = 0; i < p.size(); ++i) {
        const char c = p[i];
        if (c == '\\') {
            if (i + 1 < p.size() && !cls && p[i + 1] >= '1' && p[i + 1] <= '9') return true;      // back-reference
            ++i;                                                // whatever is escaped is not syntax
            continue;
        }
        if (cls) {
            if (c == '[') ++cls;
            else if (c == ']') --cls;
            continue;
        }
        if (c == '[') {
            cls = 1;
            if (i + 1 < p.size() && p[i + 1] == '^') ++i;
            if (i + 1 < p.size() && p[i + 1] == ']') ++i;       // a leading ] is a literal
            continue;
        }
        if (c == '(' && i + 2 < p.size() && p[i + 1] == '?') {
            if (p[i + 2] == '=' || p[i + 2] == '!') return true;
            if (p[i + 2] == '<' && i + 3 < p.size() && (p[i + 3] == '=' || p[i + 3] == '!')) return true;
        }
    }
    return false;
}

const char* type_of(const JValue* v) {
    if (!v || !v->is_obj()) return nullptr;
    const JValue* t = v->get("type");
    if (!t) return nullptr;
    return t->is_str() ? t->s.c_str() : "";
}

// parsing.rs:10-90.  ok=false => unsupported
bool parse_normalizer(const JValue* v, bool& nfc, std::string& err) {
    const char* t = type_of(v);
    if (!t) { nfc = true; return true; }                       // missing/null/malformed => NFC (:89)
    std::string ty = t;
    if (ty == "NFC") { nfc = true; return true; }
    if (ty == "Sequence") {
        const JValue* seq = v->get("normalizers");
        nfc = false;
        if (!seq || !seq->is_arr()) return true;                // => None
        for (auto& x : seq->arr) {
            bool sub = false;
            JValue tmp = x;
            if (!parse_normalizer(x.t == JValue::Null ? nullptr : &x, sub, err)) return false;
            nfc = nfc || sub;                                   // NFC is idempotent
        }
        return true;
    }
    static const char* out_of_scope[] = {"NFD", "NFKC", "NFKD", "Lowercase", "Strip", "StripAccents", "Replace",
                                         "Prepend", "BertNormalizer", "Precompiled"};
    for (const char* o : out_of_scope)
        if (ty == o) { err = "normalizer '" + ty + "' is outside the ByteLevel-BPE hot path"; return false; }
    nfc = false;                                                // unknown type => None
    return true;
}

// parsing.rs:93-190.  Collects ByteLevel stages; returns 0 ok, 1 = "None", 2 = unsupported
// One parsed pre-tokenizer stage: the ByteLevel stage (with its add_prefix_space) or a compiled Split stage.
struct PreStage { bool bytelevel = false; bool metaspace = false; bool add_prefix_space = false; uint32_t replacement = 0x2581; SplitStage split; };

// parsing.rs:93-190.  0 = parsed (stages appended), 1 = None (unknown type), 2 = unsupported (err set)
int parse_pre(const JValue* v, std::vector<PreStage>& stages, std::string& err) {
    const char* t = type_of(v);
    if (!t) { PreStage b; b.bytelevel = true; stages.push_back(b); return 0; }   // default ByteLevel{false} (:187-189)
    std::string ty = t;
    if (ty == "ByteLevel") {
        const JValue* a = v->get("add_prefix_space");
        PreStage b;
        b.bytelevel = true;
        b.add_prefix_space = a && a->t == JValue::Bool ? a->b : false;    // use_regex / trim_offsets ignored (:99-107)
        stages.push_back(b);
        return 0;
    }
    if (ty == "Metaspace") {                                    // parsing.rs:108-123
        PreStage ms;
        ms.metaspace = true;
        const JValue* r = v->get("replacement");
        if (r && r->is_str() && !r->s.empty()) { size_t i = 0; ms.replacement = next_cp(r->s, i); }
        const JValue* a = v->get("add_prefix_space");
        ms.add_prefix_space = a && a->t == JValue::Bool ? a->b : true;
        stages.push_back(ms);
        return 0;
    }
    if (ty == "Split") {                                        // parsing.rs:145-167 -> SplitWithBehavior
        const JValue* pat = v->get("pattern");
        const JValue* rx = pat ? pat->get("Regex") : nullptr;
        PreStage sp;
        sp.split.pattern = (rx && rx->is_str()) ? rx->s : "";   // a {"String": ...} pattern reads as "" (:147-151)
        if (rust_regex_certainly_rejected(sp.split.pattern)) return 0;    // Regex::new fails: text passes through (pretokenizers.rs:299-302)
        const JValue* inv = v->get("invert");
        sp.split.invert = inv && inv->t == JValue::Bool ? inv->b : false;
        const JValue* bh = v->get("behavior");
        const std::string b = bh && bh->is_str() ? bh->s : "Removed";
        sp.split.behavior = b == "Isolated" ? SPLIT_ISOLATED : b == "MergedWithPrevious" ? SPLIT_MERGED_PREV : b == "MergedWithNext" ? SPLIT_MERGED_NEXT
                            : b == "Contiguous" ? SPLIT_CONTIGUOUS : SPLIT_REMOVED;
        if (compile_split_regex(sp.split.pattern, sp.split.dfa, err) != 0) return 2;
        stages.push_back(sp);
        return 0;
    }
    if (ty == "Sequence") {
        const JValue* seq = v->get("pretokenizers");
        if (!seq || !seq->is_arr()) return 1;
        bool any = false;
        for (auto& x : seq->arr) {
            int r = parse_pre(x.t == JValue::Null ? nullptr : &x, stages, err);
            if (r == 2) return 2;
            if (r == 0) any = true;
        }
        return any ? 0 : 1;
    }
    static const char* out_of_scope[] = {"Whitespace", "WhitespaceSplit", "Punctuation", "BertPreTokenizer",
                                         "CharDelimiterSplit", "UnicodeScripts", "Digits"};
    for (const char* o : out_of_scope)
        if (ty == o) { err = "pre_tokenizer '" + ty + "' is outside the ByteLevel-BPE hot path"; return 2; }
    return 1;                                                   // unknown type => None
}

// parsing.rs:253-270
std::string template_from_array(const JValue& arr) {
    std::string out;
    bool first = true;
    for (auto& item : arr.arr) {
        if (!item.is_obj()) continue;
        std::string part;
        bool have = false;
        if (const JValue* sp = item.get("SpecialToken")) {
            const JValue* id = sp->get("id");
            if (id && id->is_str()) { part = id->s; have = true; }
        } else if (const JValue* sq = item.get("Sequence")) {
            const JValue* id = sq->get("id");
            if (id && id->is_str()) { part = "$" + id->s; have = true; }
        }
        if (!have) continue;
        if (!first) out += ' ';
        out += part;
        first = false;
    }
    return out;
}

bool rust_is_whitespace(uint32_t c) {
    return (c >= 9 && c <= 13) || c == 0x20 || c == 0x85 || c == 0xA0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200A) ||
           c == 0x2028 || c == 0x2029 || c == 0x202F || c == 0x205F || c == 0x3000;
}

// The template walk of postprocessors.rs:88-148 with pair_ids = None, recorded instead of executed.
void template_items(const std::string& tpl, const std::unordered_map<std::string, uint32_t>& specials, std::vector<int64_t>& items) {
    std::vector<uint32_t> ch;
    for (size_t i = 0; i < tpl.size();) ch.push_back(next_cp(tpl, i));
    size_t i = 0;
    while (i < ch.size()) {
        if (ch[i] == '$' && i + 1 < ch.size()) {
            if (ch[i + 1] == 'A') { items.push_back(-1); i += 2; }
            else if (ch[i + 1] == 'B') { i += 2; }
            else i += 1;
        } else if (ch[i] == '<' || ch[i] == '[') {
            const uint32_t endc = ch[i] == '<' ? '>' : ']';
            size_t start = i;
            while (i < ch.size() && ch[i] != endc) ++i;
            if (i < ch.size()) ++i;
            size_t a = start, b = i;                            // str::trim
            while (a < b && rust_is_whitespace(ch[a])) ++a;
            while (b > a && rust_is_whitespace(ch[b - 1])) --b;
            std::string tok;
            for (size_t k = a; k < b; ++k) put_utf8(tok, ch[k]);
            auto it = specials.find(tok);
            if (it != specials.end()) items.push_back((int64_t)it->second);
        } else {
            i += 1;
        }
    }
}

// parsing.rs:193-250
void parse_post_processor(const JValue* v, const std::unordered_map<std::string, uint32_t>& specials, HostModel& m) {
    m.pp_items.clear();
    m.has_post_processor = false;
    auto get = [&](const char* k, uint32_t dflt) { auto it = specials.find(k); return it == specials.end() ? dflt : it->second; };
    const char* t = type_of(v);
    std::string ty = t ? t : "";
    if (ty == "TemplateProcessing") {
        const JValue* s = v->get("single");
        std::string single = (s && s->is_arr()) ? template_from_array(*s) : "<s> $A </s>";
        template_items(single, specials, m.pp_items);
        m.has_post_processor = true;
    } else if (ty == "RobertaProcessing") {
        m.pp_items = {(int64_t)get("<s>", 0), -1, (int64_t)get("</s>", 2)};
        m.has_post_processor = true;
    } else if (ty == "BertProcessing") {
        m.pp_items = {(int64_t)get("[CLS]", 101), -1, (int64_t)get("[SEP]", 102)};
        m.has_post_processor = true;
    } else {
        m.pp_items = {-1};
    }
}

int klass_ascii(uint8_t c) {   // 0 other, 1 letter, 2 number, 3 whitespace
    if ((c | 0x20) >= 'a' && (c | 0x20) <= 'z') return 1;
    if (c >= '0' && c <= '9') return 2;
    if (c == ' ' || (c >= 9 && c <= 13)) return 3;
    return 0;
}

// Can this added token ever occur inside ONE byte-mapped pre-token (mod.rs:566-610 searches words,
// not raw text)?  Conservative: "false" only when provably impossible.  A pre-token is a
// contraction, a whitespace run, or [one leading 0x20] + a run of a single class.
void analyse_added(AddedTok& a, const std::unordered_map<uint32_t, uint8_t>& cp_to_byte) {
    a.may_match = false;
    a.bytes.clear();
    for (size_t i = 0; i < a.content.size();) {
        uint32_t cp = next_cp(a.content, i);
        auto it = cp_to_byte.find(cp);
        if (it == cp_to_byte.end()) return;                     // char outside the mapped alphabet
        a.bytes.push_back(it->second);
    }
    if (a.bytes.empty()) return;
    std::string s((const char*)a.bytes.data(), a.bytes.size());
    static const char* contr[] = {"'s", "'t", "'re", "'ve", "'m", "'ll", "'d"};
    for (const char* c : contr) if (std::string(c).find(s) != std::string::npos) { a.may_match = true; return; }
    int seen = -1;
    for (size_t i = 0; i < a.bytes.size(); ++i) {
        uint8_t b = a.bytes[i];
        if (b >= 0x80) continue;                                // part of a multi-byte char: class unknown here
        int k = klass_ascii(b);
        if (i == 0 && b == 0x20) continue;                      // glued leading space, or first char of a \s+ run
        if (seen < 0) seen = k;
        else if (seen != k) return;                             // two classes in one pre-token: impossible
    }
    if (a.bytes[0] == 0x20 && a.bytes.size() > 1 && seen == 3) { a.may_match = true; return; }
    a.may_match = true;
}

// Round-parallel merging (encode_long.cuh / encode_xlong.cuh) looks, for a pair (x, y), only as far as a token
// around x or y can reach.  Needs: a monotone table, products that are the concatenation of their parts (the
// rank quirk of bpe.rs:52-79 inactive), and one token per id.
void token_reach(HostModel& m) {
    m.round_parallel = false;
    m.reach.clear();
    if (!m.merges_monotone || m.pairs.empty()) return;
    if (getenv("CTK_NO_REACH")) return;                           // debug/tests: keep the uniform-window rounds
    if (m.vocab.size() != (size_t)std::count(m.id_present.begin(), m.id_present.end(), (uint8_t)1)) return;   // two tokens, one id
    for (const PairEntry& p : m.pairs) {
        if (p.a >= m.id_to_token.size() || p.b >= m.id_to_token.size() || p.new_id >= m.id_to_token.size()) return;
        if (m.id_to_token[p.a] + m.id_to_token[p.b] != m.id_to_token[p.new_id]) return;
    }
    std::vector<uint32_t> wl(m.id_to_token.size(), 0), wr(m.id_to_token.size(), 0);
    std::vector<size_t> cut;                                    // byte offsets of the code points of one token
    for (size_t id = 0; id < m.id_to_token.size(); ++id) {
        if (!m.id_present[id]) continue;
        const std::string& t = m.id_to_token[id];
        cut.clear();
        for (size_t i = 0; i < t.size();) { cut.push_back(i); next_cp(t, i); }
        const size_t L = cut.size();
        if (L > 0xFFFF) return;
        for (size_t c = 1; c < L; ++c) {
            auto ip = m.vocab.find(t.substr(0, cut[c]));
            if (ip != m.vocab.end()) wr[ip->second] = std::max<uint32_t>(wr[ip->second], (uint32_t)(L - c));
            auto is = m.vocab.find(t.substr(cut[c]));
            if (is != m.vocab.end()) wl[is->second] = std::max<uint32_t>(wl[is->second], (uint32_t)c);
        }
    }
    m.reach.resize(wl.size());
    for (size_t i = 0; i < wl.size(); ++i) m.reach[i] = wl[i] | (wr[i] << 16);
    m.round_parallel = true;
}

}  // namespace

int load_model(const uint8_t* json, size_t len, HostModel& m, std::string& err) {
    JValue root;
    {
        JParser p((const char*)json, len);
        if (!p.parse(root, err)) return CTK_ERR_INVALID_DATA;
    }
    if (!root.is_obj()) { err = "invalid type: expected struct TokenizerJson"; return CTK_ERR_INVALID_DATA; }
    const JValue* ver = root.get("version");
    if (ver && ver->t != JValue::Null && !ver->is_str()) { err = "version: expected a string"; return CTK_ERR_INVALID_DATA; }
    const JValue* model = root.get("model");
    if (!model || !model->is_obj()) { err = "missing field `model`"; return CTK_ERR_INVALID_DATA; }
    const JValue* mt = model->get("type");
    if (mt && mt->t != JValue::Null && !mt->is_str()) { err = "model.type: expected a string"; return CTK_ERR_INVALID_DATA; }
    const JValue* vocab = model->get("vocab");
    if (!vocab || !vocab->is_obj()) { err = "missing field `vocab`"; return CTK_ERR_INVALID_DATA; }
    uint32_t max_id = 0;
    for (auto& kv : vocab->obj) {
        const JValue& v = kv.second;
        if (v.t != JValue::Num || !v.is_uint || v.u > 0xFFFFFFFFull) { err = "vocab id is not a u32"; return CTK_ERR_INVALID_DATA; }
        m.vocab[kv.first] = (uint32_t)v.u;                      // duplicate keys: last wins
    }
    for (auto& kv : m.vocab) max_id = std::max(max_id, kv.second);
    if (!m.vocab.empty() && max_id > (1u << 24)) { err = "vocab ids above 2^24 are not supported"; return CTK_ERR_UNSUPPORTED; }

    // ---- merges (mod.rs:56-101 then :252-264)
    std::vector<std::pair<std::string, std::string>> merges;
    const JValue* mj = model->get("merges");
    if (mj) {
        if (!mj->is_arr()) { err = "merges: expected a sequence"; return CTK_ERR_INVALID_DATA; }
        for (auto& it : mj->arr) {
            std::string line;
            if (it.is_str()) line = it.s;
            else if (it.is_arr()) {
                if (it.arr.size() == 2 && it.arr[0].is_str() && it.arr[1].is_str()) line = it.arr[0].s + " " + it.arr[1].s;
                else continue;
            } else continue;
            size_t sp = line.find(' ');
            if (sp == std::string::npos) continue;
            if (line.find(' ', sp + 1) != std::string::npos) continue;      // split(' ') must give exactly 2 parts
            merges.emplace_back(line.substr(0, sp), line.substr(sp + 1));
        }
    }
    // ---- BpeTokenizer::new (bpe.rs:52-79)
    std::unordered_map<uint64_t, uint32_t> rank_of;
    std::vector<uint32_t> ops_new_id;
    for (size_t rank = 0; rank < merges.size(); ++rank) {
        auto ia = m.vocab.find(merges[rank].first);
        auto ib = m.vocab.find(merges[rank].second);
        if (ia == m.vocab.end() || ib == m.vocab.end()) continue;
        auto im = m.vocab.find(merges[rank].first + merges[rank].second);
        if (im == m.vocab.end()) continue;
        rank_of[((uint64_t)ia->second << 32) | ib->second] = (uint32_t)rank;    // insert: later overwrites
        ops_new_id.push_back(im->second);
    }
    m.pairs.reserve(rank_of.size());
    for (auto& kv : rank_of) {
        if (kv.second >= ops_new_id.size()) {
            err = "merges table would make the reference panic (bpe.rs:141: rank beyond the compacted merges vector)";
            return CTK_ERR_UNSUPPORTED;
        }
        m.pairs.push_back({(uint32_t)(kv.first >> 32), (uint32_t)kv.first, kv.second, ops_new_id[kv.second]});
    }
    std::sort(m.pairs.begin(), m.pairs.end(), [](const PairEntry& x, const PairEntry& y) { return x.rank < y.rank; });
    {   // monotone?  (a trainer's table is: a token exists before any pair uses it; duplicates / the rank quirk can break it)
        std::unordered_map<uint32_t, uint32_t> made_at;        // id -> highest rank of a merge producing it
        for (const PairEntry& p : m.pairs) { auto it = made_at.find(p.new_id); if (it == made_at.end() || it->second < p.rank) made_at[p.new_id] = p.rank; }
        m.merges_monotone = true;
        for (const PairEntry& p : m.pairs) {
            auto ia = made_at.find(p.a), ib = made_at.find(p.b);
            if ((ia != made_at.end() && ia->second >= p.rank) || (ib != made_at.end() && ib->second >= p.rank)) { m.merges_monotone = false; break; }
        }
        m.max_token_span = 1;
        if (m.merges_monotone) {                               // spans in rank order: components are final before they are used
            std::unordered_map<uint32_t, uint32_t> span;
            for (const PairEntry& p : m.pairs) {
                auto ia = span.find(p.a), ib = span.find(p.b);
                uint64_t s = (uint64_t)(ia == span.end() ? 1u : ia->second) + (ib == span.end() ? 1u : ib->second);
                if (s > 0x7FFFFFFFull) s = 0x7FFFFFFFull;
                uint32_t& dst = span[p.new_id];
                if (dst < s) dst = (uint32_t)s;
                if (m.max_token_span < s) m.max_token_span = (uint32_t)s;
            }
        }
    }

    // ---- added tokens (mod.rs:274-305)
    const JValue* aj = root.get("added_tokens");
    std::unordered_map<std::string, size_t> added_idx;
    std::unordered_map<std::string, uint32_t> special_map;
    if (aj) {
        if (!aj->is_arr()) { err = "added_tokens: expected a sequence"; return CTK_ERR_INVALID_DATA; }
        for (auto& t : aj->arr) {
            const JValue *id = t.get("id"), *content = t.get("content"), *special = t.get("special");
            if (!t.is_obj() || !id || id->t != JValue::Num || !id->is_uint || id->u > 0xFFFFFFFFull || !content ||
                !content->is_str() || !special || special->t != JValue::Bool) {
                err = "added_tokens: missing or mistyped id/content/special";
                return CTK_ERR_INVALID_DATA;
            }
            AddedTok a;
            a.content = content->s;
            a.id = (uint32_t)id->u;
            a.special = special->b;
            auto flag = [&](const char* k, bool& dst) -> bool {
                const JValue* f = t.get(k);
                if (!f) { dst = false; return true; }
                if (f->t != JValue::Bool) return false;
                dst = f->b;
                return true;
            };
            bool normalized;
            if (!flag("single_word", a.single_word) || !flag("lstrip", a.lstrip) || !flag("rstrip", a.rstrip) ||
                !flag("normalized", normalized)) { err = "added_tokens: flag is not a bool"; return CTK_ERR_INVALID_DATA; }
            auto f = added_idx.find(a.content);
            if (f == added_idx.end()) { added_idx[a.content] = m.added.size(); m.added.push_back(a); }
            else m.added[f->second] = a;
            if (a.special) special_map[a.content] = a.id;
        }
    }
    for (auto& kv : special_map) m.specials.emplace_back(kv.first, kv.second);
    std::sort(m.specials.begin(), m.specials.end());

    // ---- components
    const JValue* nj = root.get("normalizer");
    if (!parse_normalizer(nj && nj->t != JValue::Null ? nj : nullptr, m.nfc, err)) return CTK_ERR_UNSUPPORTED;
    std::vector<PreStage> stages;
    const JValue* pj = root.get("pre_tokenizer");
    int pr = parse_pre(pj && pj->t != JValue::Null ? pj : nullptr, stages, err);
    if (pr == 2) return CTK_ERR_UNSUPPORTED;
    // the hot path: any number of Split stages, then exactly one ByteLevel stage, nothing after it (a Split behind the
    // ByteLevel stage would see byte-mapped text, a second ByteLevel stage would map twice)
    if (pr == 1 || stages.empty() || !(stages.back().bytelevel || stages.back().metaspace)) {
        err = "pre_tokenizer must resolve to Split stages followed by exactly one ByteLevel (or Metaspace) stage for the hot path";
        return CTK_ERR_UNSUPPORTED;
    }
    m.split_stages.clear();
    for (size_t k = 0; k + 1 < stages.size(); ++k) {
        if (stages[k].bytelevel || stages[k].metaspace) { err = "pre_tokenizer must resolve to Split stages followed by exactly one ByteLevel stage for the hot path"; return CTK_ERR_UNSUPPORTED; }
        m.split_stages.push_back(std::move(stages[k].split));
    }
    m.metaspace = stages.back().metaspace;
    m.add_prefix_space = m.metaspace ? false : stages.back().add_prefix_space;
    m.meta_replacement = stages.back().replacement;
    m.meta_prefix = stages.back().add_prefix_space;
    if (m.metaspace && (rust_is_whitespace(m.meta_replacement) || m.meta_replacement == 0)) {
        err = "Metaspace replacement is a white-space character (words would contain the spaces they are split on)";
        return CTK_ERR_UNSUPPORTED;
    }
    const JValue* dj = root.get("decoder");
    const char* dt = type_of(dj && dj->t != JValue::Null ? dj : nullptr);
    m.dec_metaspace = false;
    if (dt && std::string(dt) == "Metaspace") {                  // parsing.rs:279-293
        m.dec_metaspace = true;
        const JValue* r = dj->get("replacement");
        m.dec_meta_replacement = 0x2581;
        if (r && r->is_str() && !r->s.empty()) { size_t i = 0; m.dec_meta_replacement = next_cp(r->s, i); }
        const JValue* a = dj->get("add_prefix_space");
        m.dec_meta_strip = a && a->t == JValue::Bool ? a->b : true;
    } else if (dt && std::string(dt) != "ByteLevel") { err = std::string("decoder '") + dt + "' is outside the hot path (ByteLevel and Metaspace are built)"; return CTK_ERR_UNSUPPORTED; }

    // ---- derived tables
    uint32_t b2c[256];
    byte_map(b2c);
    std::unordered_map<uint32_t, uint8_t> c2b;
    for (int b = 0; b < 256; ++b) {
        c2b[b2c[b]] = (uint8_t)b;
        std::string s;
        put_utf8(s, b2c[b]);
        auto it = m.vocab.find(s);
        m.byte_init_id[b] = it == m.vocab.end() ? kNoId : it->second;
    }
    m.any_added_may_match = false;
    for (auto& a : m.added) {
        if (a.content.empty()) { err = "empty added token (the reference loops forever on it)"; return CTK_ERR_UNSUPPORTED; }
        if (m.metaspace) {                                       // words are raw text without white space: any token without white space may occur
            a.bytes.assign(a.content.begin(), a.content.end());
            a.may_match = true;
            for (size_t i = 0; i < a.content.size();) if (rust_is_whitespace(next_cp(a.content, i))) a.may_match = false;
            if (a.may_match && a.single_word) {
                err = "added token with single_word inside a Metaspace pipeline (needs char::is_alphanumeric of its neighbours)";
                return CTK_ERR_UNSUPPORTED;
            }
        } else analyse_added(a, c2b);
        m.any_added_may_match = m.any_added_may_match || a.may_match;
    }
    m.char_ids.clear();
    for (auto& kv : m.vocab) {                                   // single-character entries: the initial symbols of bpe.rs:94-97
        if (kv.first.empty()) continue;
        size_t i = 0;
        const uint32_t cp = next_cp(kv.first, i);
        if (i == kv.first.size()) m.char_ids.emplace_back(cp, kv.second);
    }
    std::sort(m.char_ids.begin(), m.char_ids.end());
    m.meta_empty_ids.clear();
    if (m.metaspace && m.meta_prefix) {                          // the word "<replacement>": an added token equal to it (mod.rs:566-594), else its vocabulary id, else nothing
        std::string w;
        put_utf8(w, m.meta_replacement);
        const AddedTok* best = nullptr;
        for (auto& a : m.added) if (a.may_match && a.content == w) best = &a;
        if (best) m.meta_empty_ids.push_back(best->id);
        else { auto it = m.vocab.find(w); if (it != m.vocab.end()) m.meta_empty_ids.push_back(it->second); }
    }
    // Vocab::new (vocab.rs:48-51): id -> token.  Two tokens with one id: the reference keeps an
    // arbitrary one (hash iteration order); we keep the lexicographically largest, deterministically.
    size_t nid = m.vocab.empty() ? 0 : (size_t)max_id + 1;
    m.id_to_token.assign(nid, std::string());
    m.id_present.assign(nid, 0);
    for (auto& kv : m.vocab) {
        if (!m.id_present[kv.second] || m.id_to_token[kv.second] < kv.first) m.id_to_token[kv.second] = kv.first;
        m.id_present[kv.second] = 1;
    }
    m.dec_off.assign(nid + 1, 0);
    m.dec_special.assign(nid, 0);
    m.dec_blob.clear();
    m.dec_max_bytes = 0;
    for (size_t id = 0; id < nid; ++id) {
        m.dec_off[id] = (uint32_t)m.dec_blob.size();
        if (!m.id_present[id]) continue;
        const std::string& tok = m.id_to_token[id];
        size_t before = m.dec_blob.size();
        for (size_t i = 0; i < tok.size() && m.dec_metaspace;) {   // decoders.rs:121-124: join, then replacement -> ' '
            const size_t i0 = i;
            const uint32_t cp = next_cp(tok, i);
            if (cp == m.dec_meta_replacement) m.dec_blob.push_back(0x20);
            else m.dec_blob.insert(m.dec_blob.end(), tok.begin() + i0, tok.begin() + i);
        }
        for (size_t i = 0; i < tok.size() && !m.dec_metaspace;) {  // decoders.rs:100-116
            uint32_t cp = next_cp(tok, i);
            if (cp == 0x120) { m.dec_blob.push_back(0x20); continue; }
            auto it = c2b.find(cp);
            if (it != c2b.end()) m.dec_blob.push_back(it->second);
            else if (cp < 0x80) m.dec_blob.push_back((uint8_t)cp);
        }
        m.dec_max_bytes = std::max(m.dec_max_bytes, m.dec_blob.size() - before);
        if (special_map.count(tok)) m.dec_special[id] = 1;
    }
    m.dec_off[nid] = (uint32_t)m.dec_blob.size();
    // ---- rich Encoding outputs
    {
        std::unordered_map<std::string, uint32_t> smap(special_map.begin(), special_map.end());
        const JValue* ppj = root.get("post_processor");
        parse_post_processor(ppj && ppj->t != JValue::Null ? ppj : nullptr, smap, m);
        m.token_str_len.assign(nid, 0);
        for (size_t id = 0; id < nid; ++id) if (m.id_present[id]) m.token_str_len[id] = (uint32_t)m.id_to_token[id].size();
        auto p1 = smap.find("[PAD]"), p2 = smap.find("<pad>");
        m.pad_id = p1 != smap.end() ? p1->second : p2 != smap.end() ? p2->second : 0u;
        m.pad_token = (m.pad_id < nid && m.id_present[m.pad_id]) ? m.id_to_token[m.pad_id] : std::string("<pad>");
    }
    token_reach(m);
    return CTK_OK;
}

}  // namespace ctk
