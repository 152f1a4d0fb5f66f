// Minimal JSON reader for tokenizer.json (no third-party dependency).
// Follows serde_json's behaviour where the reference depends on it (src/huggingface/mod.rs:162):
// objects keep every key in file order and get() returns the LAST duplicate (a HashMap insert
// overwrites), strings are UTF-8 with \uXXXX escapes (surrogate pairs combined, lone surrogates
// rejected), numbers remember whether they were written as integers.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

namespace ctk {

struct JValue {
    enum Type { Null, Bool, Num, Str, Arr, Obj } t = Null;
    bool b = false;
    bool is_uint = false;      // written as a non-negative integer that fits uint64
    uint64_t u = 0;
    double num = 0;
    std::string s;
    std::vector<JValue> arr;
    std::vector<std::pair<std::string, JValue>> obj;

    const JValue* get(const char* key) const {
        if (t != Obj) return nullptr;
        const JValue* r = nullptr;
        for (auto& kv : obj)
            if (kv.first == key) r = &kv.second;
        return r;
    }
    bool is_str() const { return t == Str; }
    bool is_obj() const { return t == Obj; }
    bool is_arr() const { return t == Arr; }
};

class JParser {
public:
    JParser(const char* p, size_t n) : p_(p), e_(p + n) {}
    bool parse(JValue& out, std::string& err) {
        // serde_json tolerates a UTF-8 BOM? No: it errors.  We do the same (no skipping).
        ws();
        if (!value(out, 0)) { err = err_; return false; }
        ws();
        if (p_ != e_) { err = "trailing characters"; return false; }
        return true;
    }

private:
    const char* p_;
    const char* e_;
    std::string err_;
    bool fail(const char* m) { if (err_.empty()) err_ = m; return false; }
    void ws() { while (p_ < e_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) ++p_; }
    static void put_utf8(std::string& s, uint32_t cp) {
        if (cp < 0x80) s += (char)cp;
        else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 63)); }
        else if (cp < 0x10000) { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 63)); s += (char)(0x80 | (cp & 63)); }
        else { s += (char)(0xF0 | (cp >> 18)); s += (char)(0x80 | ((cp >> 12) & 63)); s += (char)(0x80 | ((cp >> 6) & 63)); s += (char)(0x80 | (cp & 63)); }
    }
    bool hex4(uint32_t& v) {
        if (e_ - p_ < 4) return fail("bad \\u escape");
        v = 0;
        for (int i = 0; i < 4; ++i) {
            char c = *p_++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= (uint32_t)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (uint32_t)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (uint32_t)(c - 'A' + 10);
            else return fail("bad \\u escape");
        }
        return true;
    }
    bool string(std::string& out) {
        // *p_ == '"'
        ++p_;
        out.clear();
        for (;;) {
            if (p_ >= e_) return fail("EOF in string");
            const char* q = p_;
            while (q < e_ && *q != '"' && *q != '\\' && (unsigned char)*q >= 0x20) ++q;
            out.append(p_, q - p_);
            p_ = q;
            if (p_ >= e_) return fail("EOF in string");
            char c = *p_++;
            if (c == '"') break;
            if ((unsigned char)c < 0x20) return fail("control character in string");
            // backslash
            if (p_ >= e_) return fail("EOF in escape");
            char x = *p_++;
            switch (x) {
                case '"': out += '"'; break;
                case '\\': out += '\\'; break;
                case '/': out += '/'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'n': out += '\n'; break;
                case 'r': out += '\r'; break;
                case 't': out += '\t'; break;
                case 'u': {
                    uint32_t v;
                    if (!hex4(v)) return false;
                    if (v >= 0xDC00 && v <= 0xDFFF) return fail("lone trailing surrogate");
                    if (v >= 0xD800 && v <= 0xDBFF) {
                        if (e_ - p_ < 2 || p_[0] != '\\' || p_[1] != 'u') return fail("lone leading surrogate");
                        p_ += 2;
                        uint32_t lo;
                        if (!hex4(lo)) return false;
                        if (lo < 0xDC00 || lo > 0xDFFF) return fail("lone leading surrogate");
                        v = 0x10000 + ((v - 0xD800) << 10) + (lo - 0xDC00);
                    }
                    put_utf8(out, v);
                    break;
                }
                default: return fail("bad escape");
            }
        }
        return valid_utf8(out) ? true : fail("invalid UTF-8 in string");
    }
    static bool valid_utf8(const std::string& s) {
        const unsigned char* b = (const unsigned char*)s.data();
        size_t n = s.size(), i = 0;
        while (i < n) {
            unsigned char c = b[i];
            if (c < 0x80) { ++i; continue; }
            size_t need; unsigned char lo = 0x80, hi = 0xBF;
            if (c >= 0xC2 && c <= 0xDF) need = 1;
            else if (c == 0xE0) { need = 2; lo = 0xA0; }
            else if (c >= 0xE1 && c <= 0xEC) need = 2;
            else if (c == 0xED) { need = 2; hi = 0x9F; }
            else if (c >= 0xEE && c <= 0xEF) need = 2;
            else if (c == 0xF0) { need = 3; lo = 0x90; }
            else if (c >= 0xF1 && c <= 0xF3) need = 3;
            else if (c == 0xF4) { need = 3; hi = 0x8F; }
            else return false;
            if (i + need >= n) return false;
            if (b[i + 1] < lo || b[i + 1] > hi) return false;
            for (size_t k = 2; k <= need; ++k) if ((b[i + k] & 0xC0) != 0x80) return false;
            i += need + 1;
        }
        return true;
    }
    bool number(JValue& out) {
        const char* s = p_;
        bool neg = false, integral = true;
        if (p_ < e_ && *p_ == '-') { neg = true; ++p_; }
        if (p_ >= e_ || *p_ < '0' || *p_ > '9') return fail("bad number");
        if (*p_ == '0') ++p_;
        else while (p_ < e_ && *p_ >= '0' && *p_ <= '9') ++p_;
        if (p_ < e_ && *p_ == '.') {
            integral = false; ++p_;
            if (p_ >= e_ || *p_ < '0' || *p_ > '9') return fail("bad number");
            while (p_ < e_ && *p_ >= '0' && *p_ <= '9') ++p_;
        }
        if (p_ < e_ && (*p_ == 'e' || *p_ == 'E')) {
            integral = false; ++p_;
            if (p_ < e_ && (*p_ == '+' || *p_ == '-')) ++p_;
            if (p_ >= e_ || *p_ < '0' || *p_ > '9') return fail("bad number");
            while (p_ < e_ && *p_ >= '0' && *p_ <= '9') ++p_;
        }
        std::string tmp(s, p_ - s);
        out.t = JValue::Num;
        out.num = strtod(tmp.c_str(), nullptr);
        out.is_uint = false;
        if (integral && !neg && tmp.size() <= 19) { out.is_uint = true; out.u = strtoull(tmp.c_str(), nullptr, 10); }
        return true;
    }
    bool value(JValue& out, int depth) {
        if (depth > 128) return fail("recursion limit exceeded");
        ws();
        if (p_ >= e_) return fail("EOF while parsing a value");
        char c = *p_;
        if (c == '{') {
            ++p_;
            out.t = JValue::Obj;
            ws();
            if (p_ < e_ && *p_ == '}') { ++p_; return true; }
            for (;;) {
                ws();
                if (p_ >= e_ || *p_ != '"') return fail("key must be a string");
                std::string k;
                if (!string(k)) return false;
                ws();
                if (p_ >= e_ || *p_ != ':') return fail("expected ':'");
                ++p_;
                out.obj.emplace_back(std::move(k), JValue());
                if (!value(out.obj.back().second, depth + 1)) return false;
                ws();
                if (p_ < e_ && *p_ == ',') { ++p_; continue; }
                if (p_ < e_ && *p_ == '}') { ++p_; return true; }
                return fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            ++p_;
            out.t = JValue::Arr;
            ws();
            if (p_ < e_ && *p_ == ']') { ++p_; return true; }
            for (;;) {
                out.arr.emplace_back();
                if (!value(out.arr.back(), depth + 1)) return false;
                ws();
                if (p_ < e_ && *p_ == ',') { ++p_; continue; }
                if (p_ < e_ && *p_ == ']') { ++p_; return true; }
                return fail("expected ',' or ']'");
            }
        }
        if (c == '"') { out.t = JValue::Str; return string(out.s); }
        if (c == 't') { if (e_ - p_ >= 4 && !memcmp(p_, "true", 4)) { p_ += 4; out.t = JValue::Bool; out.b = true; return true; } return fail("bad literal"); }
        if (c == 'f') { if (e_ - p_ >= 5 && !memcmp(p_, "false", 5)) { p_ += 5; out.t = JValue::Bool; out.b = false; return true; } return fail("bad literal"); }
        if (c == 'n') { if (e_ - p_ >= 4 && !memcmp(p_, "null", 4)) { p_ += 4; out.t = JValue::Null; return true; } return fail("bad literal"); }
        return number(out);
    }
};

}  // namespace ctk
