// Minimal JSON DOM (RFC 8259) for the Blender-exported scene files (reference docs/scene_format.md).
// Semantics follow JSON.parse as used at reference js/ui-controller.js:248-253: duplicate keys keep the LAST
// value, numbers are doubles.
#pragma once
#include <charconv>
#include <cmath>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

namespace brtjson {

enum Kind { NUL, BOOL, NUM, STR, ARR, OBJ };

struct Value {
    Kind kind = NUL;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Value> arr;
    std::vector<std::pair<std::string, Value>> obj;

    const Value* get(const char* key) const {            // undefined -> nullptr
        if (kind != OBJ) return nullptr;
        for (size_t i = obj.size(); i-- > 0;) if (obj[i].first == key) return &obj[i].second;
        return nullptr;
    }
    bool is_array() const { return kind == ARR; }
    bool is_object() const { return kind == OBJ; }
    bool is_string() const { return kind == STR; }
};

class Parser {
public:
    Parser(const char* s, size_t n) : p_(s), end_(s + n) {}
    bool parse(Value& out, std::string& err) {
        ws();
        if (!value(out, 0)) { err = err_.empty() ? "invalid JSON" : err_; return false; }
        ws();
        if (p_ != end_) { err = "trailing characters after JSON value"; return false; }
        return true;
    }
private:
    const char* p_; const char* end_; std::string err_;
    void ws() { while (p_ < end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) p_++; }
    bool fail(const char* m) { if (err_.empty()) err_ = std::string(m) + " at byte " + std::to_string((long long)(end_ - p_)) + " from end"; return false; }
    bool lit(const char* w) { size_t n = strlen(w); if ((size_t)(end_ - p_) >= n && memcmp(p_, w, n) == 0) { p_ += n; return true; } return false; }
    bool value(Value& v, int depth) {
        if (depth > 256) return fail("nesting too deep");
        if (p_ >= end_) return fail("unexpected end");
        char c = *p_;
        if (c == '{') return object(v, depth);
        if (c == '[') return array(v, depth);
        if (c == '"') { v.kind = STR; return string(v.str); }
        if (c == 't') { if (!lit("true")) return fail("bad literal"); v.kind = BOOL; v.b = true; return true; }
        if (c == 'f') { if (!lit("false")) return fail("bad literal"); v.kind = BOOL; v.b = false; return true; }
        if (c == 'n') { if (!lit("null")) return fail("bad literal"); v.kind = NUL; return true; }
        return number(v);
    }
    bool number(Value& v) {
        const char* s = p_;
        if (p_ < end_ && *p_ == '-') p_++;
        if (p_ >= end_ || *p_ < '0' || *p_ > '9') return fail("bad number");
        if (*p_ == '0') p_++; else while (p_ < end_ && *p_ >= '0' && *p_ <= '9') p_++;
        if (p_ < end_ && *p_ == '.') { p_++; if (p_ >= end_ || *p_ < '0' || *p_ > '9') return fail("bad number"); while (p_ < end_ && *p_ >= '0' && *p_ <= '9') p_++; }
        if (p_ < end_ && (*p_ == 'e' || *p_ == 'E')) {
            p_++; if (p_ < end_ && (*p_ == '+' || *p_ == '-')) p_++;
            if (p_ >= end_ || *p_ < '0' || *p_ > '9') return fail("bad number");
            while (p_ < end_ && *p_ >= '0' && *p_ <= '9') p_++;
        }
        double d = 0;
        auto r = std::from_chars(s, p_, d);
        if (r.ec == std::errc::result_out_of_range) d = (*s == '-') ? -HUGE_VAL : HUGE_VAL;   // JSON.parse("1e999") = Infinity
        else if (r.ec != std::errc()) return fail("bad number");
        v.kind = NUM; v.num = d;
        return true;
    }
    static void utf8(std::string& o, unsigned cp) {
        if (cp < 0x80) o += (char)cp;
        else if (cp < 0x800) { o += (char)(0xC0 | (cp >> 6)); o += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { o += (char)(0xE0 | (cp >> 12)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
        else { o += (char)(0xF0 | (cp >> 18)); o += (char)(0x80 | ((cp >> 12) & 0x3F)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
    }
    bool hex4(unsigned& out) {
        if (end_ - p_ < 4) return false;
        out = 0;
        for (int i = 0; i < 4; i++) {
            char c = *p_++; out <<= 4;
            if (c >= '0' && c <= '9') out |= c - '0'; else if (c >= 'a' && c <= 'f') out |= c - 'a' + 10;
            else if (c >= 'A' && c <= 'F') out |= c - 'A' + 10; else return false;
        }
        return true;
    }
    bool string(std::string& out) {
        p_++;  // opening quote
        out.clear();
        while (p_ < end_) {
            unsigned char c = (unsigned char)*p_++;
            if (c == '"') return true;
            if (c < 0x20) return fail("control character in string");
            if (c != '\\') { out += (char)c; continue; }
            if (p_ >= end_) break;
            char e = *p_++;
            switch (e) {
            case '"': out += '"'; break; case '\\': out += '\\'; break; case '/': out += '/'; break;
            case 'b': out += '\b'; break; case 'f': out += '\f'; break; case 'n': out += '\n'; break;
            case 'r': out += '\r'; break; case 't': out += '\t'; break;
            case 'u': {
                unsigned cp; if (!hex4(cp)) return fail("bad \\u escape");
                if (cp >= 0xD800 && cp < 0xDC00 && end_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') {
                    const char* save = p_; p_ += 2; unsigned lo;
                    if (hex4(lo) && lo >= 0xDC00 && lo < 0xE000) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00); else p_ = save;
                }
                utf8(out, cp); break;
            }
            default: return fail("bad escape");
            }
        }
        return fail("unterminated string");
    }
    bool array(Value& v, int depth) {
        v.kind = ARR; p_++; ws();
        if (p_ < end_ && *p_ == ']') { p_++; return true; }
        for (;;) {
            v.arr.emplace_back();
            ws();
            if (!value(v.arr.back(), depth + 1)) return false;
            ws();
            if (p_ >= end_) return fail("unterminated array");
            if (*p_ == ',') { p_++; continue; }
            if (*p_ == ']') { p_++; return true; }
            return fail("expected , or ]");
        }
    }
    bool object(Value& v, int depth) {
        v.kind = OBJ; p_++; ws();
        if (p_ < end_ && *p_ == '}') { p_++; return true; }
        for (;;) {
            ws();
            if (p_ >= end_ || *p_ != '"') return fail("expected string key");
            std::string k; if (!string(k)) return false;
            ws();
            if (p_ >= end_ || *p_ != ':') return fail("expected :");
            p_++; ws();
            v.obj.emplace_back(std::move(k), Value());
            if (!value(v.obj.back().second, depth + 1)) return false;
            ws();
            if (p_ >= end_) return fail("unterminated object");
            if (*p_ == ',') { p_++; continue; }
            if (*p_ == '}') { p_++; return true; }
            return fail("expected , or }");
        }
    }
};

inline bool parse(const char* s, size_t n, Value& out, std::string& err) { Parser p(s, n); return p.parse(out, err); }

}  // namespace brtjson
