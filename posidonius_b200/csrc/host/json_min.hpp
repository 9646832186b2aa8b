// json_min.hpp — a small JSON DOM reader/writer for the serde_json images of the reference's integrators.
// Numbers are parsed with strtod (correctly rounded) and written as the shortest decimal that round-trips.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace pbjson {

struct Value;
typedef std::shared_ptr<Value> Ptr;

struct Value {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0.;
    bool is_integer = false;       // written without exponent/fraction in the source
    unsigned long long u = 0;      // exact value when is_integer
    std::string str;
    std::vector<Ptr> arr;
    std::vector<std::pair<std::string, Ptr>> obj;  // insertion order kept

    const Value* find(const std::string& key) const {
        for (const auto& kv : obj) if (kv.first == key) return kv.second.get();
        return nullptr;
    }
    const Value& at(const std::string& key) const {
        const Value* v = find(key);
        if (!v) throw std::runtime_error("missing JSON key '" + key + "'");
        return *v;
    }
    const Value& at(size_t i) const {
        if (kind != Array || i >= arr.size()) throw std::runtime_error("JSON array index out of range");
        return *arr[i];
    }
    double number() const {
        if (kind != Number) throw std::runtime_error("JSON number expected");
        return num;
    }
    bool boolean() const {
        if (kind != Bool) throw std::runtime_error("JSON bool expected");
        return b;
    }
};

class Parser {
  public:
    explicit Parser(const std::string& text) : s(text), p(0) {}
    Ptr parse() {
        Ptr v = value();
        ws();
        if (p != s.size()) fail("trailing characters");
        return v;
    }

  private:
    const std::string& s;
    size_t p;
    [[noreturn]] void fail(const char* what) { throw std::runtime_error(std::string("JSON parse error: ") + what + " at byte " + std::to_string(p)); }
    void ws() { while (p < s.size() && (s[p] == ' ' || s[p] == '\n' || s[p] == '\t' || s[p] == '\r')) p++; }
    Ptr value() {
        ws();
        if (p >= s.size()) fail("unexpected end");
        char c = s[p];
        Ptr v = std::make_shared<Value>();
        if (c == '{') {
            v->kind = Value::Object;
            p++;
            ws();
            if (p < s.size() && s[p] == '}') { p++; return v; }
            for (;;) {
                ws();
                std::string k = string();
                ws();
                if (p >= s.size() || s[p] != ':') fail("':' expected");
                p++;
                v->obj.emplace_back(k, value());
                ws();
                if (p < s.size() && s[p] == ',') { p++; continue; }
                if (p < s.size() && s[p] == '}') { p++; break; }
                fail("',' or '}' expected");
            }
        } else if (c == '[') {
            v->kind = Value::Array;
            p++;
            ws();
            if (p < s.size() && s[p] == ']') { p++; return v; }
            for (;;) {
                v->arr.push_back(value());
                ws();
                if (p < s.size() && s[p] == ',') { p++; continue; }
                if (p < s.size() && s[p] == ']') { p++; break; }
                fail("',' or ']' expected");
            }
        } else if (c == '"') {
            v->kind = Value::String;
            v->str = string();
        } else if (c == 't' && s.compare(p, 4, "true") == 0) { v->kind = Value::Bool; v->b = true; p += 4; }
        else if (c == 'f' && s.compare(p, 5, "false") == 0) { v->kind = Value::Bool; v->b = false; p += 5; }
        else if (c == 'n' && s.compare(p, 4, "null") == 0) { v->kind = Value::Null; p += 4; }
        else if (c == 'N' && s.compare(p, 3, "NaN") == 0) { v->kind = Value::Number; v->num = NAN; p += 3; }   // python json writes NaN
        else {
            const char* start = s.c_str() + p;
            char* end = nullptr;
            double d = strtod(start, &end);
            if (end == start) fail("value expected");
            v->kind = Value::Number;
            v->num = d;
            bool integer = true;
            for (const char* q = start; q < end; q++) if (*q == '.' || *q == 'e' || *q == 'E' || *q == '-') integer = false;
            v->is_integer = integer;
            if (integer) v->u = strtoull(start, nullptr, 10);
            p += (size_t)(end - start);
        }
        return v;
    }
    std::string string() {
        if (p >= s.size() || s[p] != '"') fail("string expected");
        p++;
        std::string out;
        while (p < s.size() && s[p] != '"') {
            if (s[p] == '\\') {
                p++;
                if (p >= s.size()) fail("bad escape");
                char e = s[p];
                switch (e) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': out += '?'; p += 4; break;   // keys of the integrator images are ASCII
                    default: out += e;
                }
                p++;
            } else out += s[p++];
        }
        if (p >= s.size()) fail("unterminated string");
        p++;
        return out;
    }
};

// shortest decimal representation that round-trips (serde_json prints floats with ryu; any round-tripping text is equivalent)
inline std::string format_double(double v) {
    if (std::isnan(v) || std::isinf(v)) return "null";   // serde_json writes null for non-finite floats
    char buf[40];
    for (int prec = 1; prec <= 17; prec++) {
        snprintf(buf, sizeof buf, "%.*g", prec, v);
        if (strtod(buf, nullptr) == v) break;
    }
    std::string s(buf);
    if (s.find('.') == std::string::npos && s.find('e') == std::string::npos && s.find("inf") == std::string::npos) s += ".0";
    return s;
}

// Pretty writer with serde_json's layout (2 spaces, "key": value).
class Writer {
  public:
    std::string out;
    void begin_object() { open('{'); }
    void end_object() { close('}'); }
    void begin_array() { open('['); }
    void end_array() { close(']'); }
    void key(const char* k) { comma(); indent(); out += '"'; out += k; out += "\": "; pending_value = true; }
    void number(double v) { prefix(); out += format_double(v); }
    void integer(unsigned long long v) { prefix(); out += std::to_string(v); }
    void boolean(bool v) { prefix(); out += v ? "true" : "false"; }
    void string(const char* s) { prefix(); out += '"'; out += s; out += '"'; }

  private:
    std::vector<int> counts;
    bool pending_value = false;
    void indent() { out.append(2 * counts.size(), ' '); }
    void comma() {
        if (!counts.empty()) { if (counts.back() > 0) out += ","; out += "\n"; counts.back()++; }
    }
    void prefix() {
        if (pending_value) { pending_value = false; return; }
        comma();
        indent();
    }
    void open(char c) { prefix(); out += c; counts.push_back(0); }
    void close(char c) {
        int n = counts.back();
        counts.pop_back();
        if (n > 0) { out += "\n"; indent(); }
        out += c;
    }
};

}  // namespace pbjson
