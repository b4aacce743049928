// json.hpp -- minimal JSON reader for config.json (objects, arrays, numbers, strings, bools, null).
#pragma once
#include <cmath>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace aha {

struct Json {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Json> arr;
    std::map<std::string, Json> obj;

    bool has(const std::string& k) const { return kind == Obj && obj.count(k) && obj.at(k).kind != Null; }
    const Json& at(const std::string& k) const {
        if (kind != Obj || !obj.count(k)) throw std::runtime_error("config.json: missing key '" + k + "'");
        return obj.at(k);
    }
    double number(const std::string& k) const {
        const Json& v = at(k);
        if (v.kind != Num) throw std::runtime_error("config.json: key '" + k + "' is not a number");
        return v.num;
    }
    double number_or(const std::string& k, double d) const { return has(k) && obj.at(k).kind == Num ? obj.at(k).num : d; }
    int integer(const std::string& k) const { return (int)std::llround(number(k)); }
    int integer_or(const std::string& k, int d) const { return has(k) ? integer(k) : d; }
    bool boolean_or(const std::string& k, bool d) const { return has(k) && obj.at(k).kind == Bool ? obj.at(k).b : d; }
    std::string string_or(const std::string& k, const std::string& d) const {
        return has(k) && obj.at(k).kind == Str ? obj.at(k).str : d;
    }
    std::vector<int> int_array(const std::string& k) const {
        const Json& v = at(k);
        if (v.kind != Arr) throw std::runtime_error("config.json: key '" + k + "' is not an array");
        std::vector<int> r;
        for (auto& e : v.arr) r.push_back((int)std::llround(e.num));
        return r;
    }
};

class JsonParser {
  public:
    explicit JsonParser(const std::string& s) : s_(s) {}
    Json parse() {
        Json v = value();
        ws();
        if (p_ != s_.size()) fail("trailing characters");
        return v;
    }

  private:
    const std::string& s_;
    size_t p_ = 0;
    [[noreturn]] void fail(const std::string& m) { throw std::runtime_error("config.json parse error at " + std::to_string(p_) + ": " + m); }
    void ws() { while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\n' || s_[p_] == '\t' || s_[p_] == '\r')) ++p_; }
    Json value() {
        ws();
        if (p_ >= s_.size()) fail("unexpected end");
        char c = s_[p_];
        Json v;
        if (c == '{') {
            v.kind = Json::Obj; ++p_; ws();
            if (s_[p_] == '}') { ++p_; return v; }
            for (;;) {
                ws(); std::string k = string(); ws();
                if (s_[p_] != ':') fail("expected ':'");
                ++p_; v.obj[k] = value(); ws();
                if (s_[p_] == ',') { ++p_; continue; }
                if (s_[p_] == '}') { ++p_; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            v.kind = Json::Arr; ++p_; ws();
            if (s_[p_] == ']') { ++p_; return v; }
            for (;;) {
                v.arr.push_back(value()); ws();
                if (s_[p_] == ',') { ++p_; continue; }
                if (s_[p_] == ']') { ++p_; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            v.kind = Json::Str; v.str = string();
        } else if (s_.compare(p_, 4, "true") == 0) { v.kind = Json::Bool; v.b = true; p_ += 4;
        } else if (s_.compare(p_, 5, "false") == 0) { v.kind = Json::Bool; v.b = false; p_ += 5;
        } else if (s_.compare(p_, 4, "null") == 0) { v.kind = Json::Null; p_ += 4;
        } else if (s_.compare(p_, 8, "Infinity") == 0) { v.kind = Json::Num; v.num = INFINITY; p_ += 8;
        } else if (s_.compare(p_, 3, "NaN") == 0) { v.kind = Json::Num; v.num = NAN; p_ += 3;
        } else {
            size_t e = p_;
            while (e < s_.size() && (isdigit((unsigned char)s_[e]) || s_[e] == '-' || s_[e] == '+' || s_[e] == '.' || s_[e] == 'e' || s_[e] == 'E')) ++e;
            if (e == p_) fail("unexpected character");
            v.kind = Json::Num; v.num = std::stod(s_.substr(p_, e - p_)); p_ = e;
        }
        return v;
    }
    std::string string() {
        if (s_[p_] != '"') fail("expected string");
        ++p_;
        std::string r;
        while (p_ < s_.size() && s_[p_] != '"') {
            if (s_[p_] == '\\' && p_ + 1 < s_.size()) {
                char n = s_[p_ + 1];
                switch (n) {
                    case 'n': r += '\n'; break; case 't': r += '\t'; break; case 'r': r += '\r'; break;
                    case 'u': r += '?'; p_ += 4; break;
                    default: r += n;
                }
                p_ += 2;
            } else r += s_[p_++];
        }
        if (p_ >= s_.size()) fail("unterminated string");
        ++p_;
        return r;
    }
};

}  // namespace aha
