// TEST INFRASTRUCTURE ONLY -- part of oracle/_ref.  The reference pins nlohmann/json v3.11.2 (CMakeLists.txt:28) and uses it for
// exactly two things: writing a flat object of strings / ints with `os << std::setw(4) << j` (img2img_build.cpp:29-50) and
// reading one back with `is >> j`, `j.at(k).get<std::string>()`, `j.at(k).get_to(int&)` (img2img_load.cpp:54-77).  This is
// a flat-object stand-in with the same call surface; the writer reproduces nlohmann's dump(4) layout (one `"key": value`
// per line, four-space indent, insertion order for ordered_json, no trailing newline).
#pragma once

#include <cctype>
#include <initializer_list>
#include <iomanip>
#include <istream>
#include <ostream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace nlohmann {

class value {
public:
    bool isString = false;
    std::string s;
    long long i = 0;
    value() {}
    value(const char* v) : isString(true), s(v) {}
    value(const std::string& v) : isString(true), s(v) {}
    value(int v) : i(v) {}
    value(long long v) : i(v) {}
    template <class T>
    T get() const;
    void get_to(int& out) const {
        if (isString) throw std::runtime_error("json: type must be number, but is string");
        out = (int)i;
    }
    void get_to(std::string& out) const {
        if (!isString) throw std::runtime_error("json: type must be string, but is number");
        out = s;
    }
};
template <>
inline std::string value::get<std::string>() const {
    if (!isString) throw std::runtime_error("json: type must be string, but is number");
    return s;
}
template <>
inline int value::get<int>() const {
    if (isString) throw std::runtime_error("json: type must be number, but is string");
    return (int)i;
}

class ordered_json {
public:
    std::vector<std::pair<std::string, value>> items;
    ordered_json() {}
    ordered_json(std::initializer_list<std::pair<std::string, value>> l) : items(l) {}
    const value& at(const std::string& k) const {
        for (const auto& kv : items)
            if (kv.first == k) return kv.second;
        throw std::out_of_range("json: key '" + k + "' not found");
    }
    static std::string escape(const std::string& in) {
        std::string o;
        for (char ch : in) {
            switch (ch) {
                case '"': o += "\\\""; break;
                case '\\': o += "\\\\"; break;
                case '\n': o += "\\n"; break;
                case '\t': o += "\\t"; break;
                default: o += ch;
            }
        }
        return o;
    }
    std::string dump(int indent) const {
        if (items.empty()) return "{}";
        const std::string pad(indent > 0 ? indent : 0, ' ');
        std::string o = "{";
        for (size_t k = 0; k < items.size(); ++k) {
            o += indent > 0 ? "\n" + pad : "";
            o += "\"" + escape(items[k].first) + "\":" + (indent > 0 ? " " : "");
            o += items[k].second.isString ? "\"" + escape(items[k].second.s) + "\"" : std::to_string(items[k].second.i);
            if (k + 1 < items.size()) o += ",";
        }
        o += indent > 0 ? "\n}" : "}";
        return o;
    }
    void parse(std::istream& is) {
        items.clear();
        auto ws = [&] { while (std::isspace(is.peek())) is.get(); };
        auto expect = [&](char c) {
            ws();
            if (is.get() != c) throw std::runtime_error(std::string("json: parse error, expected '") + c + "'");
        };
        auto str = [&]() {
            expect('"');
            std::string o;
            for (int ch = is.get(); ch != '"'; ch = is.get()) {
                if (ch == EOF) throw std::runtime_error("json: unterminated string");
                if (ch == '\\') {
                    ch = is.get();
                    o += ch == 'n' ? '\n' : ch == 't' ? '\t' : (char)ch;
                } else {
                    o += (char)ch;
                }
            }
            return o;
        };
        expect('{');
        ws();
        if (is.peek() == '}') { is.get(); return; }
        for (;;) {
            std::string key = str();
            expect(':');
            ws();
            value v;
            if (is.peek() == '"') {
                v = value(str());
            } else {
                long long n = 0;
                if (!(is >> n)) throw std::runtime_error("json: parse error, expected a number");
                v = value(n);
            }
            items.emplace_back(std::move(key), std::move(v));
            ws();
            const int c = is.get();
            if (c == '}') break;
            if (c != ',') throw std::runtime_error("json: parse error, expected ',' or '}'");
        }
    }
};
typedef ordered_json json;

inline std::ostream& operator<<(std::ostream& os, const ordered_json& j) {
    const int indent = (int)os.width();
    os.width(0);
    return os << j.dump(indent);
}
inline std::istream& operator>>(std::istream& is, ordered_json& j) {
    j.parse(is);
    return is;
}

}  // namespace nlohmann
