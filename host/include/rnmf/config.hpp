// config.hpp -- mirror of src/config.rs: Config<T>{geom, model, residual_iters}, init(), read_lua().
//
// The reference embeds Lua 5.3 through rlua (config.rs:90-158).  Neither Lua nor Rust exists in
// this image, so read_lua() evaluates the subset of Lua that configuration files of this kind use,
// with Lua's own semantics for it: comments (`--`, `--[[ ]]`), `[local] name = expr`,
// numbers (strtod, as Lua), strings, true/false/nil, variables, `+ - * / // % ^`, unary minus,
// parentheses, table constructors `{ k = v, [k] = v, v }`, field access `t.k` / `t[k]`, and calls
// of the registered constructors RealVec2 / UIntVec2 / IntVec2 / Model (config.rs:109-126) plus
// math.sqrt/abs/floor/pi.  A `.lua` case of the reference therefore runs with no edits.
#pragma once

#include <cctype>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "rnmf.hpp"

namespace rnmf {
namespace lua {

struct Value;
using Table = std::map<std::string, Value>;

struct Value {
    enum Kind { Nil, Bool, Number, String, TableK, Vec } kind = Nil;
    bool b = false;
    double num = 0.0;
    std::string str;                    // String, or the constructor name for Vec/Model tables
    std::shared_ptr<Table> table;       // TableK (array part stored under keys "1","2",..)
    std::vector<double> vec;            // Vec: RealVec2 / UIntVec2 / IntVec2 payload

    static Value number(double v) { Value x; x.kind = Number; x.num = v; return x; }
    static Value boolean(bool v) { Value x; x.kind = Bool; x.b = v; return x; }
    static Value string(std::string s) { Value x; x.kind = String; x.str = std::move(s); return x; }
};

class Interpreter {
public:
    std::map<std::string, Value> globals;

    void exec(const std::string& src)
    {
        s_ = src; p_ = 0;
        for (;;) {
            skip();
            if (p_ >= s_.size()) break;
            if (s_[p_] == ';') { ++p_; continue; }
            statement();
        }
    }

private:
    std::string s_;
    std::size_t p_ = 0;

    [[noreturn]] void fail(const std::string& m) const
    {
        std::size_t line = 1;
        for (std::size_t i = 0; i < p_ && i < s_.size(); ++i) line += s_[i] == '\n';
        throw Panic("failed executing lua file: line " + std::to_string(line) + ": " + m);
    }
    void skip()
    {
        for (;;) {
            while (p_ < s_.size() && std::isspace((unsigned char)s_[p_])) ++p_;
            if (p_ + 1 < s_.size() && s_[p_] == '-' && s_[p_ + 1] == '-') {
                if (s_.compare(p_, 4, "--[[") == 0) {
                    const std::size_t e = s_.find("]]", p_ + 4);
                    p_ = e == std::string::npos ? s_.size() : e + 2;
                } else {
                    while (p_ < s_.size() && s_[p_] != '\n') ++p_;
                }
                continue;
            }
            break;
        }
    }
    bool eat(char c)
    {
        skip();
        if (p_ < s_.size() && s_[p_] == c) { ++p_; return true; }
        return false;
    }
    void expect(char c)
    {
        if (!eat(c)) fail(std::string("expected '") + c + "'");
    }
    std::string ident()
    {
        skip();
        std::size_t b = p_;
        if (p_ < s_.size() && (std::isalpha((unsigned char)s_[p_]) || s_[p_] == '_')) {
            while (p_ < s_.size() && (std::isalnum((unsigned char)s_[p_]) || s_[p_] == '_')) ++p_;
        }
        return s_.substr(b, p_ - b);
    }
    void statement()
    {
        std::size_t save = p_;
        std::string name = ident();
        if (name == "local") name = ident();
        if (name.empty()) fail("statement expected");
        skip();
        if (p_ < s_.size() && s_[p_] == '=' && !(p_ + 1 < s_.size() && s_[p_ + 1] == '=')) {
            ++p_;
            globals[name] = expr();
            return;
        }
        // a bare call such as print(...) is evaluated and discarded
        p_ = save;
        (void)expr();
    }
    // precedence (low -> high): + -  |  * / // %  |  unary -  |  ^ (right assoc)
    Value expr() { return additive(); }
    static double need_num(const Value& v, const Interpreter* self)
    {
        if (v.kind != Value::Number) self->fail("arithmetic on a non-number value");
        return v.num;
    }
    Value additive()
    {
        Value l = multiplicative();
        for (;;) {
            skip();
            if (p_ < s_.size() && s_[p_] == '+') { ++p_; Value r = multiplicative(); l = Value::number(need_num(l, this) + need_num(r, this)); }
            else if (p_ < s_.size() && s_[p_] == '-' && !(p_ + 1 < s_.size() && s_[p_ + 1] == '-')) { ++p_; Value r = multiplicative(); l = Value::number(need_num(l, this) - need_num(r, this)); }
            else return l;
        }
    }
    Value multiplicative()
    {
        Value l = unary();
        for (;;) {
            skip();
            if (p_ < s_.size() && s_[p_] == '*') { ++p_; Value r = unary(); l = Value::number(need_num(l, this) * need_num(r, this)); }
            else if (p_ + 1 < s_.size() && s_[p_] == '/' && s_[p_ + 1] == '/') { p_ += 2; Value r = unary(); l = Value::number(std::floor(need_num(l, this) / need_num(r, this))); }
            else if (p_ < s_.size() && s_[p_] == '/') { ++p_; Value r = unary(); l = Value::number(need_num(l, this) / need_num(r, this)); }
            else if (p_ < s_.size() && s_[p_] == '%') { ++p_; Value r = unary(); const double a = need_num(l, this), b = need_num(r, this); l = Value::number(a - std::floor(a / b) * b); }
            else return l;
        }
    }
    Value unary()
    {
        skip();
        if (p_ < s_.size() && s_[p_] == '-' && !(p_ + 1 < s_.size() && s_[p_ + 1] == '-')) { ++p_; Value v = unary(); return Value::number(-need_num(v, this)); }
        return power();
    }
    Value power()
    {
        Value base = postfix();
        skip();
        if (p_ < s_.size() && s_[p_] == '^') { ++p_; Value e = unary(); return Value::number(std::pow(need_num(base, this), need_num(e, this))); }
        return base;
    }
    Value postfix()
    {
        Value v = primary();
        for (;;) {
            skip();
            if (p_ < s_.size() && s_[p_] == '.' && !(p_ + 1 < s_.size() && s_[p_ + 1] == '.')) {
                ++p_;
                v = index(v, ident());
            } else if (p_ < s_.size() && s_[p_] == '[') {
                ++p_;
                Value k = expr();
                expect(']');
                v = index(v, key_of(k));
            } else return v;
        }
    }
    std::string key_of(const Value& k) const
    {
        if (k.kind == Value::String) return k.str;
        if (k.kind == Value::Number) { std::ostringstream o; if (k.num == std::floor(k.num)) o << (long long)k.num; else o << k.num; return o.str(); }
        fail("unsupported table key");
    }
    Value index(const Value& v, const std::string& k) const
    {
        if (v.kind == Value::Vec) {
            const int i = std::atoi(k.c_str());
            if (i >= 1 && (std::size_t)i <= v.vec.size()) return Value::number(v.vec[i - 1]);
            return Value();
        }
        if (v.kind != Value::TableK) fail("attempt to index a non-table value with '" + k + "'");
        auto it = v.table->find(k);
        return it == v.table->end() ? Value() : it->second;
    }
    Value primary()
    {
        skip();
        if (p_ >= s_.size()) fail("unexpected end of file");
        const char c = s_[p_];
        if (c == '(') { ++p_; Value v = expr(); expect(')'); return v; }
        if (c == '{') return table_ctor();
        if (c == '"' || c == '\'') {
            ++p_;
            std::string out;
            while (p_ < s_.size() && s_[p_] != c) { if (s_[p_] == '\\' && p_ + 1 < s_.size()) ++p_; out += s_[p_++]; }
            if (p_ >= s_.size()) fail("unterminated string");
            ++p_;
            return Value::string(out);
        }
        if (std::isdigit((unsigned char)c) || (c == '.' && p_ + 1 < s_.size() && std::isdigit((unsigned char)s_[p_ + 1]))) {
            char* end = nullptr;
            const double v = std::strtod(s_.c_str() + p_, &end);      // Lua's own conversion (lua_str2number = strtod)
            p_ = (std::size_t)(end - s_.c_str());
            return Value::number(v);
        }
        std::string name = ident();
        if (name.empty()) fail(std::string("unexpected character '") + c + "'");
        if (name == "true") return Value::boolean(true);
        if (name == "false") return Value::boolean(false);
        if (name == "nil") return Value();
        skip();
        if (name == "math") return math_lib();
        if (p_ < s_.size() && (s_[p_] == '(' || s_[p_] == '{')) return call(name);
        auto it = globals.find(name);
        return it == globals.end() ? Value() : it->second;
    }
    std::vector<Value> args()
    {
        std::vector<Value> a;
        skip();
        if (p_ < s_.size() && s_[p_] == '{') { a.push_back(table_ctor()); return a; }   // f{...} call sugar
        expect('(');
        if (eat(')')) return a;
        do { a.push_back(expr()); } while (eat(','));
        expect(')');
        return a;
    }
    Value math_lib()
    {
        expect('.');
        const std::string f = ident();
        if (f == "pi") return Value::number(3.141592653589793);
        if (f == "huge") return Value::number(HUGE_VAL);
        auto a = args();
        if (a.size() != 1) fail("math." + f + ": one argument expected");
        const double x = need_num(a[0], this);
        if (f == "sqrt") return Value::number(std::sqrt(x));
        if (f == "abs") return Value::number(std::fabs(x));
        if (f == "floor") return Value::number(std::floor(x));
        if (f == "ceil") return Value::number(std::ceil(x));
        if (f == "sin") return Value::number(std::sin(x));
        if (f == "cos") return Value::number(std::cos(x));
        if (f == "exp") return Value::number(std::exp(x));
        fail("math." + f + " is not in the supported subset");
    }
    Value call(const std::string& name)
    {
        auto a = args();
        // constructors registered by read_lua, config.rs:109-126
        if (name == "RealVec2" || name == "UIntVec2" || name == "IntVec2" || name == "RealVec3" || name == "UIntVec3" || name == "IntVec3") {
            const std::size_t want = name.back() == '2' ? 2 : 3;
            if (a.size() != want) fail(name + ": expected " + std::to_string(want) + " arguments");
            Value v;
            v.kind = Value::Vec;
            v.str = name;
            for (auto& x : a) {
                const double d = need_num(x, this);
                if (name[0] != 'R' && d != std::floor(d)) fail(name + ": integer arguments expected");
                if (name[0] == 'U' && d < 0) fail(name + ": non-negative arguments expected");
                v.vec.push_back(d);
            }
            return v;
        }
        if (name == "Model") {
            if (a.size() != 1 || a[0].kind != Value::TableK) fail("Model: a table argument expected");
            Value v = a[0];
            v.str = "Model";
            return v;
        }
        if (name == "print") return Value();
        fail("call of unknown function '" + name + "'");
    }
    Value table_ctor()
    {
        expect('{');
        Value t;
        t.kind = Value::TableK;
        t.table = std::make_shared<Table>();
        int next = 1;
        for (;;) {
            skip();
            if (eat('}')) break;
            if (p_ < s_.size() && s_[p_] == '[') {
                ++p_;
                Value k = expr();
                expect(']');
                expect('=');
                (*t.table)[key_of(k)] = expr();
            } else {
                const std::size_t save = p_;
                const std::string name = ident();
                skip();
                if (!name.empty() && p_ < s_.size() && s_[p_] == '=' && !(p_ + 1 < s_.size() && s_[p_ + 1] == '=')) {
                    ++p_;
                    (*t.table)[name] = expr();
                } else {
                    p_ = save;
                    (*t.table)[std::to_string(next++)] = expr();
                }
            }
            if (!eat(',') && !eat(';')) { expect('}'); break; }
        }
        return t;
    }
};

// typed getters with the reference's failure style (`.expect("failed reading 'H_far'")`, model.rs:39-54)
inline const Value& field(const Table& t, const std::string& k)
{
    auto it = t.find(k);
    if (it == t.end() || it->second.kind == Value::Nil) throw Panic("failed reading '" + k + "'");
    return it->second;
}
inline Real get_real(const Table& t, const std::string& k)
{
    const Value& v = field(t, k);
    if (v.kind != Value::Number) throw Panic("failed reading '" + k + "'");
    return v.num;
}
inline std::size_t get_usize(const Table& t, const std::string& k)
{
    const Value& v = field(t, k);
    if (v.kind != Value::Number || v.num < 0 || v.num != std::floor(v.num)) throw Panic("failed reading '" + k + "'");
    return (std::size_t)v.num;
}
inline RealVec2 get_realvec2(const Table& t, const std::string& k)
{
    const Value& v = field(t, k);
    if (v.kind != Value::Vec || v.str != "RealVec2") throw Panic("failed reading '" + k + "'");
    return { v.vec[0], v.vec[1] };
}

}  // namespace lua

namespace config {

// config.rs:34-39
struct GeomConfig {
    std::size_t dim = 0;
    RealVec2 length{ 0.0, 0.0 };
    UIntVec2 n_cells{ 0, 0 };
};

// config.rs:15-31.  T must provide `static T from_lua(const lua::Table&)` (UserConfig::lua_constructor)
template <class T>
struct Config {
    GeomConfig geom;
    T model;
    std::size_t residual_iters = 1;
    explicit Config(T user) : model(std::move(user)) {}
};

// config.rs:90-158
template <class T>
Config<T> read_lua(const std::string& lua_loc, T user_model)
{
    Config<T> conf(std::move(user_model));
    std::ifstream in(lua_loc);
    if (!in) throw Panic("Could not read lua file");                       // config.rs:97
    std::stringstream ss;
    ss << in.rdbuf();
    lua::Interpreter L;
    L.exec(ss.str());
    auto g = [&](const char* k) -> const lua::Value* {
        auto it = L.globals.find(k);
        return it == L.globals.end() ? nullptr : &it->second;
    };
    const lua::Value* dim = g("dim");
    if (!dim || dim->kind != lua::Value::Number) throw Panic("failed setting dimensionality");   // config.rs:138-139
    conf.geom.dim = (std::size_t)dim->num;
    if (conf.geom.dim == 2) {
        const lua::Value* len = g("length");
        const lua::Value* nc = g("n_cells");
        if (!len || len->kind != lua::Value::Vec || len->str != "RealVec2") throw Panic("failed reading 'length'");
        if (!nc || nc->kind != lua::Value::Vec || nc->str != "UIntVec2") throw Panic("failed reading 'n_cells'");
        conf.geom.length = { len->vec[0], len->vec[1] };
        conf.geom.n_cells = { (std::size_t)nc->vec[0], (std::size_t)nc->vec[1] };
    } else {
        std::cout << "Error: " << conf.geom.dim << " not yet supported. Exiting" << std::endl;     // config.rs:148-150
    }
    const lua::Value* model = g("model");
    if (!model || model->kind != lua::Value::TableK || model->str != "Model") throw Panic("failed reading model parameters");   // :153
    conf.model = T::from_lua(*model->table);
    return conf;
}

// config.rs:57-86: banner, precision line, the positional Lua path
template <class T>
Config<T> init(const std::vector<std::string>& args, T user_model)
{
    std::cout << "rnmf-b200 (librnmf_b200 abi " << rnmf_abi_version() << ")" << std::endl;
    std::cout << "using double precision" << std::endl;
    if (args.size() <= 1) {
        std::cout << "Error: Lua configuration script not given." << std::endl;
        throw Panic("Lua configuration script not given");
    }
    return read_lua(args[1], std::move(user_model));
}

}  // namespace config
}  // namespace rnmf
