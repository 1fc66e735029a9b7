"""Mirror of src/config.rs for the Python host layer: `read_lua(path) -> poisson.Config`.

The reference embeds Lua 5.3 through rlua (config.rs:90-158); this image has neither Lua nor Rust, so,
like host/include/rnmf/config.hpp (the C++ twin, against which the CPU tests compare this reader), the
subset of Lua that case files of this kind use is evaluated with Lua's own semantics: comments
(`--`, `--[[ ]]`), `[local] name = expr`, integer / float numbers (`/` and `^` give floats, `//` and `%`
floor), strings, true / false / nil, variables, `+ - * / // % ^`, unary minus, parentheses, table
constructors `{ k = v, [k] = v, v }`, field access `t.k` / `t[k]`, the registered constructors RealVec2 /
UIntVec2 / IntVec2 / Model (config.rs:109-126) and math.sqrt / abs / floor / pi.  A `.lua` case of the
reference therefore runs with no edits.  Failures raise `ConfigPanic` with the reference's messages
(it panics: config.rs:97,135,153; model.rs:39-54)."""
from __future__ import annotations

import math
from typing import Any, Dict, List


class ConfigPanic(Exception):
    """the reference panics on a bad case file; the Python mirror raises"""


class LuaVec:
    def __init__(self, ctor: str, vals: List[float]):
        self.ctor, self.vals = ctor, vals


class LuaTable(dict):
    ctor = ""            # "Model" for Model({...})


class _Interp:
    def __init__(self, src: str):
        self.s, self.p = src, 0
        self.globals: Dict[str, Any] = {}

    # ---- lexing helpers ---------------------------------------------------------------------
    def fail(self, msg: str):
        line = 1 + self.s.count("\n", 0, min(self.p, len(self.s)))
        raise ConfigPanic(f"failed executing lua file: line {line}: {msg}")

    def skip(self):
        s = self.s
        while True:
            while self.p < len(s) and s[self.p].isspace():
                self.p += 1
            if s.startswith("--", self.p):
                if s.startswith("--[[", self.p):
                    e = s.find("]]", self.p + 4)
                    self.p = len(s) if e < 0 else e + 2
                else:
                    e = s.find("\n", self.p)
                    self.p = len(s) if e < 0 else e
                continue
            return

    def eat(self, tok: str) -> bool:
        self.skip()
        if self.s.startswith(tok, self.p):
            self.p += len(tok)
            return True
        return False

    def expect(self, tok: str):
        if not self.eat(tok):
            self.fail(f"expected '{tok}'")

    def name(self) -> str:
        self.skip()
        b = self.p
        while self.p < len(self.s) and (self.s[self.p].isalnum() or self.s[self.p] == "_") and not (self.p == b and self.s[b].isdigit()):
            self.p += 1
        return self.s[b:self.p]

    # ---- statements ---------------------------------------------------------------------------
    def run(self):
        while True:
            self.skip()
            if self.p >= len(self.s):
                return
            if self.s[self.p] == ";":
                self.p += 1
                continue
            n = self.name()
            if n == "local":
                n = self.name()
            if not n:
                self.fail("statement expected")
            self.skip()
            if self.s.startswith("==", self.p) or not self.eat("="):
                self.fail("expected '='")
            self.skip()
            if self.s.startswith("=", self.p):
                self.fail("unexpected character '='")
            self.globals[n] = self.expr()

    # ---- expressions (Lua precedence: + - < * / // % < unary < ^) ------------------------------
    def num(self, v):
        if isinstance(v, bool) or not isinstance(v, (int, float)):
            self.fail("arithmetic on a non-number value")
        return v

    def expr(self):
        v = self.term()
        while True:
            self.skip()
            if self.s.startswith("--", self.p):
                return v
            if self.eat("+"):
                v = self.num(v) + self.num(self.term())
            elif self.eat("-"):
                v = self.num(v) - self.num(self.term())
            else:
                return v

    def term(self):
        v = self.unary()
        while True:
            if self.eat("//"):
                r = self.num(self.unary())
                v = self.num(v) // r if isinstance(v, int) and isinstance(r, int) else float(math.floor(self.num(v) / r))
            elif self.eat("*"):
                v = self.num(v) * self.num(self.unary())
            elif self.eat("/"):
                v = float(self.num(v)) / float(self.num(self.unary()))
            elif self.eat("%"):
                r = self.num(self.unary())
                v = self.num(v) - math.floor(self.num(v) / r) * r
            else:
                return v

    def unary(self):
        self.skip()
        if not self.s.startswith("--", self.p) and self.eat("-"):
            return -self.num(self.unary())
        return self.power()

    def power(self):
        base = self.suffixed()
        if self.eat("^"):
            return float(self.num(base)) ** float(self.num(self.unary()))       # right associative, binds tighter than unary minus on its left
        return base

    def suffixed(self):
        v = self.primary()
        while True:
            self.skip()
            if self.p < len(self.s) and self.s[self.p] == "." and not self.s.startswith("..", self.p):
                self.p += 1
                v = self.index(v, self.name())
            elif self.eat("["):
                k = self.expr()
                self.expect("]")
                v = self.index(v, self.key(k))
            else:
                return v

    def key(self, k) -> str:
        if isinstance(k, str):
            return k
        if isinstance(k, (int, float)) and not isinstance(k, bool) and float(k) == math.floor(k):
            return str(int(k))
        self.fail("unsupported table key")

    def index(self, v, k: str):
        if not isinstance(v, LuaTable):
            self.fail(f"attempt to index a non-table value with '{k}'")
        return v.get(k)

    def primary(self):
        self.skip()
        if self.p >= len(self.s):
            self.fail("unexpected end of file")
        c = self.s[self.p]
        if c == "(":
            self.p += 1
            v = self.expr()
            self.expect(")")
            return v
        if c == "{":
            return self.table()
        if c in "\"'":
            e = self.s.find(c, self.p + 1)
            if e < 0:
                self.fail("unterminated string")
            v = self.s[self.p + 1:e]
            self.p = e + 1
            return v
        if c.isdigit() or (c == "." and self.p + 1 < len(self.s) and self.s[self.p + 1].isdigit()):
            return self.number()
        n = self.name()
        if not n:
            self.fail(f"unexpected character '{c}'")
        if n == "true":
            return True
        if n == "false":
            return False
        if n == "nil":
            return None
        if n == "math":
            return self.math()
        self.skip()
        if self.p < len(self.s) and self.s[self.p] in "({":
            return self.call(n)
        return self.globals.get(n)

    def number(self):
        b = self.p
        s = self.s
        if s.startswith(("0x", "0X"), b):
            self.p += 2
            while self.p < len(s) and s[self.p] in "0123456789abcdefABCDEF":
                self.p += 1
            return int(s[b:self.p], 16)
        is_float = False
        while self.p < len(s) and (s[self.p].isdigit() or s[self.p] == "."):
            is_float |= s[self.p] == "."
            self.p += 1
        if self.p < len(s) and s[self.p] in "eE":
            is_float = True
            self.p += 1
            if self.p < len(s) and s[self.p] in "+-":
                self.p += 1
            while self.p < len(s) and s[self.p].isdigit():
                self.p += 1
        txt = s[b:self.p]
        try:
            return float(txt) if is_float else int(txt)
        except ValueError:
            self.fail(f"malformed number '{txt}'")

    def args(self) -> List[Any]:
        self.skip()
        if self.p < len(self.s) and self.s[self.p] == "{":        # f{...}
            return [self.table()]
        self.expect("(")
        out: List[Any] = []
        if not self.eat(")"):
            while True:
                out.append(self.expr())
                if not self.eat(","):
                    break
            self.expect(")")
        return out

    def math(self):
        self.expect(".")
        f = self.name()
        if f == "pi":
            return math.pi
        if f == "huge":
            return math.inf
        a = self.args()
        if len(a) != 1:
            self.fail(f"math.{f}: one argument expected")
        x = self.num(a[0])
        if f == "sqrt":
            return math.sqrt(x)
        if f == "abs":
            return abs(x)
        if f == "floor":
            return math.floor(x)
        if f == "ceil":
            return math.ceil(x)
        if f in ("sin", "cos", "exp"):
            return getattr(math, f)(x)
        self.fail(f"math.{f} is not in the supported subset")

    def call(self, n: str):
        a = self.args()
        if n in ("RealVec2", "UIntVec2", "IntVec2", "RealVec3", "UIntVec3", "IntVec3"):     # config.rs:109-126
            want = int(n[-1])
            if len(a) != want:
                self.fail(f"{n}: expected {want} arguments")
            vals = []
            for x in a:
                d = float(self.num(x))
                if n[0] != "R" and d != math.floor(d):
                    self.fail(f"{n}: integer arguments expected")
                if n[0] == "U" and d < 0:
                    self.fail(f"{n}: non-negative arguments expected")
                vals.append(d)
            return LuaVec(n, vals)
        if n == "Model":
            if len(a) != 1 or not isinstance(a[0], LuaTable):
                self.fail("Model: a table argument expected")
            a[0].ctor = "Model"
            return a[0]
        if n == "print":
            return None
        self.fail(f"call of unknown function '{n}'")

    def table(self) -> LuaTable:
        self.expect("{")
        t = LuaTable()
        nxt = 1
        while True:
            self.skip()
            if self.eat("}"):
                break
            if self.eat("["):
                k = self.key(self.expr())
                self.expect("]")
                self.expect("=")
                t[k] = self.expr()
            else:
                save = self.p
                n = self.name()
                self.skip()
                if n and self.s.startswith("=", self.p) and not self.s.startswith("==", self.p):
                    self.p += 1
                    t[n] = self.expr()
                else:
                    self.p = save
                    t[str(nxt)] = self.expr()
                    nxt += 1
            if not self.eat(",") and not self.eat(";"):
                self.expect("}")
                break
        return t


def _get(t: LuaTable, k: str):
    v = t.get(k)
    if v is None:
        raise ConfigPanic(f"failed reading '{k}'")                    # model.rs:39-54 `.expect("failed reading '...'")`
    return v


def _real(t, k) -> float:
    v = _get(t, k)
    if isinstance(v, bool) or not isinstance(v, (int, float)):
        raise ConfigPanic(f"failed reading '{k}'")
    return float(v)


def _uint(t, k) -> int:
    v = _get(t, k)
    if isinstance(v, bool) or not isinstance(v, (int, float)) or v < 0 or float(v) != math.floor(v):
        raise ConfigPanic(f"failed reading '{k}'")
    return int(v)


def _vec2(t, k):
    v = _get(t, k)
    if not isinstance(v, LuaVec) or v.ctor != "RealVec2":
        raise ConfigPanic(f"failed reading '{k}'")
    return (v.vals[0], v.vals[1])


def read_lua(path: str):
    """config.rs:90-158 read_lua + model.rs:36-57 UserModel::lua_constructor -> poisson.Config"""
    from .poisson import Config, GeomConfig, UserModel
    try:
        with open(path) as f:
            src = f.read()
    except OSError:
        raise ConfigPanic("Could not read lua file")                  # config.rs:97
    it = _Interp(src)
    it.run()
    g = it.globals
    dim = g.get("dim")
    if isinstance(dim, bool) or not isinstance(dim, (int, float)):
        raise ConfigPanic("failed setting dimensionality")            # config.rs:138-139
    length, n_cells = g.get("length"), g.get("n_cells")
    if not isinstance(length, LuaVec) or length.ctor != "RealVec2":
        raise ConfigPanic("failed reading 'length'")
    if not isinstance(n_cells, LuaVec) or n_cells.ctor != "UIntVec2":
        raise ConfigPanic("failed reading 'n_cells'")
    model = g.get("model")
    if not isinstance(model, LuaTable) or model.ctor != "Model":
        raise ConfigPanic("failed reading model parameters")          # config.rs:153
    um = UserModel(h_far=_vec2(model, "H_far"), bubble_centre=_vec2(model, "bubble_centre"),
                   bubble_radius=_real(model, "bubble_radius"), mu=_vec2(model, "mu"), n_iter=_uint(model, "n_iter"),
                   n_sub_iter=_uint(model, "n_sub_iter"), relax=_real(model, "relax"), tol=_real(model, "tol"))
    return Config(GeomConfig(int(dim), tuple(length.vals), tuple(int(v) for v in n_cells.vals)), um, 1)
