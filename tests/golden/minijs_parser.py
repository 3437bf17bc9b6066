"""TEST INFRASTRUCTURE ONLY — tokenizer + parser for the subset of ECMAScript 2020 that the reference's compute modules use
(js/rng.js … js/planet-worker.js: modules, classes with getters, arrow functions, destructuring, template literals, spread,
optional chaining, `??`, for-of / for-in, switch, try/catch).  Produces a tuple-shaped AST consumed by minijs.py.  Not a general
JavaScript parser: no regular-expression literals, generators, async functions, labels or tagged templates — the parser raises
on what it does not know instead of guessing.

Why it exists: the reference is browser JavaScript and this image holds no JavaScript runtime, so the reference itself could not
be executed to pin the oracle.  With this parser and the evaluator in minijs.py the UNMODIFIED files under /root/reference/js are
executed (in the build container only) to produce the golden vectors committed under tests/golden/ (make_reference_vectors.py).
"""
from __future__ import annotations

PUNCT = sorted([
    ">>>=", "...", "===", "!==", "**=", "<<=", ">>=", ">>>", "&&=", "||=", "??=", "=>", "==", "!=", "<=", ">=", "&&", "||", "??", "?.",
    "++", "--", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "**", "<<", ">>", "{", "}", "(", ")", "[", "]", ";", ",", "<", ">", "+",
    "-", "*", "/", "%", "&", "|", "^", "!", "~", "?", ":", "=", "."], key=len, reverse=True)
KEYWORDS = {"var", "let", "const", "function", "return", "if", "else", "for", "while", "do", "break", "continue", "new", "delete", "typeof",
            "instanceof", "in", "of", "class", "extends", "super", "this", "null", "undefined", "true", "false", "import", "export", "from",
            "as", "default", "try", "catch", "finally", "throw", "switch", "case", "void", "get", "set", "static"}
# words that are only keywords in some positions and may otherwise be identifiers / property names
CONTEXTUAL = {"of", "from", "as", "get", "set", "static", "undefined"}


class Tok:
    __slots__ = ("kind", "val", "line", "nl")       # kind: num | str | tpl | id | kw | punct | eof ; nl: newline before the token

    def __init__(self, kind, val, line, nl):
        self.kind, self.val, self.line, self.nl = kind, val, line, nl

    def __repr__(self):
        return f"{self.kind}:{self.val!r}@{self.line}"


class JSSyntaxError(Exception):
    pass


_ESC = {"n": "\n", "t": "\t", "r": "\r", "b": "\b", "f": "\f", "v": "\v", "0": "\0", "\n": ""}


def _read_escape(src, i):
    c = src[i]
    if c == "u":
        if src[i + 1] == "{":
            j = src.index("}", i)
            return chr(int(src[i + 2:j], 16)), j + 1
        return chr(int(src[i + 1:i + 5], 16)), i + 5
    if c == "x":
        return chr(int(src[i + 1:i + 3], 16)), i + 3
    return _ESC.get(c, c), i + 1


def tokenize(src: str, fname: str = "<js>"):
    toks, i, n, line, nl = [], 0, len(src), 1, False
    while i < n:
        c = src[i]
        if c == "\n":
            line += 1; nl = True; i += 1; continue
        if c in " \t\r\ufeff\xa0":
            i += 1; continue
        if c == "/" and i + 1 < n and src[i + 1] == "/":
            while i < n and src[i] != "\n":
                i += 1
            continue
        if c == "/" and i + 1 < n and src[i + 1] == "*":
            j = src.index("*/", i + 2)
            if "\n" in src[i:j]:
                nl = True
            line += src.count("\n", i, j)
            i = j + 2; continue
        if c.isdigit() or (c == "." and i + 1 < n and src[i + 1].isdigit()):
            j = i
            if c == "0" and i + 1 < n and src[i + 1] in "xX":
                j = i + 2
                while j < n and src[j] in "0123456789abcdefABCDEF":
                    j += 1
                val = float(int(src[i + 2:j], 16))
            elif c == "0" and i + 1 < n and src[i + 1] in "bB":
                j = i + 2
                while j < n and src[j] in "01":
                    j += 1
                val = float(int(src[i + 2:j], 2))
            else:
                while j < n and (src[j].isdigit() or src[j] == "_"):
                    j += 1
                if j < n and src[j] == ".":
                    j += 1
                    while j < n and src[j].isdigit():
                        j += 1
                if j < n and src[j] in "eE":
                    k = j + 1
                    if k < n and src[k] in "+-":
                        k += 1
                    if k < n and src[k].isdigit():
                        j = k
                        while j < n and src[j].isdigit():
                            j += 1
                val = float(src[i:j].replace("_", ""))
            if j < n and (src[j].isalpha() or src[j] == "_"):
                raise JSSyntaxError(f"{fname}:{line}: bad numeric literal")
            toks.append(Tok("num", val, line, nl)); nl = False; i = j; continue
        if c in "'\"":
            j, out = i + 1, []
            while src[j] != c:
                if src[j] == "\\":
                    s, j = _read_escape(src, j + 1)
                    out.append(s)
                else:
                    if src[j] == "\n":
                        raise JSSyntaxError(f"{fname}:{line}: unterminated string")
                    out.append(src[j]); j += 1
            toks.append(Tok("str", "".join(out), line, nl)); nl = False; i = j + 1; continue
        if c == "`":
            # template literal → ('tpl', [str, exprSource, str, …]) with the expression sources kept for the parser
            j, parts, cur, l0 = i + 1, [], [], line
            while src[j] != "`":
                if src[j] == "\\":
                    s, j = _read_escape(src, j + 1)
                    cur.append(s)
                elif src[j] == "$" and src[j + 1] == "{":
                    depth, k = 1, j + 2
                    while depth:
                        if src[k] == "{":
                            depth += 1
                        elif src[k] == "}":
                            depth -= 1
                        elif src[k] in "'\"`":
                            q = src[k]; k += 1
                            while src[k] != q:
                                k += 2 if src[k] == "\\" else 1
                        k += 1
                    parts.append("".join(cur)); cur = []
                    parts.append((src[j + 2:k - 1], line))
                    line += src.count("\n", j, k)
                    j = k
                else:
                    if src[j] == "\n":
                        line += 1
                    cur.append(src[j]); j += 1
            parts.append("".join(cur))
            toks.append(Tok("tpl", parts, l0, nl)); nl = False; i = j + 1; continue
        if c.isalpha() or c in "_$" or ord(c) > 127:
            j = i + 1
            while j < n and (src[j].isalnum() or src[j] in "_$" or ord(src[j]) > 127):
                j += 1
            w = src[i:j]
            toks.append(Tok("kw" if w in KEYWORDS else "id", w, line, nl)); nl = False; i = j; continue
        for p in PUNCT:
            if src.startswith(p, i):
                # `?.` followed by a digit is a conditional with a decimal literal (a ?.5 : 1)
                if p == "?." and i + 2 < n and src[i + 2].isdigit():
                    continue
                toks.append(Tok("punct", p, line, nl)); nl = False; i += len(p); break
        else:
            raise JSSyntaxError(f"{fname}:{line}: unexpected character {c!r}")
    toks.append(Tok("eof", None, line, True))
    return toks


ASSIGN_OPS = {"=", "+=", "-=", "*=", "/=", "%=", "**=", "<<=", ">>=", ">>>=", "&=", "|=", "^=", "&&=", "||=", "??="}
BIN_PREC = {"??": 1, "||": 2, "&&": 3, "|": 4, "^": 5, "&": 6, "==": 7, "!=": 7, "===": 7, "!==": 7, "<": 8, ">": 8, "<=": 8, ">=": 8,
            "instanceof": 8, "in": 8, "<<": 9, ">>": 9, ">>>": 9, "+": 10, "-": 10, "*": 11, "/": 11, "%": 11, "**": 12}


class Parser:
    def __init__(self, src: str, fname: str = "<js>"):
        self.fname = fname
        self.t = tokenize(src, fname)
        self.i = 0
        self.no_in = False

    # ---- token helpers ----
    @property
    def tok(self):
        return self.t[self.i]

    def peek(self, k=1):
        return self.t[min(self.i + k, len(self.t) - 1)]

    def err(self, msg):
        raise JSSyntaxError(f"{self.fname}:{self.tok.line}: {msg} (at {self.tok!r})")

    def is_p(self, v):
        t = self.tok
        return t.kind == "punct" and t.val == v

    def is_kw(self, v):
        t = self.tok
        return t.kind == "kw" and t.val == v

    def eat_p(self, v):
        if self.is_p(v):
            self.i += 1
            return True
        return False

    def eat_kw(self, v):
        if self.is_kw(v):
            self.i += 1
            return True
        return False

    def expect_p(self, v):
        if not self.eat_p(v):
            self.err(f"expected {v!r}")

    def expect_kw(self, v):
        if not self.eat_kw(v):
            self.err(f"expected {v!r}")

    def ident(self):
        t = self.tok
        if t.kind == "id" or (t.kind == "kw" and t.val in CONTEXTUAL):
            self.i += 1
            return t.val
        self.err("expected identifier")

    def prop_name(self):
        t = self.tok
        if t.kind in ("id", "kw"):
            self.i += 1
            return t.val
        if t.kind == "str":
            self.i += 1
            return t.val
        if t.kind == "num":
            self.i += 1
            from .minijs import number_to_string
            return number_to_string(t.val)
        self.err("expected property name")

    def semicolon(self):
        if self.eat_p(";"):
            return
        if self.is_p("}") or self.tok.kind == "eof" or self.tok.nl:
            return
        self.err("expected ';'")

    # ---- program ----
    def parse_program(self):
        body = []
        while self.tok.kind != "eof":
            body.append(self.statement())
        return ("program", body)

    # ---- statements ----
    def statement(self):
        t = self.tok
        line = t.line
        if t.kind == "punct":
            if t.val == "{":
                return self.block()
            if t.val == ";":
                self.i += 1
                return ("empty",)
        if t.kind == "kw":
            v = t.val
            if v in ("var", "let", "const"):
                d = self.var_decl()
                self.semicolon()
                return d
            if v == "function":
                self.i += 1
                name = self.ident()
                return ("funcdecl", name, self.function_rest(name, False), line)
            if v == "class":
                return self.class_decl()
            if v == "if":
                self.i += 1
                self.expect_p("(")
                test = self.expression()
                self.expect_p(")")
                cons = self.statement()
                alt = self.statement() if self.eat_kw("else") else None
                return ("if", test, cons, alt)
            if v == "for":
                return self.for_statement()
            if v == "while":
                self.i += 1
                self.expect_p("(")
                test = self.expression()
                self.expect_p(")")
                return ("while", test, self.statement())
            if v == "do":
                self.i += 1
                body = self.statement()
                self.expect_kw("while")
                self.expect_p("(")
                test = self.expression()
                self.expect_p(")")
                self.eat_p(";")
                return ("dowhile", body, test)
            if v == "return":
                self.i += 1
                arg = None
                if not (self.is_p(";") or self.is_p("}") or self.tok.kind == "eof" or self.tok.nl):
                    arg = self.expression()
                self.semicolon()
                return ("return", arg)
            if v == "break":
                self.i += 1
                if self.tok.kind == "id" and not self.tok.nl:
                    self.err("labelled break is not supported")
                self.semicolon()
                return ("break",)
            if v == "continue":
                self.i += 1
                if self.tok.kind == "id" and not self.tok.nl:
                    self.err("labelled continue is not supported")
                self.semicolon()
                return ("continue",)
            if v == "throw":
                self.i += 1
                arg = self.expression()
                self.semicolon()
                return ("throw", arg, line)
            if v == "try":
                self.i += 1
                block = self.block()
                param = handler = final = None
                if self.eat_kw("catch"):
                    if self.eat_p("("):
                        param = self.binding_target()
                        self.expect_p(")")
                    handler = self.block()
                if self.eat_kw("finally"):
                    final = self.block()
                if handler is None and final is None:
                    self.err("try without catch or finally")
                return ("try", block, param, handler, final)
            if v == "switch":
                self.i += 1
                self.expect_p("(")
                disc = self.expression()
                self.expect_p(")")
                self.expect_p("{")
                cases = []
                while not self.eat_p("}"):
                    if self.eat_kw("case"):
                        test = self.expression()
                    else:
                        self.expect_kw("default")
                        test = None
                    self.expect_p(":")
                    body = []
                    while not (self.is_kw("case") or self.is_kw("default") or self.is_p("}")):
                        body.append(self.statement())
                    cases.append((test, body))
                return ("switch", disc, cases)
            if v == "import" and not (self.peek().kind == "punct" and self.peek().val == "."):
                return self.import_decl()
            if v == "export":
                return self.export_decl()
        if t.kind == "id" and self.peek().kind == "punct" and self.peek().val == ":":
            self.err("labelled statements are not supported")
        e = self.expression()
        self.semicolon()
        return ("expr", e, line)

    def block(self):
        self.expect_p("{")
        body = []
        while not self.eat_p("}"):
            body.append(self.statement())
        return ("block", body)

    def var_decl(self):
        kind = self.tok.val
        self.i += 1
        decls = []
        while True:
            target = self.binding_target()
            init = self.assignment() if self.eat_p("=") else None
            decls.append((target, init))
            if not self.eat_p(","):
                break
        return ("var", kind, decls)

    def binding_target(self):
        """identifier | object pattern | array pattern (without default)"""
        if self.is_p("{"):
            self.i += 1
            props, rest = [], None
            while not self.eat_p("}"):
                if self.eat_p("..."):
                    rest = self.ident()
                else:
                    if self.is_p("["):
                        self.err("computed keys in patterns are not supported")
                    key = self.prop_name()
                    if self.eat_p(":"):
                        value = self.binding_element()
                    else:
                        value = ("id", key)
                        if self.eat_p("="):
                            value = ("assignpat", value, self.assignment())
                    props.append((key, value))
                if not self.eat_p(","):
                    self.expect_p("}")
                    break
            return ("objpat", props, rest)
        if self.is_p("["):
            self.i += 1
            elems, rest = [], None
            while not self.eat_p("]"):
                if self.is_p(","):
                    self.i += 1
                    elems.append(None)
                    continue
                if self.eat_p("..."):
                    rest = self.binding_target()
                else:
                    elems.append(self.binding_element())
                if not self.eat_p(","):
                    self.expect_p("]")
                    break
            return ("arrpat", elems, rest)
        return ("id", self.ident())

    def binding_element(self):
        target = self.binding_target()
        if self.eat_p("="):
            return ("assignpat", target, self.assignment())
        return target

    def for_statement(self):
        self.expect_kw("for")
        self.expect_p("(")
        init = None
        if self.is_kw("var") or self.is_kw("let") or self.is_kw("const"):
            kind = self.tok.val
            # for (const x of y) / for (const k in o)
            save = self.i
            self.i += 1
            target = self.binding_target()
            if self.eat_kw("of"):
                it = self.assignment()
                self.expect_p(")")
                return ("forof", kind, target, it, self.statement())
            if self.eat_kw("in"):
                it = self.expression()
                self.expect_p(")")
                return ("forin", kind, target, it, self.statement())
            self.i = save
            self.no_in = True
            init = self.var_decl()
            self.no_in = False
        elif not self.is_p(";"):
            self.no_in = True
            e = self.expression()
            self.no_in = False
            if self.eat_kw("of"):
                it = self.assignment()
                self.expect_p(")")
                return ("forof", None, self.to_pattern(e), it, self.statement())
            if self.eat_kw("in"):
                it = self.expression()
                self.expect_p(")")
                return ("forin", None, self.to_pattern(e), it, self.statement())
            init = ("expr", e, self.tok.line)
        self.expect_p(";")
        test = None if self.is_p(";") else self.expression()
        self.expect_p(";")
        update = None if self.is_p(")") else self.expression()
        self.expect_p(")")
        return ("for", init, test, update, self.statement())

    def class_decl(self):
        line = self.tok.line
        self.expect_kw("class")
        name = self.ident() if self.tok.kind == "id" else None
        return ("classdecl", name, self.class_rest(name), line)

    def class_rest(self, name):
        sup = None
        if self.eat_kw("extends"):
            sup = self.unary()
        self.expect_p("{")
        members = []      # (kind: method|get|set|field, static, key, fn-or-expr)
        while not self.eat_p("}"):
            if self.eat_p(";"):
                continue
            static = False
            if self.is_kw("static") and not (self.peek().kind == "punct" and self.peek().val in ("(", "=")):
                self.i += 1
                static = True
            kind = "method"
            if (self.is_kw("get") or self.is_kw("set")) and not (self.peek().kind == "punct" and self.peek().val in ("(", "=", ";")):
                kind = self.tok.val
                self.i += 1
            key = self.prop_name()
            if self.is_p("("):
                members.append((kind, static, key, self.function_rest(key, False)))
            else:
                init = self.assignment() if self.eat_p("=") else None
                self.semicolon()
                members.append(("field", static, key, init))
        return ("class", name, sup, members)

    def import_decl(self):
        self.expect_kw("import")
        specs = []          # (imported name | 'default' | '*', local name)
        if self.tok.kind == "str":
            src = self.tok.val
            self.i += 1
            self.semicolon()
            return ("import", specs, src)
        if self.tok.kind == "id":
            specs.append(("default", self.ident()))
            self.eat_p(",")
        if self.eat_p("*"):
            self.expect_kw("as")
            specs.append(("*", self.ident()))
        elif self.eat_p("{"):
            while not self.eat_p("}"):
                imported = self.prop_name()
                local = self.ident() if self.eat_kw("as") else imported
                specs.append((imported, local))
                if not self.eat_p(","):
                    self.expect_p("}")
                    break
        self.expect_kw("from")
        src = self.tok.val
        if self.tok.kind != "str":
            self.err("expected module specifier")
        self.i += 1
        self.semicolon()
        return ("import", specs, src)

    def export_decl(self):
        self.expect_kw("export")
        if self.eat_kw("default"):
            if self.is_kw("function"):
                self.i += 1
                name = self.ident() if self.tok.kind == "id" else "default"
                return ("export", "default", ("funcdecl", name, self.function_rest(name, False), self.tok.line))
            if self.is_kw("class"):
                return ("export", "default", self.class_decl())
            e = self.assignment()
            self.semicolon()
            return ("export", "defaultexpr", e)
        if self.eat_p("{"):
            specs = []
            while not self.eat_p("}"):
                local = self.prop_name()
                exported = self.prop_name() if self.eat_kw("as") else local
                specs.append((local, exported))
                if not self.eat_p(","):
                    self.expect_p("}")
                    break
            src = None
            if self.eat_kw("from"):
                src = self.tok.val
                self.i += 1
            self.semicolon()
            return ("export", "specs", specs, src)
        return ("export", "decl", self.statement())

    # ---- functions ----
    def params(self):
        self.expect_p("(")
        out, rest = [], None
        while not self.eat_p(")"):
            if self.eat_p("..."):
                rest = self.binding_target()
            else:
                out.append(self.binding_element())
            if not self.eat_p(","):
                self.expect_p(")")
                break
        return out, rest

    def function_rest(self, name, is_arrow):
        line = self.tok.line
        params, rest = self.params()
        saved, self.no_in = self.no_in, False
        body = self.block()
        self.no_in = saved
        return ("fn", name, params, rest, body, False, False, line)       # (…, isArrow, isExpressionBody, line)

    # ---- expressions ----
    def expression(self):
        e = self.assignment()
        if self.is_p(","):
            items = [e]
            while self.eat_p(","):
                items.append(self.assignment())
            return ("seq", items)
        return e

    def is_arrow_ahead(self):
        """at '(' : does the matching ')' precede '=>' ?"""
        depth, j = 0, self.i
        while True:
            t = self.t[j]
            if t.kind == "eof":
                return False
            if t.kind == "punct":
                if t.val in "([{":
                    depth += 1
                elif t.val in ")]}":
                    depth -= 1
                    if depth == 0:
                        nx = self.t[j + 1]
                        return nx.kind == "punct" and nx.val == "=>"
            j += 1

    def arrow_body(self, params, rest, line):
        saved, self.no_in = self.no_in, False
        if self.is_p("{"):
            body = self.block()
            self.no_in = saved
            return ("fn", None, params, rest, body, True, False, line)
        e = self.assignment()
        self.no_in = saved
        return ("fn", None, params, rest, e, True, True, line)

    def assignment(self):
        t = self.tok
        if t.kind == "id" and self.peek().kind == "punct" and self.peek().val == "=>":
            self.i += 2
            return self.arrow_body([("id", t.val)], None, t.line)
        if t.kind == "punct" and t.val == "(" and self.is_arrow_ahead():
            params, rest = self.params()
            self.expect_p("=>")
            return self.arrow_body(params, rest, t.line)
        left = self.conditional()
        if self.tok.kind == "punct" and self.tok.val in ASSIGN_OPS:
            op = self.tok.val
            self.i += 1
            right = self.assignment()
            if op == "=" and left[0] in ("arr", "obj"):
                left = self.to_pattern(left)
            elif left[0] not in ("id", "member"):
                self.err("invalid assignment target")
            return ("assign", op, left, right)
        return left

    def to_pattern(self, e):
        k = e[0]
        if k in ("id", "member"):
            return e
        if k == "assign" and e[1] == "=":
            return ("assignpat", self.to_pattern(e[2]), e[3])
        if k == "arr":
            elems, rest = [], None
            for x in e[1]:
                if x is not None and x[0] == "spread":
                    rest = self.to_pattern(x[1])
                else:
                    elems.append(None if x is None else self.to_pattern(x))
            return ("arrpat", elems, rest)
        if k == "obj":
            props, rest = [], None
            for p in e[1]:
                if p[0] == "spread":
                    rest = p[1][1]
                elif p[0] == "prop" and p[1][0] == "str":
                    props.append((p[1][1], self.to_pattern(p[2])))
                else:
                    self.err("unsupported destructuring target")
            return ("objpat", props, rest)
        self.err("invalid destructuring target")

    def conditional(self):
        test = self.binary(0)
        if self.eat_p("?"):
            saved, self.no_in = self.no_in, False
            a = self.assignment()
            self.no_in = saved
            self.expect_p(":")
            b = self.assignment()
            return ("cond", test, a, b)
        return test

    def binary(self, min_prec):
        left = self.unary()
        while True:
            t = self.tok
            op = t.val if (t.kind == "punct" or (t.kind == "kw" and t.val in ("instanceof", "in"))) else None
            if op == "in" and self.no_in:
                break
            prec = BIN_PREC.get(op)
            if prec is None or prec <= min_prec:
                break
            self.i += 1
            right = self.binary(prec - 1 if op == "**" else prec)
            left = ("logical", op, left, right) if op in ("&&", "||", "??") else ("bin", op, left, right)
        return left

    def unary(self):
        t = self.tok
        if t.kind == "punct" and t.val in ("!", "-", "+", "~"):
            self.i += 1
            return ("unary", t.val, self.unary())
        if t.kind == "punct" and t.val in ("++", "--"):
            self.i += 1
            return ("update", t.val, True, self.unary())
        if t.kind == "kw" and t.val in ("typeof", "void", "delete"):
            self.i += 1
            return ("unary", t.val, self.unary())
        e = self.postfix()
        return e

    def postfix(self):
        e = self.call_member()
        t = self.tok
        if t.kind == "punct" and t.val in ("++", "--") and not t.nl:
            self.i += 1
            return ("update", t.val, False, e)
        return e

    def arguments(self):
        self.expect_p("(")
        args = []
        while not self.eat_p(")"):
            if self.eat_p("..."):
                args.append(("spread", self.assignment()))
            else:
                args.append(self.assignment())
            if not self.eat_p(","):
                self.expect_p(")")
                break
        return args

    def call_member(self):
        line = self.tok.line
        if self.eat_kw("new"):
            callee = self.member_only()
            args = self.arguments() if self.is_p("(") else []
            e = ("new", callee, args, line)
        else:
            e = self.primary()
        while True:
            t = self.tok
            if t.kind != "punct":
                break
            if t.val == ".":
                self.i += 1
                e = ("member", e, ("str", self.prop_name()), False)
            elif t.val == "?.":
                self.i += 1
                if self.is_p("("):
                    e = ("call", e, self.arguments(), True, t.line)
                elif self.eat_p("["):
                    k = self.expression()
                    self.expect_p("]")
                    e = ("member", e, k, True)
                else:
                    e = ("member", e, ("str", self.prop_name()), True)
            elif t.val == "[":
                self.i += 1
                saved, self.no_in = self.no_in, False
                k = self.expression()
                self.no_in = saved
                self.expect_p("]")
                e = ("member", e, k, False)
            elif t.val == "(":
                e = ("call", e, self.arguments(), False, t.line)
            else:
                break
        return e

    def member_only(self):
        """callee of `new`: primary followed by member accesses (no calls)"""
        if self.eat_kw("new"):
            callee = self.member_only()
            args = self.arguments() if self.is_p("(") else []
            e = ("new", callee, args, self.tok.line)
        else:
            e = self.primary()
        while True:
            if self.eat_p("."):
                e = ("member", e, ("str", self.prop_name()), False)
            elif self.is_p("["):
                self.i += 1
                k = self.expression()
                self.expect_p("]")
                e = ("member", e, k, False)
            else:
                return e

    def primary(self):
        t = self.tok
        k, v = t.kind, t.val
        if k == "num":
            self.i += 1
            return ("num", v)
        if k == "str":
            self.i += 1
            return ("str", v)
        if k == "tpl":
            self.i += 1
            parts = []
            for p in v:
                if isinstance(p, str):
                    parts.append(p)
                else:
                    sub = Parser(p[0], self.fname)
                    for tk in sub.t:
                        tk.line += p[1] - 1
                    parts.append(sub.expression())
                    if sub.tok.kind != "eof":
                        sub.err("unexpected token in template expression")
            return ("tpl", parts)
        if k == "id":
            self.i += 1
            return ("id", v)
        if k == "kw":
            if v == "this":
                self.i += 1
                return ("this",)
            if v == "null":
                self.i += 1
                return ("null",)
            if v == "undefined":
                self.i += 1
                return ("id", "undefined")
            if v == "true":
                self.i += 1
                return ("bool", True)
            if v == "false":
                self.i += 1
                return ("bool", False)
            if v == "function":
                self.i += 1
                name = self.ident() if self.tok.kind == "id" else None
                return self.function_rest(name, False)
            if v == "class":
                self.i += 1
                name = self.ident() if self.tok.kind == "id" else None
                return self.class_rest(name)
            if v == "super":
                self.i += 1
                return ("super",)
            if v == "import" and self.peek().kind == "punct" and self.peek().val == ".":
                self.i += 2
                if self.prop_name() != "meta":
                    self.err("only import.meta is supported")
                return ("importmeta",)
            if v in CONTEXTUAL:
                self.i += 1
                return ("id", v)
        if k == "punct":
            if v == "(":
                self.i += 1
                saved, self.no_in = self.no_in, False
                e = self.expression()
                self.no_in = saved
                self.expect_p(")")
                return ("paren", e)
            if v == "[":
                self.i += 1
                saved, self.no_in = self.no_in, False
                elems = []
                while not self.eat_p("]"):
                    if self.is_p(","):
                        self.i += 1
                        elems.append(None)
                        continue
                    if self.eat_p("..."):
                        elems.append(("spread", self.assignment()))
                    else:
                        elems.append(self.assignment())
                    if not self.eat_p(","):
                        self.expect_p("]")
                        break
                self.no_in = saved
                return ("arr", elems)
            if v == "{":
                return self.object_literal()
        self.err("unexpected token")

    def object_literal(self):
        self.expect_p("{")
        saved, self.no_in = self.no_in, False
        props = []          # ('prop', keyExpr, valueExpr) | ('spread', expr) | ('get'|'set', keyExpr, fn)
        while not self.eat_p("}"):
            if self.eat_p("..."):
                props.append(("spread", self.assignment()))
            else:
                t = self.tok
                nxt = self.peek()
                if t.kind == "kw" and t.val in ("get", "set") and not (nxt.kind == "punct" and nxt.val in (",", ":", "(", "}")):
                    self.i += 1
                    key = ("str", self.prop_name())
                    props.append((t.val, key, self.function_rest(key[1], False)))
                else:
                    if self.eat_p("["):
                        key = self.assignment()
                        self.expect_p("]")
                        shorthand_ok = False
                    else:
                        shorthand_ok = t.kind == "id" or (t.kind == "kw" and t.val in CONTEXTUAL)
                        key = ("str", self.prop_name())
                    if self.eat_p(":"):
                        props.append(("prop", key, self.assignment()))
                    elif self.is_p("("):
                        props.append(("prop", key, self.function_rest(key[1] if key[0] == "str" else None, False)))
                    elif shorthand_ok:
                        if self.eat_p("="):       # only valid when the literal is reinterpreted as a pattern
                            props.append(("prop", key, ("assign", "=", ("id", key[1]), self.assignment())))
                        else:
                            props.append(("prop", key, ("id", key[1])))
                    else:
                        self.err("bad object literal")
            if not self.eat_p(","):
                self.expect_p("}")
                break
        self.no_in = saved
        return ("obj", props)


def parse(src: str, fname: str = "<js>"):
    return Parser(src, fname).parse_program()
