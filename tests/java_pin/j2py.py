#!/usr/bin/env python
"""tests/java_pin/j2py.py -- a mechanical Java -> Python transliterator for the subset of Java that
lib/src/kmergutsjava/KmerGutsJava.java is written in, plus the small runtime (java.util / java.io stand-ins) the
transliterated code runs on.

Why: the image has no JVM, so the reference cannot be executed here and the CPU oracle (oracle/kg_oracle.c) was checked
against the Java only by reading it.  This tool removes the human from that step: it parses the UNMODIFIED Java source
(tokenizer + recursive-descent parser for statements and expressions), and emits Python with the same control flow, the
same names and the same evaluation order, statement by statement.  Nothing about k-mers, hits or calls is known to this
file -- it knows Java syntax and the semantics of the Java types the source uses:

  * int / long / float / byte arithmetic: results are reduced to the declared type where the source assigns or returns an
    arithmetic expression (two's complement wrap for int and long, IEEE round-to-nearest to binary32 for float: a double
    sum of two floats rounded to float IS the float sum); integer `/` truncates, `%` takes the sign of the dividend;
  * x++ / x-- inside expressions (walrus), `for` with `continue` (the update still runs), switch with fall-through,
    try / catch / finally, anonymous classes (hoisted to local classes; captured locals are Python closures), labelled
    nothing (the source has none; the parser refuses what it does not know instead of guessing);
  * String.format (%d %s %c %f %N.Mf with Java's HALF_UP rounding of the shortest repr), String.trim (<= U+0020),
    BufferedReader.readLine (\\n, \\r, \\r\\n), StringTokenizer, ArrayList (index-checked), HashMap / LinkedHashMap
    (insertion order, equals()-based keys), Collections.sort (stable, comparator-driven), InputStream.read / skip,
    GZIPInputStream, PrintWriter.

Not modelled (stated so nobody reads more into a green run than it proves): int overflow INSIDE a comparison operand
(e.g. `a + b < c` with a + b beyond 2^31; the results of assignments and returns are wrapped), HashMap iteration order
(the source never iterates a HashMap), the default charset (files are read as ISO-8859-1, one char per byte),
ConcurrentModificationException, and pre-JDK-19 FloatingDecimal's rare non-shortest digit strings.

Used by tests/java_pin/transliterated_pin.py (the driver) and tests/test_java_transliteration.py.  Test infrastructure
only: nothing in the product imports it.
"""
import re

# =====================================================================================================================
# tokenizer
# =====================================================================================================================
_TOKEN = re.compile(r"""
   (?P<ws>\s+) | (?P<lc>//[^\n]*) | (?P<bc>/\*.*?\*/)
 | (?P<str>"(?:\\.|[^"\\])*") | (?P<chr>'(?:\\.|[^'\\])+')
 | (?P<num>0[xX][0-9a-fA-F]+[lL]? | \d+\.\d*(?:[eE][+-]?\d+)?[fFdD]? | \.\d+(?:[eE][+-]?\d+)?[fFdD]?
          | \d+[eE][+-]?\d+[fFdD]? | \d+[fFdDlL]?)
 | (?P<id>[A-Za-z_$][A-Za-z_$0-9]*)
 | (?P<op><<= | \+\+ | -- | && | \|\| | == | != | <= | >= | \+= | -= | \*= | /= | %= | &= | \|= | \^= | << | ->
          | [-+*/%&|^!~<>=?:;,.(){}\[\]@])
""", re.X | re.S)

KEYWORDS = {"abstract", "boolean", "break", "byte", "case", "catch", "char", "class", "continue", "default", "do", "double",
            "else", "extends", "final", "finally", "float", "for", "if", "implements", "import", "instanceof", "int",
            "interface", "long", "new", "package", "private", "protected", "public", "return", "short", "static", "switch",
            "this", "throw", "throws", "try", "void", "while", "true", "false", "null", "synchronized", "transient", "volatile"}
PRIMITIVES = {"int", "long", "float", "double", "byte", "char", "short", "boolean"}
MODIFIERS = {"public", "private", "protected", "static", "final", "abstract", "synchronized", "transient", "volatile"}


class Tok:
    __slots__ = ("kind", "text", "pos", "end")

    def __init__(self, kind, text, pos, end):
        self.kind, self.text, self.pos, self.end = kind, text, pos, end

    def __repr__(self):
        return f"{self.kind}:{self.text!r}@{self.pos}"


def tokenize(src):
    out, i = [], 0
    while i < len(src):
        m = _TOKEN.match(src, i)
        if not m:
            raise SyntaxError(f"cannot tokenize at offset {i}: {src[i:i + 40]!r}")
        k = m.lastgroup
        if k not in ("ws", "lc", "bc"):
            out.append(Tok(k, m.group(), m.start(), m.end()))
        i = m.end()
    out.append(Tok("eof", "", len(src), len(src)))
    return out


# =====================================================================================================================
# parser: Java subset -> tuples
# =====================================================================================================================
class ParseError(Exception):
    pass


class Parser:
    def __init__(self, src, lenient=False):
        """lenient: SYNTAX CHECK ONLY -- also accept constructors, records, try-with-resources, lambdas, X.class and
        annotation arguments (the repo's own Java files use them); the emitter has no rule for those nodes and refuses them."""
        self.src = src
        self.t = tokenize(src)
        self.i = 0
        self.lenient = lenient

    # ---- token helpers ----
    def peek(self, k=0):
        return self.t[min(self.i + k, len(self.t) - 1)]

    def at(self, text, k=0):
        tk = self.peek(k)
        return tk.text == text and tk.kind in ("op", "id")

    def accept(self, text):
        if self.at(text):
            self.i += 1
            return True
        return False

    def expect(self, text):
        if not self.accept(text):
            tk = self.peek()
            line = self.src.count("\n", 0, tk.pos) + 1
            raise ParseError(f"line {line}: expected {text!r}, found {tk.text!r}")

    def ident(self):
        tk = self.peek()
        if tk.kind != "id" or tk.text in KEYWORDS:
            line = self.src.count("\n", 0, tk.pos) + 1
            raise ParseError(f"line {line}: expected an identifier, found {tk.text!r}")
        self.i += 1
        return tk.text

    def line(self):
        return self.src.count("\n", 0, self.peek().pos) + 1

    # ---- types ----
    def try_type(self):
        """Type at the cursor -> its text without generic arguments (e.g. 'List', 'byte[]'), or None (cursor restored)."""
        save = self.i
        tk = self.peek()
        if tk.kind != "id" or (tk.text in KEYWORDS and tk.text not in PRIMITIVES):
            return None
        name = tk.text
        self.i += 1
        if name not in PRIMITIVES:
            while self.at(".") and self.peek(1).kind == "id" and self.peek(1).text not in KEYWORDS:
                self.i += 1
                name = self.peek().text  # keep the last component: imports make simple names unambiguous here
                self.i += 1
            if self.at("<"):
                if not self._skip_generic_args():
                    self.i = save
                    return None
        while self.at("[") and self.at("]", 1):
            self.i += 2
            name += "[]"
        return name

    def _skip_generic_args(self):
        depth = 0
        while True:
            tk = self.peek()
            if tk.text == "<":
                depth += 1
            elif tk.text == ">":
                depth -= 1
            elif tk.kind == "id" or tk.text in (",", ".", "?", "[", "]"):
                pass
            else:
                return False
            self.i += 1
            if depth == 0:
                return True

    # ---- compilation unit / classes ----
    def compilation_unit(self):
        classes = []
        while self.peek().kind != "eof":
            if self.accept("package") or self.accept("import"):
                while not self.accept(";"):
                    self.i += 1
                continue
            classes.append(self.type_decl())
        return classes

    def modifiers(self):
        mods = set()
        while True:
            if self.at("@"):
                self.i += 1
                self.ident()
                while self.at(".") and self.peek(1).kind == "id":
                    self.i += 2
                if self.lenient and self.at("("):
                    self.args()
                continue
            if self.peek().text in MODIFIERS:
                mods.add(self.peek().text)
                self.i += 1
                continue
            return mods

    def type_decl(self):
        mods = self.modifiers()
        if self.accept("interface"):
            kind = "interface"
        elif self.lenient and self.at("record") and self.peek(1).kind == "id" and self.peek(2).text == "(":
            self.i += 1
            kind = "record"
        else:
            self.expect("class")
            kind = "class"
        name = self.ident()
        if self.at("<"):
            self._skip_generic_args()
        if kind == "record":
            self.params()
        while self.at("extends") or self.at("implements"):
            self.i += 1
            self.try_type()
            while self.accept(","):
                self.try_type()
        body = self.class_body(name, kind == "interface")
        return ("class", name, kind, mods, body)

    def class_body(self, cname, is_iface=False):
        """-> list of members: ('field', mods, type, name, init) | ('method', mods, rtype, name, params, body|None) | class"""
        self.expect("{")
        members = []
        while not self.accept("}"):
            if self.accept(";"):
                continue
            save = self.i
            mods = self.modifiers()
            if self.at("class") or self.at("interface") or \
                    (self.lenient and self.at("record") and self.peek(1).kind == "id" and self.peek(2).text == "("):
                self.i = save
                members.append(self.type_decl())
                continue
            if self.lenient and self.peek().text == cname and self.peek(1).text == "(":   # constructor
                self.i += 1
                params = self.params()
                if self.accept("throws"):
                    self.try_type()
                    while self.accept(","):
                        self.try_type()
                members.append(("ctor", mods, params, self.block()))
                continue
            if self.at("<"):
                raise ParseError(f"line {self.line()}: generic methods are not supported")
            if self.at("void"):
                self.i += 1
                rtype = "void"
            else:
                rtype = self.try_type()
                if rtype is None:
                    raise ParseError(f"line {self.line()}: member declaration expected, found {self.peek().text!r}")
            if self.at("("):  # constructor
                raise ParseError(f"line {self.line()}: constructors are not supported")
            name = self.ident()
            if self.at("("):
                params = self.params()
                if self.accept("throws"):
                    self.try_type()
                    while self.accept(","):
                        self.try_type()
                body = None
                if not self.accept(";"):
                    body = self.block()
                if is_iface:
                    mods = mods | {"abstract"}
                members.append(("method", mods, rtype, name, params, body))
            else:
                while True:
                    ftype = rtype
                    while self.at("[") and self.at("]", 1):
                        self.i += 2
                        ftype += "[]"
                    init = None
                    if self.accept("="):
                        init = self.var_init()
                    members.append(("field", mods, ftype, name, init))
                    if self.accept(","):
                        name = self.ident()
                        continue
                    self.expect(";")
                    break
        return members

    def params(self):
        self.expect("(")
        out = []
        while not self.accept(")"):
            self.modifiers()
            ty = self.try_type()
            if ty is None:
                raise ParseError(f"line {self.line()}: parameter type expected")
            name = self.ident()
            while self.at("[") and self.at("]", 1):
                self.i += 2
                ty += "[]"
            out.append((ty, name))
            self.accept(",")
        return out

    def var_init(self):
        if self.at("{"):
            return self.array_init()
        return self.expr()

    def array_init(self):
        self.expect("{")
        elems = []
        while not self.accept("}"):
            elems.append(self.var_init())
            self.accept(",")
        return ("arrinit", elems)

    # ---- statements ----
    def block(self):
        self.expect("{")
        stmts = []
        while not self.accept("}"):
            stmts.append(self.statement())
        return ("block", stmts)

    def try_local_decl(self):
        """Local variable declaration at the cursor (without the trailing ';' / ':') or None."""
        save = self.i
        self.modifiers()
        ty = self.try_type()
        if ty is None or self.peek().kind != "id" or self.peek().text in KEYWORDS:
            self.i = save
            return None
        if self.peek(1).text not in ("=", ";", ",", ":", "["):
            self.i = save
            return None
        decls = []
        while True:
            name = self.ident()
            vty = ty
            while self.at("[") and self.at("]", 1):
                self.i += 2
                vty += "[]"
            init = None
            if self.accept("="):
                init = self.var_init()
            decls.append((vty, name, init))
            if not self.accept(","):
                break
        return ("local", decls)

    def statement(self):
        ln = self.line()
        tk = self.peek()
        if self.at("{"):
            return self.block()
        if self.accept(";"):
            return ("empty",)
        if tk.kind == "id" and self.peek(1).text == ":" and tk.text not in KEYWORDS:
            raise ParseError(f"line {ln}: labelled statements are not supported")
        if self.accept("if"):
            self.expect("(")
            c = self.expr()
            self.expect(")")
            a = self.statement()
            b = self.statement() if self.accept("else") else None
            return ("if", c, a, b)
        if self.accept("while"):
            self.expect("(")
            c = self.expr()
            self.expect(")")
            return ("while", c, self.statement())
        if self.accept("do"):
            body = self.statement()
            self.expect("while")
            self.expect("(")
            c = self.expr()
            self.expect(")")
            self.expect(";")
            return ("dowhile", body, c)
        if self.accept("for"):
            self.expect("(")
            d = self.try_local_decl()
            if d is not None and self.accept(":"):
                (ty, name, init), = d[1]
                it = self.expr()
                self.expect(")")
                return ("foreach", ty, name, it, self.statement())
            init = []
            if d is not None:
                init.append(d)
            elif not self.at(";"):
                init.append(("expr", self.expr()))
                while self.accept(","):
                    init.append(("expr", self.expr()))
            self.expect(";")
            cond = None if self.at(";") else self.expr()
            self.expect(";")
            upd = []
            if not self.at(")"):
                upd.append(self.expr())
                while self.accept(","):
                    upd.append(self.expr())
            self.expect(")")
            return ("for", init, cond, upd, self.statement())
        if self.accept("return"):
            e = None if self.at(";") else self.expr()
            self.expect(";")
            return ("return", e)
        if self.accept("break"):
            self.expect(";")
            return ("break",)
        if self.accept("continue"):
            self.expect(";")
            return ("continue",)
        if self.accept("throw"):
            e = self.expr()
            self.expect(";")
            return ("throw", e)
        if self.accept("try"):
            resources = []
            if self.at("("):
                if not self.lenient:
                    raise ParseError(f"line {ln}: try-with-resources is not supported")
                self.i += 1
                while not self.accept(")"):
                    d = self.try_local_decl()
                    resources.append(d if d is not None else ("expr", self.expr()))
                    self.accept(";")
            body = self.block()
            if resources:
                body = ("unsupported", "try-with-resources", resources, body)
            catches, fin = [], None
            while self.accept("catch"):
                self.expect("(")
                self.modifiers()
                ty = self.try_type()
                if self.at("|"):
                    raise ParseError(f"line {ln}: multi-catch is not supported")
                name = self.ident()
                self.expect(")")
                catches.append((ty, name, self.block()))
            if self.accept("finally"):
                fin = self.block()
            return ("try", body, catches, fin)
        if self.accept("switch"):
            self.expect("(")
            e = self.expr()
            self.expect(")")
            self.expect("{")
            groups = []  # (labels, stmts); a label is an expression or None for default
            while not self.accept("}"):
                labels = []
                while self.at("case") or self.at("default"):
                    if self.accept("default"):
                        labels.append(None)
                    else:
                        self.expect("case")
                        labels.append(self.expr())
                    self.expect(":")
                stmts = []
                while not (self.at("case") or self.at("default") or self.at("}")):
                    stmts.append(self.statement())
                groups.append((labels, stmts))
            return ("switch", e, groups)
        if tk.text in ("synchronized", "assert"):
            raise ParseError(f"line {ln}: {tk.text} is not supported")
        d = self.try_local_decl()
        if d is not None:
            self.expect(";")
            return d
        e = self.expr()
        self.expect(";")
        return ("expr", e)

    # ---- expressions (precedence climbing) ----
    ASSIGN_OPS = {"=", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "<<="}
    BIN = [["||"], ["&&"], ["|"], ["^"], ["&"], ["==", "!="], ["<", ">", "<=", ">=", "instanceof"], ["<<", ">>", ">>>"],
           ["+", "-"], ["*", "/", "%"]]

    def expr(self):
        left = self.ternary()
        if self.peek().text in self.ASSIGN_OPS and self.peek().kind == "op":
            op = self.peek().text
            self.i += 1
            right = self.expr()
            return ("assign", op, left, right)
        return left

    def ternary(self):
        c = self.binary(0)
        if self.accept("?"):
            a = self.expr()
            self.expect(":")
            b = self.ternary()
            return ("cond", c, a, b)
        return c

    def _shift_op(self):
        """'>' '>' written without a gap is a shift (the tokenizer keeps '>' single for the sake of generic types)."""
        a, b, c = self.peek(), self.peek(1), self.peek(2)
        if a.text == ">" and b.text == ">" and b.pos == a.end:
            if c.text == ">" and c.pos == b.end:
                return ">>>", 3
            return ">>", 2
        return None, 0

    def binary(self, level):
        if level == len(self.BIN):
            return self.unary()
        left = self.binary(level + 1)
        while True:
            tk = self.peek()
            op, n = None, 1
            if level == 7:
                op, n = self._shift_op()
                if op is None and tk.text == "<<" and tk.kind == "op":
                    op, n = "<<", 1
            elif level == 6 and tk.text == ">" and self._shift_op()[0]:
                op = None
            elif tk.text in self.BIN[level] and tk.kind in ("op", "id"):
                op = tk.text
            if op is None:
                return left
            self.i += n
            if op == "instanceof":
                left = ("instanceof", left, self.try_type())
                continue
            right = self.binary(level + 1)
            left = ("binary", op, left, right)

    def unary(self):
        tk = self.peek()
        if tk.kind == "op" and tk.text in ("+", "-", "!", "~"):
            self.i += 1
            return ("unary", tk.text, self.unary())
        if tk.kind == "op" and tk.text in ("++", "--"):
            self.i += 1
            return ("prefix", tk.text, self.unary())
        if tk.text == "(":
            save = self.i
            self.i += 1
            ty = self.try_type()
            if ty is not None and self.accept(")"):
                nx = self.peek()
                base = ty.rstrip("[]")
                starts_operand = nx.kind in ("id", "num", "str", "chr") and nx.text not in ("instanceof",) or nx.text in ("(", "!", "~")
                if base in PRIMITIVES and (starts_operand or nx.text in ("-", "+")):
                    return ("cast", ty, self.unary())
                if base not in PRIMITIVES and base[:1].isupper() and starts_operand:
                    return ("cast", ty, self.unary())
            self.i = save
        return self.postfix(self.primary())

    def args(self):
        self.expect("(")
        out = []
        while not self.accept(")"):
            out.append(self.expr())
            self.accept(",")
        return out

    def primary(self):
        tk = self.peek()
        if tk.kind == "num":
            self.i += 1
            return ("num", tk.text)
        if tk.kind == "str":
            self.i += 1
            return ("str", tk.text)
        if tk.kind == "chr":
            self.i += 1
            return ("chr", tk.text)
        if self.lenient and tk.text == "(":   # ( params ) -> body
            j, depth = self.i, 0
            while True:
                t = self.t[j].text
                depth += (t == "(") - (t == ")")
                j += 1
                if depth == 0 or self.t[j].kind == "eof":
                    break
            if self.t[j].text == "->":
                self.i = j + 1
                return ("unsupported", "lambda", self.block() if self.at("{") else self.expr())
        if self.lenient and tk.kind == "id" and tk.text not in KEYWORDS and self.peek(1).text == "->":
            self.i += 2
            return ("unsupported", "lambda", self.block() if self.at("{") else self.expr())
        if self.lenient and tk.text in PRIMITIVES and self.peek(1).text == "." and self.peek(2).text == "class":
            self.i += 3
            return ("unsupported", "class literal")
        if tk.text == "(":
            self.i += 1
            e = self.expr()
            self.expect(")")
            return ("paren", e)
        if self.accept("true"):
            return ("bool", True)
        if self.accept("false"):
            return ("bool", False)
        if self.accept("null"):
            return ("null",)
        if self.accept("this"):
            return ("this",)
        if self.accept("new"):
            ty = self.try_type_no_dims()
            if self.at("["):
                dims = []
                while self.at("["):
                    self.i += 1
                    if self.accept("]"):
                        dims.append(None)
                    else:
                        dims.append(self.expr())
                        self.expect("]")
                init = self.array_init() if self.at("{") else None
                return ("newarr", ty, dims, init)
            a = self.args()
            body = None
            if self.at("{"):
                body = self.class_body(ty)
            return ("new", ty, a, body)
        if tk.kind == "id" and tk.text not in KEYWORDS:
            self.i += 1
            if self.at("("):
                return ("call", None, tk.text, self.args())
            return ("name", tk.text)
        raise ParseError(f"line {self.line()}: unexpected {tk.text!r} in an expression")

    def try_type_no_dims(self):
        tk = self.peek()
        if tk.kind != "id":
            raise ParseError(f"line {self.line()}: type expected after new")
        name = tk.text
        self.i += 1
        if name not in PRIMITIVES:
            while self.at(".") and self.peek(1).kind == "id":
                self.i += 1
                name = self.peek().text
                self.i += 1
            if self.at("<"):
                self._skip_generic_args()
        return name

    def postfix(self, e):
        while True:
            if self.lenient and self.at(".") and self.peek(1).text == "class":
                self.i += 2
                e = ("unsupported", "class literal")
            elif self.at("."):
                self.i += 1
                name = self.ident()
                if self.at("("):
                    e = ("call", e, name, self.args())
                else:
                    e = ("field", e, name)
            elif self.at("["):
                self.i += 1
                idx = self.expr()
                self.expect("]")
                e = ("index", e, idx)
            elif self.peek().kind == "op" and self.peek().text in ("++", "--"):
                op = self.peek().text
                self.i += 1
                e = ("postfix", op, e)
            else:
                return e


# =====================================================================================================================
# emitter: tuples -> Python source
# =====================================================================================================================
PY_RESERVED = {"and", "as", "assert", "async", "await", "def", "del", "elif", "except", "exec", "from", "global", "import", "in", "is",
               "lambda", "nonlocal", "not", "or", "pass", "print", "raise", "with", "yield", "None", "True", "False",
               "len", "list", "str", "int", "float", "id", "max", "min", "abs", "type", "object", "range", "next", "iter", "input",
               "map", "filter", "format", "bytes", "chr", "ord", "hash", "set", "dict", "tuple", "bool", "sum", "all", "any", "open",
               "file", "dir", "vars", "repr", "round", "sorted", "super", "zip", "re", "os", "sys", "math"}
INTEGRAL = {"int", "long", "byte", "short", "char"}
STRING_METHODS = {"trim", "charAt", "substring", "startsWith", "endsWith", "indexOf", "toCharArray", "equals", "hashCode",
                  "isEmpty", "split", "toString", "contains", "lastIndexOf", "toUpperCase", "toLowerCase"}
STRING_RETURNING = {"trim", "substring", "toString", "getMessage", "getName", "getCanonicalPath", "getAbsolutePath", "nextToken",
                    "readLine", "format", "getProperty", "getPath"}
ESCAPES = {"n": "\n", "t": "\t", "r": "\r", "0": "\0", "'": "'", '"': '"', "\\": "\\", "b": "\b", "f": "\f"}
WRAP = {"int": "_i32", "long": "_i64", "float": "_f32", "byte": "_i8", "short": "_i16"}
DEFAULTS = {"int": "0", "long": "0", "short": "0", "byte": "0", "float": "0.0", "double": "0.0", "boolean": "False", "char": "'\\0'"}


def py_name(n):
    return n + "_" if n in PY_RESERVED else n


def java_char(text):
    body = text[1:-1]
    if body[0] == "\\":
        if body[1] == "u":
            return chr(int(body[2:], 16))
        return ESCAPES[body[1]]
    return body


def java_string(text):
    out, i, body = [], 0, text[1:-1]
    while i < len(body):
        if body[i] == "\\":
            if body[i + 1] == "u":
                out.append(chr(int(body[i + 2:i + 6], 16)))
                i += 6
            else:
                out.append(ESCAPES[body[i + 1]])
                i += 2
        else:
            out.append(body[i])
            i += 1
    return "".join(out)


class ClassInfo:
    def __init__(self, name, members, outer=None, anon=False):
        self.name, self.outer, self.anon = name, outer, anon
        self.fields = {}    # name -> (type, is_static)
        self.methods = {}   # name -> (rtype, is_static)
        for m in members:
            if m[0] == "field":
                self.fields[m[3]] = (m[2], "static" in m[1])
            elif m[0] == "method":
                if m[3] in self.methods:
                    raise ParseError(f"overloaded method {name}.{m[3]} is not supported")
                self.methods[m[3]] = (m[2], "static" in m[1])


class Emitter:
    def __init__(self, classes):
        self.classes = {}        # simple name -> ClassInfo (top-level and nested, flattened: the source's names are unique)
        self.field_types = {}    # field name -> type, over all named classes (for `expr.field` whose receiver type is unknown)
        self.anon_count = 0
        self.top = list(classes)
        for c in classes:
            self._collect(c, None)

    def _collect(self, c, outer):
        _, name, kind, mods, body = c
        info = ClassInfo(name, body, outer)
        self.classes[name] = info
        for fname, (ftype, _) in info.fields.items():
            self.field_types.setdefault(fname, ftype)
        for m in body:
            if m[0] == "class":
                self._collect(m, info)

    def emit_module(self):
        out = ["# GENERATED by tests/java_pin/j2py.py from the reference's Java source -- do not edit, do not commit",
               "from j2py_runtime import *  # noqa: F401,F403", ""]
        statics = []
        for c in self.top:
            self.emit_class(c, 0, out, statics, [], None)
        out.append("")
        out.extend(statics)
        return "\n".join(out) + "\n"

    def emit_class(self, c, ind, out, statics, ctx, outer_fn, anon_name=None):
        """ctx: enclosing (ClassInfo, this-name or None) pairs, innermost last; outer_fn: the FunctionEmitter whose locals an
        anonymous class captures."""
        _, name, kind, mods, body = c
        pad = "    " * ind
        cname = anon_name or name
        info = self.classes[name] if anon_name is None else ClassInfo(cname, body, None, True)
        this = f"this{sum(1 for _, t in ctx if t)}"
        fields = [m for m in body if m[0] == "field"]
        methods = [m for m in body if m[0] == "method"]
        if anon_name is None:
            for n in (m for m in body if m[0] == "class"):  # static nested classes and interfaces: module level
                self.emit_class(n, ind, out, statics, [], None)
        elif any(m[0] == "class" for m in body):
            raise ParseError("classes nested in an anonymous class are not supported")
        w = out.append
        w(f"{pad}class {cname}(JObject):")
        inst_fields = [f for f in fields if "static" not in f[1]]
        if not methods and anon_name is None and inst_fields:
            w(f"{pad}    __slots__ = ({', '.join(repr(py_name(f[3])) for f in inst_fields)},)")
        my_ctx = ctx + [(info, this)]
        static_ctx = ctx + [(info, None)]
        wrote = False
        if inst_fields:
            w(f"{pad}    def __init__({this}):")
            for f in inst_fields:
                fn = FunctionEmitter(self, my_ctx, None, [], ind + 2, outer_fn)
                val = DEFAULTS.get(f[2], "None") if f[4] is None else fn.init_value(f[2], f[4])
                out.extend(fn.flush_pre())
                w(f"{pad}        {this}.{py_name(f[3])} = {val}")
            w("")
            wrote = True
        for f in fields:
            if "static" in f[1]:
                if statics is None:
                    raise ParseError("static fields in an anonymous class are not supported")
                fn = FunctionEmitter(self, static_ctx, None, [], 0, None)
                val = DEFAULTS.get(f[2], "None") if f[4] is None else fn.init_value(f[2], f[4])
                statics.extend(fn.flush_pre())
                statics.append(f"{cname}.{py_name(f[3])} = {val}")
        for m in methods:
            _, mmods, rtype, mname, params, mbody = m
            if mbody is None:
                continue
            wrote = True
            plist = [py_name(p[1]) for p in params]
            if "static" in mmods:
                w(f"{pad}    @staticmethod")
                w(f"{pad}    def {py_name(mname)}({', '.join(plist)}):")
                fn = FunctionEmitter(self, static_ctx, rtype, params, ind + 2, outer_fn)
            else:
                w(f"{pad}    def {py_name(mname)}({', '.join([this] + plist)}):")
                fn = FunctionEmitter(self, my_ctx, rtype, params, ind + 2, outer_fn)
            out.extend(fn.body(mbody) or [f"{pad}        pass"])
            w("")
        if not wrote:
            w(f"{pad}    pass")
            w("")


class FunctionEmitter:
    """Emits one method body (or one initializer).  Scopes hold the declared types of locals; `pre` collects the local
    classes that the anonymous-class expressions of the current statement were hoisted into.  Name lookup follows Java:
    own locals, fields of the own class, then (anonymous classes) the enclosing method's locals, the enclosing class ..."""

    def __init__(self, em, ctx, rtype, params, ind, outer):
        self.em, self.ctx, self.rtype, self.ind, self.outer = em, ctx, rtype, ind, outer
        self.scopes = [{p[1]: p[0] for p in params}]
        self.pre = []
        self.tmp = 0
        self.in_switch = 0

    # ---- scope / resolution ----
    def own_local(self, name):
        for s in reversed(self.scopes):
            if name in s:
                return s[name]
        return None

    def lookup_local(self, name):
        """type of `name` if it is a local of this function or a captured local of an enclosing one, else None"""
        fn, level = self, len(self.ctx) - 1
        while fn is not None:
            ty = fn.own_local(name)
            if ty is not None:
                return ty
            if level >= 0 and name in fn.ctx[level][0].fields:
                return None  # a field shadows the enclosing method's locals
            fn, level = fn.outer, level - 1
        return None

    def declare(self, name, ty):
        self.scopes[-1][name] = ty

    def resolve_name(self, name):
        """-> (python code, type)"""
        fn, level = self, len(self.ctx) - 1
        while level >= 0:
            if fn is not None:
                ty = fn.own_local(name)
                if ty is not None:
                    return py_name(name), ty
            info, this = self.ctx[level]
            if name in info.fields:
                fty, static = info.fields[name]
                if static:
                    return f"{info.name}.{py_name(name)}", fty
                if this is None:
                    raise ParseError(f"instance field {name} used from a static context")
                return f"{this}.{py_name(name)}", fty
            fn, level = (fn.outer if fn is not None else None), level - 1
        for k in self.em.classes.values():  # statics of the enclosing top-level class seen from a nested static class
            if name in k.fields and k.fields[name][1]:
                return f"{k.name}.{py_name(name)}", k.fields[name][0]
        return py_name(name), None  # a class name (Integer, Math, KmerGutsJava, Hit ...) or something the runtime provides

    def resolve_method(self, name):
        """-> (python callee, rtype) for an unqualified call"""
        for info, this in reversed(self.ctx):
            if name in info.methods:
                rty, static = info.methods[name]
                if static:
                    return f"{info.name}.{py_name(name)}", rty
                if this is None:
                    raise ParseError(f"instance method {name} called from a static context")
                return f"{this}.{py_name(name)}", rty
        for k in self.em.classes.values():
            if name in k.methods and k.methods[name][1]:
                return f"{k.name}.{py_name(name)}", k.methods[name][0]
        for info, this in reversed(self.ctx):   # inherited from Object (getClass ...)
            if this:
                return f"{this}.{py_name(name)}", None
        return py_name(name), None

    def this_name(self):
        info, this = self.ctx[-1]
        if not this:
            raise ParseError("`this` in a static context")
        return this

    def flush_pre(self):
        p, self.pre = self.pre, []
        return p

    def pad(self, extra=0):
        return "    " * (self.ind + extra)

    # ---- expressions ----
    @staticmethod
    def arithmetic(e):
        """Does evaluating e involve an operation whose Java result can differ from the exact mathematical one?"""
        k = e[0]
        if k == "paren":
            return FunctionEmitter.arithmetic(e[1])
        if k == "binary":
            return e[1] in ("+", "-", "*", "/", "<<") or FunctionEmitter.arithmetic(e[2]) or FunctionEmitter.arithmetic(e[3])
        if k == "unary":
            return e[1] in ("-", "~") or FunctionEmitter.arithmetic(e[2])
        if k == "cast":
            return True
        if k == "cond":
            return FunctionEmitter.arithmetic(e[2]) or FunctionEmitter.arithmetic(e[3])
        return False

    def wrap(self, code, target_type, e, force=False):
        f = WRAP.get(target_type)
        if f and (force or self.arithmetic(e)):
            return f"{f}({code})"
        return code

    def init_value(self, ty, init):
        if init[0] == "arrinit":
            return self.array_literal(ty, init)
        code, ety = self.expr(init)
        if ty == "String" or ety == "String":
            return code
        return self.wrap(code, ty, init)

    def array_literal(self, ty, init):
        elem = ty[:-2] if ty.endswith("[]") else ty
        return "[" + ", ".join(self.init_value(elem, x) for x in init[1]) + "]"

    def expr(self, e, stmt=False):
        """-> (code, java type or None).  stmt=True: the value is discarded (x++ as a statement)."""
        k = e[0]
        if k == "num":
            t = e[1]
            if t.lower().startswith("0x"):
                return (t[:-1], "long") if t[-1] in "lL" else (t, "int")
            if t[-1] in "lL":
                return t[:-1], "long"
            if t[-1] in "fF":
                return f"_f32({t[:-1]})", "float"
            if t[-1] in "dD":
                return repr(float(t[:-1])), "double"
            if "." in t or "e" in t.lower():
                return repr(float(t)), "double"
            return t, "int"
        if k == "str":
            return repr(java_string(e[1])), "String"
        if k == "chr":
            return repr(java_char(e[1])), "char"
        if k == "bool":
            return ("True" if e[1] else "False"), "boolean"
        if k == "null":
            return "None", None
        if k == "this":
            return self.this_name(), self.ctx[-1][0].name
        if k == "paren":
            c, t = self.expr(e[1])
            return f"({c})", t
        if k == "name":
            return self.resolve_name(e[1])
        if k == "field":
            if e[1][0] == "name" and self.lookup_local(e[1][1]) is None and e[1][1] in self.em.classes and \
                    e[2] in self.em.classes[e[1][1]].fields and self.resolve_name(e[1][1])[1] is None:   # Class.staticField
                return f"{e[1][1]}.{py_name(e[2])}", self.em.classes[e[1][1]].fields[e[2]][0]
            c, t = self.expr(e[1])
            if e[2] == "length" and (t is None or t.endswith("[]")):
                return f"len({c})", "int"
            if t in self.em.classes and e[2] in self.em.classes[t].fields:
                fty = self.em.classes[t].fields[e[2]][0]
            else:
                fty = self.em.field_types.get(e[2])
            return f"{c}.{py_name(e[2])}", fty
        if k == "index":
            a, at = self.expr(e[1])
            i, _ = self.expr(e[2])
            return f"{a}[{i}]", (at[:-2] if at and at.endswith("[]") else None)
        if k == "call":
            return self.call(e)
        if k == "new":
            return self.new(e)
        if k == "newarr":
            ty, dims, init = e[1], e[2], e[3]
            if init is not None:
                return self.array_literal(ty + "[]", init), ty + "[]" * len(dims)
            if len(dims) != 1 or dims[0] is None:
                raise ParseError("only one-dimensional `new T[n]` is supported")
            n, _ = self.expr(dims[0])
            return f"_newarr({DEFAULTS.get(ty, 'None')}, {n})", ty + "[]"
        if k == "cast":
            c, t = self.expr(e[2])
            ty = e[1]
            if ty in ("int", "long", "short", "byte"):
                return f"_cast_{ty}({c})", ty
            if ty == "double":
                return f"float({c})", "double"
            if ty == "float":
                return f"_f32({c})", "float"
            if ty == "char":
                return f"_cast_char({c})", "char"
            return c, ty  # reference cast: no-op
        if k == "unary":
            c, t = self.expr(e[2])
            if e[1] == "!":
                return f"(not {c})", "boolean"
            return f"({e[1]}{c})", t
        if k in ("postfix", "prefix"):
            return self.incdec(e, stmt)
        if k == "cond":
            c, _ = self.expr(e[1])
            a, ta = self.expr(e[2])
            b, tb = self.expr(e[3])
            return f"({a} if {c} else {b})", ta or tb
        if k == "instanceof":
            c, _ = self.expr(e[1])
            return f"isinstance({c}, {e[2]})", "boolean"
        if k == "assign":
            return self.assign(e, stmt)
        if k == "binary":
            _, op, l, r = e
            a, ta = self.expr(l)
            b, tb = self.expr(r)
            return self.binop_code(op, a, ta, b, tb)
        raise ParseError(f"cannot emit expression {k}")

    def incdec(self, e, stmt):
        kind, op, target = e
        delta = "+ 1" if op == "++" else "- 1"
        tc, tt = self.expr(target)
        if stmt:
            return f"{tc} {'+=' if op == '++' else '-='} 1", tt
        if target[0] != "name" or self.own_local(target[1]) is None:
            raise ParseError("++/-- inside an expression is supported for local variables only")
        if kind == "prefix":
            return f"({tc} := {tc} {delta})", tt
        undo = "- 1" if op == "++" else "+ 1"
        return f"(({tc} := {tc} {delta}) {undo})", tt

    @staticmethod
    def pure(e):
        """no side effects and no calls: evaluating it earlier or later cannot matter"""
        k = e[0]
        if k in ("num", "str", "chr", "bool", "null", "name", "this"):
            return True
        if k in ("paren",):
            return FunctionEmitter.pure(e[1])
        if k in ("unary", "cast"):
            return FunctionEmitter.pure(e[2])
        if k == "binary":
            return FunctionEmitter.pure(e[2]) and FunctionEmitter.pure(e[3])
        if k == "field":
            return FunctionEmitter.pure(e[1])
        return False

    def assign(self, e, stmt):
        _, op, target, value = e
        if stmt and target[0] == "index" and not self.pure(target[2]):
            # Java evaluates the array index BEFORE the right-hand side, Python after it: pin the index down first
            ac, at = self.expr(target[1])
            ic, _ = self.expr(target[2])
            self.tmp += 1
            self.pre.append(self.pad(0) + f"_ix{self.tmp} = {ic}")
            tc, tt = f"{ac}[_ix{self.tmp}]", (at[:-2] if at and at.endswith("[]") else None)
        else:
            tc, tt = self.expr(target)
        if not stmt:
            if target[0] != "name" or self.own_local(target[1]) is None or op != "=":
                raise ParseError("assignment inside an expression is supported for `local = value` only")
            vc, vt = self.expr(value)
            return f"({tc} := {self.wrap(vc, tt, value)})", tt
        if op == "=":
            if value[0] == "arrinit":
                return f"{tc} = {self.array_literal(tt or '[]', value)}", tt
            vc, vt = self.expr(value)
            return f"{tc} = {self.wrap(vc, tt, value)}", tt
        vc, vt = self.expr(value)
        if op == "+=" and (tt == "String" or vt == "String"):
            return f"{tc} = _cat({tc}, {vc})", "String"
        code, _ = self.binop_code(op[:-1], tc, tt, vc, vt)
        return f"{tc} = {self.wrap(code, tt, value, force=True)}", tt

    def binop_code(self, op, a, ta, b, tb):
        def num(t):
            return t in INTEGRAL or t in ("float", "double")
        if op == "+" and (ta == "String" or tb == "String"):
            return f"_cat({a}, {b})", "String"
        if op in ("&&", "||"):
            return f"({a} {'and' if op == '&&' else 'or'} {b})", "boolean"
        if op in ("==", "!="):
            if b == "None":
                return f"({a} {'is' if op == '==' else 'is not'} None)", "boolean"
            if a == "None":
                return f"({b} {'is' if op == '==' else 'is not'} None)", "boolean"
            if ta in PRIMITIVES or tb in PRIMITIVES:
                return f"({a} {op} {b})", "boolean"
            return f"_ref_{'eq' if op == '==' else 'ne'}({a}, {b})", "boolean"
        if op in ("<", ">", "<=", ">="):
            return f"({a} {op} {b})", "boolean"
        rt = None
        if num(ta) and num(tb):
            rt = next((c for c in ("double", "float", "long") if c in (ta, tb)), "int")
        if op == "/":
            if ta in INTEGRAL and tb in INTEGRAL:
                return f"_idiv({a}, {b})", rt
            if ta in ("float", "double") or tb in ("float", "double"):
                return f"({a} / {b})", rt
            return f"_div({a}, {b})", rt
        if op == "%":
            return f"_rem({a}, {b})", rt
        if op == ">>>":
            return f"_ushr({a}, {b}, {64 if ta == 'long' else 32})", ta
        if op in ("<<", ">>"):
            return f"({a} {op} {b})", (ta if ta in INTEGRAL else None)
        if op in ("&", "|", "^") and ta == "boolean":
            return f"({a} {op} {b})", "boolean"
        return f"({a} {op} {b})", rt

    def call(self, e):
        _, recv, name, args = e
        acodes = [self.expr(a)[0] for a in args]
        al = ", ".join(acodes)
        if recv is None:
            callee, rty = self.resolve_method(name)
            return f"{callee}({al})", rty
        if recv[0] == "name" and self.lookup_local(recv[1]) is None and recv[1] in self.em.classes and \
                name in self.em.classes[recv[1]].methods and self.resolve_name(recv[1])[1] is None:
            return f"{recv[1]}.{py_name(name)}({al})", self.em.classes[recv[1]].methods[name][0]
        rc, rt = self.expr(recv)
        rty = None
        if name in ("size", "length", "indexOf", "read", "compare", "parseInt", "hashCode", "lastIndexOf"):
            rty = "int"
        elif name in STRING_RETURNING:
            rty = "String"
        elif name in ("parseLong", "currentTimeMillis", "skip"):
            rty = "long"
        elif name == "charAt":
            rty = "char"
        elif name == "toCharArray":
            rty = "char[]"
        if name == "length" and not args:
            return f"len({rc})", "int"
        if rt in self.em.classes and name in self.em.classes[rt].methods:
            return f"{rc}.{py_name(name)}({al})", self.em.classes[rt].methods[name][0]
        if name in STRING_METHODS:
            return f"_s_{name}({', '.join([rc] + acodes)})", rty
        return f"{rc}.{py_name(name)}({al})", rty

    def new(self, e):
        _, ty, args, body = e
        acodes = [self.expr(a)[0] for a in args]
        if body is None:
            return f"{py_name(ty)}({', '.join(acodes)})", ty
        if args:
            raise ParseError("anonymous classes with constructor arguments are not supported")
        self.em.anon_count += 1
        cname = f"_Anon{self.em.anon_count}_{ty}"
        # the methods of the anonymous class become nested functions: they see this function's locals through closures
        self.em.emit_class(("class", ty, "class", set(), body), self.ind, self.pre, None, self.ctx, self, anon_name=cname)
        return f"{cname}()", ty

    # ---- statements ----
    def body(self, block):
        lines = []
        self.scopes.append({})
        for s in block[1]:
            lines.extend(self.stmt(s, 0))
        self.scopes.pop()
        return lines

    def block_lines(self, s, extra):
        """Statement s as the body of a compound statement at indentation +extra."""
        self.scopes.append({})
        if s[0] == "block":
            lines = []
            for x in s[1]:
                lines.extend(self.stmt(x, extra))
        else:
            lines = self.stmt(s, extra)
        self.scopes.pop()
        return lines or [self.pad(extra) + "pass"]

    def cond(self, e):
        c, _ = self.expr(e)
        if self.pre:
            raise ParseError("anonymous class inside a condition is not supported")
        return c

    def simple(self, code, extra):
        pre = self.flush_pre()
        if pre and extra:  # hoisted classes were emitted at the function's base indentation: shift them
            pre = [("    " * extra + l) if l else l for l in pre]
        return pre + [self.pad(extra) + code]

    @staticmethod
    def has_continue(s):
        """a `continue` that binds to the loop whose body s is"""
        k = s[0]
        if k == "continue":
            return True
        if k == "block":
            return any(FunctionEmitter.has_continue(x) for x in s[1])
        if k == "if":
            return FunctionEmitter.has_continue(s[2]) or (s[3] is not None and FunctionEmitter.has_continue(s[3]))
        if k == "try":
            return FunctionEmitter.has_continue(s[1]) or any(FunctionEmitter.has_continue(c[2]) for c in s[2]) or \
                (s[3] is not None and FunctionEmitter.has_continue(s[3]))
        if k == "switch":
            return any(FunctionEmitter.has_continue(x) for _, st in s[2] for x in st)
        return False

    def stmt(self, s, extra):
        k = s[0]
        P = self.pad(extra)
        if k == "empty":
            return []
        if k == "block":
            return self.block_lines(s, extra)
        if k == "local":
            lines = []
            for ty, name, init in s[1]:
                val = DEFAULTS.get(ty, "None") if init is None else self.init_value(ty, init)
                self.declare(name, ty)
                lines.extend(self.simple(f"{py_name(name)} = {val}", extra))
            return lines
        if k == "expr":
            e = s[1]
            code, _ = self.expr(e, stmt=e[0] in ("assign", "postfix", "prefix"))
            return self.simple(code, extra)
        if k == "return":
            if s[1] is None:
                return [P + "return"]
            code, t = self.expr(s[1])
            if t != "String":
                code = self.wrap(code, self.rtype, s[1])
            return self.simple("return " + code, extra)
        if k == "throw":
            code, _ = self.expr(s[1])
            return self.simple("raise " + code, extra)
        if k == "break":
            return [P + "break"]
        if k == "continue":
            if self.in_switch:
                raise ParseError("continue inside a switch is not supported")
            return [P + "continue"]
        if k == "if":
            lines = [P + f"if {self.cond(s[1])}:"] + self.block_lines(s[2], extra + 1)
            if s[3] is not None:
                if s[3][0] == "if":
                    sub = self.stmt(s[3], extra)
                    sub[0] = P + "el" + sub[0].lstrip()
                    lines.extend(sub)
                else:
                    lines.append(P + "else:")
                    lines.extend(self.block_lines(s[3], extra + 1))
            return lines
        if k == "while":
            sw, self.in_switch = self.in_switch, 0
            lines = [P + f"while {self.cond(s[1])}:"] + self.block_lines(s[2], extra + 1)
            self.in_switch = sw
            return lines
        if k == "dowhile":
            sw, self.in_switch = self.in_switch, 0
            self.tmp += 1
            first = f"_first{self.tmp}"
            lines = [P + f"{first} = True", P + f"while {first} or ({self.cond(s[2])}):", self.pad(extra + 1) + f"{first} = False"]
            lines += self.block_lines(s[1], extra + 1)
            self.in_switch = sw
            return lines
        if k == "for":
            _, init, cond, upd, body = s
            sw, self.in_switch = self.in_switch, 0
            self.scopes.append({})
            lines = []
            for x in init:
                lines.extend(self.stmt(x, extra))
            upd_lines = [self.expr(u, stmt=u[0] in ("assign", "postfix", "prefix"))[0] for u in upd]
            cond_code = self.cond(cond) if cond is not None else "True"
            if upd and self.has_continue(body):
                # `continue` must still run the update: it sits at the top of the loop, skipped on the first pass
                self.tmp += 1
                first = f"_first{self.tmp}"
                lines += [P + f"{first} = True", P + "while True:", self.pad(extra + 1) + f"if not {first}:"]
                lines += [self.pad(extra + 2) + u for u in upd_lines]
                lines += [self.pad(extra + 1) + f"{first} = False", self.pad(extra + 1) + f"if not ({cond_code}):",
                          self.pad(extra + 2) + "break"]
                lines += self.block_lines(body, extra + 1)
            else:
                lines.append(P + f"while {cond_code}:")
                body_lines = self.block_lines(body, extra + 1)
                if upd and body_lines == [self.pad(extra + 1) + "pass"]:
                    body_lines = []
                lines += body_lines + [self.pad(extra + 1) + u for u in upd_lines]
            self.scopes.pop()
            self.in_switch = sw
            return lines
        if k == "foreach":
            _, ty, name, it, body = s
            sw, self.in_switch = self.in_switch, 0
            self.scopes.append({name: ty})
            lines = [P + f"for {py_name(name)} in _iter({self.cond(it)}):"] + self.block_lines(body, extra + 1)
            self.scopes.pop()
            self.in_switch = sw
            return lines
        if k == "try":
            _, body, catches, fin = s
            lines = [P + "try:"] + self.block_lines(body, extra + 1)
            for ty, name, blk in catches:
                self.scopes.append({name: ty})
                lines.append(P + f"except {py_name(ty)} as {py_name(name)}:")
                lines.extend(self.block_lines(blk, extra + 1))
                self.scopes.pop()
            if fin is not None:
                lines.append(P + "finally:")
                lines.extend(self.block_lines(fin, extra + 1))
            return lines
        if k == "switch":
            _, e, groups = s
            for gi, (labels, _) in enumerate(groups):
                if None in labels and gi != len(groups) - 1:
                    raise ParseError("switch: `default` must be the last group")
            self.tmp += 1
            sv, m = f"_sw{self.tmp}", f"_m{self.tmp}"
            lines = [P + f"{sv} = {self.cond(e)}", P + f"{m} = False", P + "while True:  # switch"]
            self.in_switch += 1
            for labels, stmts in groups:
                test = "True" if None in labels else f"{m} or " + " or ".join(f"{sv} == {self.expr(l)[0]}" for l in labels)
                lines.append(self.pad(extra + 1) + f"if {test}:")
                lines.append(self.pad(extra + 2) + f"{m} = True")
                self.scopes.append({})
                for x in stmts:
                    lines.extend(self.stmt(x, extra + 2))
                self.scopes.pop()
            self.in_switch -= 1
            lines.append(self.pad(extra + 1) + "break")
            return lines
        raise ParseError(f"cannot emit statement {k}")


def syntax_check(java_source):
    """Parse only, with the lenient grammar: raises ParseError (with a line number) on anything that is not Java as this parser
    knows it.  For the repo's own Java files, which no compiler has ever seen (no JDK in the image)."""
    return Parser(java_source, lenient=True).compilation_unit()


def transliterate(java_source):
    """Java source text -> Python module text (imports j2py_runtime)."""
    classes = Parser(java_source).compilation_unit()
    return Emitter(classes).emit_module()
