"""tests/java_pin/j2py_runtime.py -- what the Python text produced by j2py.py runs on: Java's primitive arithmetic and the handful
of java.lang / java.util / java.io classes that KmerGutsJava.java touches, each restricted to the behaviour the Java API
documents (index checks, insertion order, line terminators, %f rounding ...).  Knows nothing about k-mers.

Names that collide with Python (`format`, `print`, `set`, `abs`, `in`, ...) carry a trailing underscore, exactly as
j2py.py's py_name() writes them.  `Exception` here is the root of the JAVA exceptions: a Python-level error inside the
transliterated code (a bug of the tool) is therefore NOT swallowed by the source's `catch (Exception ex)`.
"""
import builtins
import functools
import gzip
import math
import os
import struct
import sys
import time
import traceback
from decimal import ROUND_HALF_UP, Decimal

# =====================================================================================================================
# java.lang.Object / exceptions
# =====================================================================================================================


class JObject:
    __slots__ = ()

    def getClass(self):
        return type(self)

    def equals(self, other):
        return self is other

    def hashCode(self):
        return id(self) & 0x7FFFFFFF

    def toString(self):
        return f"{type(self).__name__}@{self.hashCode():x}"


class Throwable(builtins.Exception):
    def __init__(self, msg=None):
        if isinstance(msg, BaseException):   # new X(cause): the message is cause.toString()
            self.cause = msg
            msg = msg.toString() if isinstance(msg, Throwable) else repr(msg)
        builtins.Exception.__init__(self, msg)
        self.msg = msg

    def getMessage(self):
        return self.msg

    def toString(self):
        name = "java.lang." + type(self).__name__
        return name if self.msg is None else f"{name}: {self.msg}"

    def printStackTrace(self):
        traceback.print_exception(type(self), self, self.__traceback__, file=sys.stderr)


class Exception(Throwable):  # noqa: A001 -- java.lang.Exception, on purpose (see the module docstring)
    pass


class RuntimeException(Exception):
    pass


class IllegalStateException(RuntimeException):
    pass


class IllegalArgumentException(RuntimeException):
    pass


class NumberFormatException(IllegalArgumentException):
    pass


class IndexOutOfBoundsException(RuntimeException):
    pass


class StringIndexOutOfBoundsException(IndexOutOfBoundsException):
    pass


class ArrayIndexOutOfBoundsException(IndexOutOfBoundsException):
    pass


class NegativeArraySizeException(RuntimeException):
    pass


class ArithmeticException(RuntimeException):
    pass


class NullPointerException(RuntimeException):
    pass


class NoSuchElementException(RuntimeException):
    pass


class IOException(Exception):
    pass


class EOFException(IOException):
    pass


class FileNotFoundException(IOException):
    pass


# =====================================================================================================================
# primitive arithmetic
# =====================================================================================================================
def _i32(x):
    x &= 0xFFFFFFFF
    return x - 0x100000000 if x & 0x80000000 else x


def _i64(x):
    x &= 0xFFFFFFFFFFFFFFFF
    return x - 0x10000000000000000 if x & 0x8000000000000000 else x


def _i16(x):
    x &= 0xFFFF
    return x - 0x10000 if x & 0x8000 else x


def _i8(x):
    x &= 0xFF
    return x - 0x100 if x & 0x80 else x


_F32 = struct.Struct("<f")


def _f32(x):
    """round a double to the nearest binary32 (what assigning to a Java float does)"""
    try:
        return _F32.unpack(_F32.pack(x))[0]
    except OverflowError:
        return math.copysign(math.inf, x)


def _float_to_integral(x, lo, hi):
    if x != x:
        return 0
    if x <= lo:
        return lo
    if x >= hi:
        return hi
    return int(x)  # truncates toward zero


def _cast_int(x):
    if isinstance(x, float):
        return _float_to_integral(x, -0x80000000, 0x7FFFFFFF)
    if isinstance(x, str):
        return ord(x)
    return _i32(x)


def _cast_long(x):
    if isinstance(x, float):
        return _float_to_integral(x, -0x8000000000000000, 0x7FFFFFFFFFFFFFFF)
    if isinstance(x, str):
        return ord(x)
    return _i64(x)


def _cast_short(x):
    return _i16(_cast_int(x))


def _cast_byte(x):
    return _i8(_cast_int(x))


def _cast_char(x):
    return x if isinstance(x, str) else chr(_cast_int(x) & 0xFFFF)


def _idiv(a, b):
    if b == 0:
        raise ArithmeticException("/ by zero")
    q = abs(a) // abs(b)
    return -q if (a < 0) != (b < 0) else q


def _div(a, b):
    if isinstance(a, int) and isinstance(b, int):
        return _idiv(a, b)
    try:
        return a / b
    except ZeroDivisionError:
        return math.nan if a == 0 or a != a else math.copysign(math.inf, a) * math.copysign(1.0, b)


def _rem(a, b):
    if isinstance(a, int) and isinstance(b, int):
        if b == 0:
            raise ArithmeticException("/ by zero")
        r = abs(a) % abs(b)
        return -r if a < 0 else r
    return math.fmod(a, b) if b != 0 else math.nan


def _ushr(a, n, bits):
    return (a & ((1 << bits) - 1)) >> (n & (bits - 1))


def _ref_eq(a, b):
    if isinstance(a, (int, float, str)) or isinstance(b, (int, float, str)):
        return a == b
    return a is b


def _ref_ne(a, b):
    return not _ref_eq(a, b)


def _newarr(default, n):
    if n < 0:
        raise NegativeArraySizeException(str(n))
    return [default] * n


def _iter(x):
    return iter(x._a) if isinstance(x, ArrayList) else iter(x)


def _jstr(x):
    """String.valueOf(x)"""
    if x is None:
        return "null"
    if x is True:
        return "true"
    if x is False:
        return "false"
    if isinstance(x, str):
        return x
    if isinstance(x, int):
        return str(x)
    if isinstance(x, float):
        return Double.toString(x)
    if isinstance(x, (JObject, Throwable)):
        return x.toString()
    return str(x)


def _cat(a, b):
    return _jstr(a) + _jstr(b)


# =====================================================================================================================
# java.lang.String (methods as functions: Java strings are Python strs, Java chars are 1-character strs)
# =====================================================================================================================
def _s_trim(s):
    a, b = 0, len(s)
    while a < b and s[a] <= " ":
        a += 1
    while b > a and s[b - 1] <= " ":
        b -= 1
    return s[a:b]


def _s_charAt(s, i):
    if i < 0 or i >= len(s):
        raise StringIndexOutOfBoundsException(f"index {i}, length {len(s)}")
    return s[i]


def _s_substring(s, a, b=None):
    e = len(s) if b is None else b
    if a < 0 or e > len(s) or a > e:
        raise StringIndexOutOfBoundsException(f"begin {a}, end {e}, length {len(s)}")
    return s[a:e]


def _s_startsWith(s, p):
    return s.startswith(p)


def _s_endsWith(s, p):
    return s.endswith(p)


def _s_indexOf(s, x, start=0):
    return s.find(x, max(start, 0))


def _s_lastIndexOf(s, x):
    return s.rfind(x)


def _s_contains(s, x):
    return x in s if isinstance(s, str) else s.contains(x)


def _s_toCharArray(s):
    return list(s)


def _s_equals(a, b):
    if isinstance(a, str):
        return isinstance(b, str) and a == b
    if isinstance(a, (int, float)):
        return a == b
    return a.equals(b)


def _s_hashCode(a):
    if isinstance(a, str):
        h = 0
        for ch in a:
            h = (31 * h + ord(ch)) & 0xFFFFFFFF
        return _i32(h)
    if isinstance(a, int):
        return _i32(a ^ (a >> 32))
    return a.hashCode()


def _s_isEmpty(a):
    return len(a) == 0 if isinstance(a, str) else a.isEmpty()


def _s_toString(a):
    return a if isinstance(a, str) else a.toString()


def _s_toUpperCase(a):
    return a.upper()


def _s_toLowerCase(a):
    return a.lower()


def _s_split(s, regex):
    import re
    parts = re.split(regex, s)
    while parts and parts[-1] == "":
        parts.pop()
    return parts


class String:
    @staticmethod
    def valueOf(x):
        return _jstr(x)

    @staticmethod
    def format_(fmt, *args):
        out, i, k = [], 0, 0
        n = len(fmt)
        while i < n:
            ch = fmt[i]
            if ch != "%":
                out.append(ch)
                i += 1
                continue
            j = i + 1
            while j < n and (fmt[j].isdigit() or fmt[j] in ".-,+ 0#("):
                j += 1
            spec, conv = fmt[i + 1:j], fmt[j]
            i = j + 1
            if conv == "%":
                out.append("%")
                continue
            if conv == "n":
                out.append("\n")
                continue
            arg = args[k]
            k += 1
            width, _, prec = spec.partition(".")
            left = width.startswith("-")
            w = int(width.lstrip("-")) if width.lstrip("-").isdigit() else 0
            if conv == "d":
                if not isinstance(arg, int) or isinstance(arg, bool):
                    raise IllegalArgumentException(f"%d with {type(arg).__name__}")
                text = str(arg)
            elif conv == "s":
                text = _jstr(arg)
            elif conv == "c":
                text = arg if isinstance(arg, str) else chr(arg)
            elif conv == "f":
                text = java_format_f(arg, int(prec) if prec else 6)
            else:
                raise IllegalArgumentException(f"conversion %{conv} is not modelled")
            if len(text) < w:
                text = text.ljust(w) if left else text.rjust(w)
            out.append(text)
        return "".join(out)


def java_format_f(x, prec):
    """Formatter's %f: the argument (a float is widened to double first) goes through FloatingDecimal's shortest
    round-trip digit string, which is then rounded HALF_UP to `prec` places (FormattedFloatingDecimal.applyPrecision)."""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (math.inf, -math.inf):
        return "Infinity" if x > 0 else "-Infinity"
    d = Decimal(repr(abs(x))).quantize(Decimal(1).scaleb(-prec), rounding=ROUND_HALF_UP)
    text = format(d, "f")
    return ("-" if math.copysign(1.0, x) < 0 else "") + text


class Integer:
    MAX_VALUE = 0x7FFFFFFF
    MIN_VALUE = -0x80000000

    @staticmethod
    def parseInt(s):
        if s is None:
            raise NumberFormatException("null")
        body = s[1:] if s[:1] in ("-", "+") else s
        if not body or not all("0" <= c <= "9" for c in body):
            raise NumberFormatException(f'For input string: "{s}"')
        v = int(s)
        if v < Integer.MIN_VALUE or v > Integer.MAX_VALUE:
            raise NumberFormatException(f'For input string: "{s}"')
        return v

    @staticmethod
    def compare(a, b):
        return -1 if a < b else (0 if a == b else 1)

    @staticmethod
    def valueOf(x):
        return Integer.parseInt(x) if isinstance(x, str) else x


class Long:
    MAX_VALUE = 0x7FFFFFFFFFFFFFFF
    MIN_VALUE = -0x8000000000000000

    @staticmethod
    def parseLong(s):
        if s is None:
            raise NumberFormatException("null")
        body = s[1:] if s[:1] in ("-", "+") else s
        if not body or not all("0" <= c <= "9" for c in body):
            raise NumberFormatException(f'For input string: "{s}"')
        v = int(s)
        if v < Long.MIN_VALUE or v > Long.MAX_VALUE:
            raise NumberFormatException(f'For input string: "{s}"')
        return v

    compare = Integer.compare


class Float:
    @staticmethod
    def intBitsToFloat(bits):
        return struct.unpack("<f", struct.pack("<I", bits & 0xFFFFFFFF))[0]

    @staticmethod
    def floatToIntBits(f):
        return _i32(struct.unpack("<I", struct.pack("<f", f))[0])


class Double:
    @staticmethod
    def toString(x):
        if x != x:
            return "NaN"
        if x in (math.inf, -math.inf):
            return "Infinity" if x > 0 else "-Infinity"
        a = abs(x)
        if a != 0 and (a < 1e-3 or a >= 1e7):   # computerised scientific notation
            m, e = repr(x).lower().split("e") if "e" in repr(x).lower() else (None, None)
            if m is None:
                d = Decimal(repr(x))
                e = d.adjusted()
                m = str(d.scaleb(-e).normalize())
            if "." not in m:
                m += ".0"
            return f"{m}E{int(e)}"
        r = repr(x)
        return r if "." in r else r + ".0"


class Math:
    @staticmethod
    def abs_(x):
        return x if x == Integer.MIN_VALUE else abs(x)   # Math.abs(Integer.MIN_VALUE) stays negative

    @staticmethod
    def max_(a, b):
        return a if a >= b else b

    @staticmethod
    def min_(a, b):
        return a if a <= b else b


# =====================================================================================================================
# java.util
# =====================================================================================================================
class ArrayList(JObject):
    __slots__ = ("_a",)

    def __init__(self, x=None):
        if x is None or isinstance(x, int):   # new ArrayList<T>(initialCapacity)
            self._a = []
        elif isinstance(x, ArrayList):
            self._a = list(x._a)
        else:
            self._a = list(x)

    def _check(self, i):
        if i < 0 or i >= len(self._a):
            raise IndexOutOfBoundsException(f"Index: {i}, Size: {len(self._a)}")

    def add(self, *a):
        if len(a) == 2:   # add(int index, E element)
            if a[0] < 0 or a[0] > len(self._a):
                raise IndexOutOfBoundsException(f"Index: {a[0]}, Size: {len(self._a)}")
            self._a.insert(a[0], a[1])
            return None
        self._a.append(a[0])
        return True

    def get(self, i):
        if i < 0 or i >= len(self._a):
            raise IndexOutOfBoundsException(f"Index: {i}, Size: {len(self._a)}")
        return self._a[i]

    def set_(self, i, v):
        self._check(i)
        old = self._a[i]
        self._a[i] = v
        return old

    def remove(self, i):
        if isinstance(i, int):   # remove(int index)
            self._check(i)
            return self._a.pop(i)
        for k, v in enumerate(self._a):   # remove(Object)
            if _s_equals(v, i):
                del self._a[k]
                return True
        return False

    def size(self):
        return len(self._a)

    def isEmpty(self):
        return not self._a

    def clear(self):
        self._a.clear()

    def contains(self, x):
        return any(_s_equals(v, x) for v in self._a)

    def addAll(self, other):
        self._a.extend(_iter(other))
        return True

    # Queue (LinkedList)
    def poll(self):
        return self._a.pop(0) if self._a else None

    def peek(self):
        return self._a[0] if self._a else None

    def __iter__(self):
        return iter(self._a)

    def __len__(self):
        return len(self._a)


LinkedList = ArrayList


class _Key:
    """A map key that is a Java object: hashed by its field values, compared with ITS OWN (transliterated) equals()."""
    __slots__ = ("o", "h")

    def __init__(self, o):
        self.o = o
        self.h = hash(tuple(sorted(vars(o).items()))) if hasattr(o, "__dict__") else \
            hash(tuple(getattr(o, s) for s in type(o).__slots__))

    def __hash__(self):
        return self.h

    def __eq__(self, other):
        return self.o.equals(other.o)


def _k(key):
    return _Key(key) if isinstance(key, JObject) else key


class HashMap(JObject):
    """Also LinkedHashMap: a Python dict keeps insertion order, and re-inserting a key keeps its place (as LinkedHashMap does
    in insertion-order mode).  Nothing in the source iterates a plain HashMap."""
    __slots__ = ("_d",)

    def __init__(self, *_):
        self._d = {}

    def put(self, k, v):
        k = _k(k)
        old = self._d.get(k)
        self._d[k] = v
        return old

    def get(self, k):
        return self._d.get(_k(k))

    def containsKey(self, k):
        return _k(k) in self._d

    def remove(self, k):
        return self._d.pop(_k(k), None)

    def size(self):
        return len(self._d)

    def isEmpty(self):
        return not self._d

    def clear(self):
        self._d.clear()

    def keySet(self):
        return ArrayList([k.o if isinstance(k, _Key) else k for k in self._d])

    def values(self):
        return ArrayList(self._d.values())


LinkedHashMap = HashMap


class Collections:
    @staticmethod
    def sort(lst, comparator=None):
        """List.sort: a stable merge sort (TimSort), as Python's -- driven by the SOURCE's comparator"""
        if comparator is None:
            lst._a.sort()
        else:
            lst._a.sort(key=functools.cmp_to_key(comparator.compare))


class Arrays:
    @staticmethod
    def asList(*a):
        return ArrayList(a[0] if len(a) == 1 and isinstance(a[0], (list, tuple)) else a)


class StringBuilder(JObject):
    __slots__ = ("_p",)

    def __init__(self, x=None):
        self._p = [x] if isinstance(x, str) else []

    def append(self, x):
        self._p.append(_jstr(x))
        return self

    def length(self):
        return sum(len(p) for p in self._p)

    def __len__(self):
        return self.length()

    def toString(self):
        return "".join(self._p)


class StringTokenizer(JObject):
    __slots__ = ("_t", "_i")

    def __init__(self, s, delims=" \t\n\r\f"):
        toks, cur = [], []
        for ch in s:
            if ch in delims:
                if cur:
                    toks.append("".join(cur))
                    cur = []
            else:
                cur.append(ch)
        if cur:
            toks.append("".join(cur))
        self._t, self._i = toks, 0

    def hasMoreTokens(self):
        return self._i < len(self._t)

    def nextToken(self):
        if self._i >= len(self._t):
            raise NoSuchElementException()
        self._i += 1
        return self._t[self._i - 1]


# =====================================================================================================================
# java.io
# =====================================================================================================================
class File(JObject):
    __slots__ = ("_p",)

    def __init__(self, a, b=None):
        if a is None and b is None:
            raise NullPointerException()
        a = a._p if isinstance(a, File) else a
        self._p = a if b is None else os.path.join(a, b)

    def getName(self):
        return os.path.basename(self._p)

    def getPath(self):
        return self._p

    def getAbsolutePath(self):
        return os.path.abspath(self._p)

    def getCanonicalPath(self):
        return os.path.realpath(self._p)

    def exists(self):
        return os.path.exists(self._p)

    def delete(self):
        try:
            os.remove(self._p)
            return True
        except OSError:
            return False

    def mkdirs(self):
        os.makedirs(self._p, exist_ok=True)
        return True

    def toString(self):
        return self._p


def _path(x):
    return x._p if isinstance(x, File) else x


class InputStream(JObject):
    """An in-memory byte stream: read() -> 0..255 or -1, skip(n) -> bytes skipped"""
    __slots__ = ("_b", "_i")

    def __init__(self, data=b""):
        self._b, self._i = data, 0

    def read(self):
        i = self._i
        if i >= len(self._b):
            return -1
        self._i = i + 1
        return self._b[i]

    def skip(self, n):
        k = max(0, min(n, len(self._b) - self._i))
        self._i += k
        return k

    def _rest(self):
        r = self._b[self._i:]
        self._i = len(self._b)
        return r

    def close(self):
        pass


class FileInputStream(InputStream):
    __slots__ = ()

    def __init__(self, f):
        try:
            with open(_path(f), "rb") as fh:
                InputStream.__init__(self, fh.read())
        except OSError as e:
            raise FileNotFoundException(str(e))


def BufferedInputStream(s):
    return s


class GZIPInputStream(InputStream):
    __slots__ = ()

    def __init__(self, s):
        try:
            InputStream.__init__(self, gzip.decompress(s._rest()))
        except (OSError, EOFError) as e:
            raise IOException(str(e))


class DataInputStream(InputStream):
    __slots__ = ()

    def __init__(self, s):
        InputStream.__init__(self, s._rest())

    def _take(self, n):
        if len(self._b) - self._i < n:
            self._i = len(self._b)
            raise EOFException()
        r = self._b[self._i:self._i + n]
        self._i += n
        return r

    def readLong(self):
        return struct.unpack(">q", self._take(8))[0]

    def readInt(self):
        return struct.unpack(">i", self._take(4))[0]


class FileOutputStream(JObject):
    __slots__ = ("_f",)

    def __init__(self, f):
        self._f = open(_path(f), "wb")

    def write(self, b):
        self._f.write(b)

    def close(self):
        self._f.close()


def BufferedOutputStream(s):
    return s


class DataOutputStream(JObject):
    __slots__ = ("_s",)

    def __init__(self, s):
        self._s = s

    def writeLong(self, v):
        self._s.write(struct.pack(">q", v))

    def writeInt(self, v):
        self._s.write(struct.pack(">i", v))

    def close(self):
        self._s.close()


class InputStreamReader(JObject):
    """bytes -> chars, ISO-8859-1 (one char per byte; the fixtures are ASCII)"""
    __slots__ = ("_s",)

    def __init__(self, s):
        self._s = s


def FileReader(f):
    return InputStreamReader(FileInputStream(f))


class BufferedReader(JObject):
    __slots__ = ("_t", "_i")

    def __init__(self, r):
        self._t = r._s._rest().decode("latin-1")
        self._i = 0

    def readLine(self):
        """a line ends with \\n, \\r, or \\r\\n; null at the end of the stream"""
        t, i = self._t, self._i
        n = len(t)
        if i >= n:
            return None
        a = t.find("\n", i)
        b = t.find("\r", i)
        if a < 0 and b < 0:
            self._i = n
            return t[i:]
        if b < 0 or (0 <= a < b):
            self._i = a + 1
            return t[i:a]
        self._i = b + 2 if t[b + 1:b + 2] == "\n" else b + 1
        return t[i:b]

    def close(self):
        pass


class _StdinStream(InputStream):
    __slots__ = ()

    def __init__(self):
        InputStream.__init__(self, b"")


class FileWriter(JObject):
    __slots__ = ("_f",)

    def __init__(self, f):
        self._f = open(_path(f), "w", encoding="latin-1", newline="")


class PrintStream(JObject):
    __slots__ = ("_f",)

    def __init__(self, f):
        self._f = f

    def println(self, x=""):
        self._f.write(_jstr(x) + "\n")

    def print_(self, x):
        self._f.write(_jstr(x))

    def flush(self):
        self._f.flush()

    def close(self):
        pass


class PrintWriter(PrintStream):
    __slots__ = ("_own",)

    def __init__(self, target):
        if isinstance(target, FileWriter):
            self._f, self._own = target._f, True
        elif isinstance(target, PrintStream):
            self._f, self._own = target._f, False
        else:
            self._f, self._own = target, False

    def close(self):
        if self._own:
            self._f.close()
        else:
            self._f.flush()


class System:
    out = PrintStream(sys.stdout)
    err = PrintStream(sys.stderr)
    in_ = _StdinStream()

    @staticmethod
    def currentTimeMillis():
        return int(time.time() * 1000)

    @staticmethod
    def getProperty(name):
        return {"java.io.tmpdir": os.environ.get("TMPDIR", "/tmp"), "line.separator": "\n"}.get(name)

    @staticmethod
    def exit(code):
        raise SystemExit(code)


__all__ = [n for n in dict(globals()) if not n.startswith("__") and n not in
           ("builtins", "functools", "gzip", "math", "os", "struct", "sys", "time", "traceback", "Decimal", "ROUND_HALF_UP")]
