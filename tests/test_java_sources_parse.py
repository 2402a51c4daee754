"""The repo's own Java files have never met a compiler (no JDK in the image or on the GPU boxes).  The least that can be
checked here: they are syntactically Java -- tests/java_pin/j2py.py's parser in its lenient (syntax-only) mode reads them,
and the same parser rejects them once a brace, a semicolon or a parenthesis is taken away.  Types, imports and the FFM API
usage are NOT checked by this; `tests/java_pin/pin_oracle.sh` compiles both on a box with a JDK."""
import importlib.util
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = [os.path.join(ROOT, "kmergutsjava_b200", "java", "KmerGutsGpu.java"), os.path.join(ROOT, "tests", "java_pin", "GoldenDump.java")]


@pytest.fixture(scope="module")
def j2py():
    spec = importlib.util.spec_from_file_location("j2py", os.path.join(ROOT, "tests", "java_pin", "j2py.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p) for p in FILES])
def test_java_file_is_syntactically_java(j2py, path):
    src = open(path).read()
    classes = j2py.syntax_check(src)
    assert len(classes) == 1 and classes[0][1] == os.path.basename(path)[:-5]
    members = classes[0][4]
    assert sum(1 for m in members if m[0] == "method") >= 3
    # the check has teeth: damage the file in a few places and the parser must say so
    code_lines = [i for i, l in enumerate(src.splitlines()) if l.rstrip().endswith(";") and "//" not in l and "*" not in l]
    lines = src.splitlines()
    for victim in (code_lines[len(code_lines) // 3], code_lines[2 * len(code_lines) // 3]):
        broken = lines[:]
        broken[victim] = broken[victim].rstrip()[:-1]          # drop a semicolon
        with pytest.raises(j2py.ParseError):
            j2py.syntax_check("\n".join(broken))
    k = src.rindex("}")
    with pytest.raises(j2py.ParseError):
        j2py.syntax_check(src[:k] + src[k + 1:])              # drop the last closing brace
    k = re.search(r"\)\s*(throws \w+\s*)?\{", src).start()
    with pytest.raises(j2py.ParseError):
        j2py.syntax_check(src[:k] + src[k + 1:])              # drop the closing parenthesis of the first method header


def test_java_snippets_in_integration_md_parse(j2py):
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```java\n(.*?)```", md, re.S)
    assert len(blocks) >= 2
    for b in blocks:
        j2py.syntax_check("class Snippet { void f() throws Throwable {\n" + b + "\n} }")


def test_binding_names_the_abi_it_binds(j2py):
    """every kg_* symbol the FFM binding looks up is declared in include/*.h (a typo would only show at run time on a JVM)"""
    src = open(FILES[0]).read()
    wanted = set(re.findall(r'h\("(kg_[a-z_0-9]+)"', src))
    assert len(wanted) >= 15
    headers = "".join(open(os.path.join(ROOT, "include", h)).read() for h in os.listdir(os.path.join(ROOT, "include")))
    declared = set(re.findall(r"\b(kg_[a-z_0-9]+)\s*\(", headers))
    assert wanted <= declared, sorted(wanted - declared)


def test_binding_descriptors_match_the_c_declarations():
    """Every FunctionDescriptor of the FFM binding against the C declaration of the symbol in include/*.h: same arity, and per
    argument pointer -> ADDRESS, int -> JAVA_INT, size_t / uint64_t -> JAVA_LONG, float -> JAVA_FLOAT; return int -> JAVA_INT,
    void -> ofVoid, pointer -> ADDRESS.  (A mismatch compiles fine on a JVM and corrupts the call at run time.)"""
    src = open(FILES[0]).read()
    inc = os.path.join(ROOT, "include")
    headers = "".join(open(os.path.join(inc, h)).read() for h in sorted(os.listdir(inc)))
    headers = re.sub(r"/\*.*?\*/", " ", headers, flags=re.S)
    headers = re.sub(r"//[^\n]*", " ", headers)

    def kind(ctype):
        t = ctype.strip()
        if "*" in t or "[" in t:
            return "ADDRESS"
        base = re.sub(r"\b(const|unsigned|signed)\b", "", t).split()
        base = base[0] if base else ""
        return {"int": "JAVA_INT", "int32_t": "JAVA_INT", "uint32_t": "JAVA_INT", "size_t": "JAVA_LONG", "uint64_t": "JAVA_LONG",
                "int64_t": "JAVA_LONG", "long": "JAVA_LONG", "float": "JAVA_FLOAT", "void": "void"}.get(base, "?" + t)

    checked = 0
    for name, how, args in re.findall(r'h\("(kg_[a-z_0-9]+)",\s*FunctionDescriptor\.(of|ofVoid)\(([^;]*?)\)\);', src):
        m = re.search(r"([\w\s\*]+?)\b" + name + r"\s*\(([^)]*)\)\s*;", headers)
        assert m, f"{name} is not declared in include/"
        ret, params = m.group(1), m.group(2)
        cparams = [] if params.strip() in ("", "void") else [kind(re.sub(r"\b\w+\s*(\[[^\]]*\])?\s*$", lambda x: x.group(1) or "", p.strip())
                                                                  if not p.strip().endswith("*") else p) for p in params.split(",")]
        jargs = [a.strip() for a in args.split(",") if a.strip()]
        if how == "of":
            jret, jargs = jargs[0], jargs[1:]
        else:
            jret = "void"
        assert jret == kind(ret), (name, jret, ret)
        assert jargs == cparams, (name, jargs, cparams, params)
        checked += 1
    assert checked >= 15


def test_binding_struct_layouts_match_the_c_structs():
    """PARAMS / CALL / OTU / HIT of the FFM binding against kg_params / kg_call / kg_otu / kg_hit in include/kmerguts.h: same
    field names in the same order, int32/uint32 -> JAVA_INT, float -> JAVA_FLOAT, int32[5] -> sequenceLayout(5, JAVA_INT); and
    the hard-coded byte offsets the reader uses (o + 4, o + 24 ...) are the ones these layouts imply."""
    src = open(FILES[0]).read()
    hdr = open(os.path.join(ROOT, "include", "kmerguts.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    bufsz = int(re.search(r"#define\s+KG_OI_BUFSZ\s+(\d+)", hdr).group(1))
    for jname, cname in (("PARAMS", "kg_params"), ("CALL", "kg_call"), ("OTU", "kg_otu"), ("HIT", "kg_hit")):
        body = re.search(r"typedef struct " + cname + r"\s*\{(.*?)\}\s*" + cname + r"\s*;", hdr, re.S).group(1)
        cfields = []
        for decl in body.split(";"):
            m = re.match(r"\s*(u?int32_t|float)\s+(\w+)\s*(\[\s*(\w+)\s*\])?\s*$", decl)
            if m:
                n = None if m.group(4) is None else (bufsz if m.group(4) == "KG_OI_BUFSZ" else int(m.group(4)))
                cfields.append((m.group(2), "JAVA_FLOAT" if m.group(1) == "float" else "JAVA_INT", n))
            else:
                assert not decl.strip(), (cname, decl)
        jbody = re.search(r"StructLayout " + jname + r" = MemoryLayout\.structLayout\((.*?)\);", src, re.S).group(1)
        jfields = []
        for m in re.finditer(r'(?:MemoryLayout\.sequenceLayout\((\d+),\s*(JAVA_\w+)\)|(JAVA_\w+))\.withName\("(\w+)"\)', jbody):
            jfields.append((m.group(4), m.group(2) or m.group(3), int(m.group(1)) if m.group(1) else None))
        assert jfields == cfields, (jname, jfields, cfields)
    # offsets used when the records are read back
    call_reads = re.findall(r"cs\.get\((JAVA_\w+), o(?: \+ (\d+))?\)", src)
    assert [(t, int(o or 0)) for t, o in call_reads] == [("JAVA_INT", 0), ("JAVA_INT", 4), ("JAVA_INT", 8), ("JAVA_INT", 12), ("JAVA_INT", 16),
                                                          ("JAVA_INT", 20), ("JAVA_FLOAT", 24), ("JAVA_INT", 28)]
    assert "o + 4 + 4L * j" in src and "o + 24 + 4L * j" in src and bufsz == 5     # kg_otu: n, count[5] at 4, oI[5] at 24
    assert [int(x) for x in re.findall(r"p\.set\(JAVA_INT, (\d+),", src)] == [0, 4, 8, 12, 16]   # kg_params
