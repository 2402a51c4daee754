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
