"""The Rust binding under rust/ cannot be compiled in this image (no rustc / cargo), so these tests pin
what can be pinned without a compiler: the `extern "C"` block is regenerated from include/fb2_w4a16.h
and must equal the committed one; every symbol the library exports is declared; every `ffi::` call in
the safe wrapper, and every wrapper call in the ferrum-kernels drop-in file, names a declared function
with the declared number of arguments; the sources are bracket-balanced."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FFI_RS = os.path.join(ROOT, "rust", "fb2-w4a16-sys", "src", "ffi.rs")
LIB_RS = os.path.join(ROOT, "rust", "fb2-w4a16-sys", "src", "lib.rs")
FERRUM_RS = os.path.join(ROOT, "rust", "ferrum-kernels", "src", "quant_linear", "fb2_w4a16.rs")
EXAMPLE_RS = os.path.join(ROOT, "rust", "fb2-w4a16-sys", "examples", "cuda_vs_cpu.rs")
HEADER = os.path.join(ROOT, "include", "fb2_w4a16.h")
SO = os.path.join(ROOT, "ferrum-infer-rs_b200", "libfb2_w4a16.so")


def _strip(src: str) -> str:
    """Drop comments, string and char literals so brackets / commas inside them do not count."""
    src = re.sub(r"//[^\n]*", "", src)
    src = re.sub(r'"(?:\\.|[^"\\])*"', '""', src)
    return re.sub(r"'(?:\\.|[^'\\])'", "' '", src)


def _args_at(src: str, open_paren: int):
    """Top-level comma-separated pieces of the parenthesised list that starts at `open_paren`."""
    depth, start, parts = 0, open_paren + 1, []
    pairs = {"(": ")", "[": "]", "{": "}"}
    for i in range(open_paren, len(src)):
        c = src[i]
        if c in pairs:
            depth += 1
        elif c in pairs.values():
            depth -= 1
            if depth == 0:
                parts.append(src[start:i])
                break
        elif c == "," and depth == 1:
            parts.append(src[start:i])
            start = i + 1
    return [p.strip() for p in parts if p.strip()]


def _ffi_decls():
    src = open(FFI_RS).read()
    block = src[src.index('extern "C" {'):]
    out = {}
    for m in re.finditer(r"pub fn (fb2_\w+)\(", block):
        out[m.group(1)] = len(_args_at(block, m.end() - 1))
    return out


def test_ffi_rs_is_what_the_header_generates():
    gen = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gen_rust_ffi.py")], capture_output=True,
                         text=True, check=True).stdout
    assert gen == open(FFI_RS).read(), "rust/fb2-w4a16-sys/src/ffi.rs is stale: run scripts/gen_rust_ffi.py --write"


def test_every_header_symbol_is_declared_in_rust():
    header = re.sub(r"/\*.*?\*/", " ", open(HEADER).read(), flags=re.S)
    names = set(re.findall(r"FB2_API\s+[^;(]*?\b(fb2_\w+)\s*\(", header))
    decls = _ffi_decls()
    assert names and names == set(decls), (names ^ set(decls))


def test_rust_declarations_resolve_in_the_library():
    if not os.path.exists(SO):
        pytest.skip("library not built")
    lib = ctypes.CDLL(SO)
    missing = [n for n in _ffi_decls() if not hasattr(lib, n)]
    assert not missing, missing


def test_wrapper_calls_match_the_declarations():
    decls = _ffi_decls()
    src = _strip(open(LIB_RS).read())
    calls = list(re.finditer(r"ffi::(fb2_\w+)\(", src))
    assert len(calls) > 40
    for m in calls:
        name = m.group(1)
        assert name in decls, f"lib.rs calls undeclared {name}"
        got = len(_args_at(src, m.end() - 1))
        assert got == decls[name], f"lib.rs calls {name} with {got} arguments, header declares {decls[name]}"
    consts = set(re.findall(r"pub const (FB2_\w+)", open(FFI_RS).read()))
    for c in re.findall(r"ffi::(FB2_\w+)", src):
        assert c in consts, c


def _wrapper_fns():
    """name -> number of parameters besides self, for every pub fn of lib.rs."""
    src = _strip(open(LIB_RS).read())
    out = {}
    for m in re.finditer(r"pub (?:unsafe )?fn (\w+)(?:<[^>]*>)?\(", src):
        params = [p for p in _args_at(src, m.end() - 1) if not re.match(r"(&(mut )?)?self\b|self\s*:", p)]
        out.setdefault(m.group(1), set()).add(len(params))
    return out


def test_ferrum_drop_in_calls_match_the_wrapper():
    fns = _wrapper_fns()
    src = _strip(open(FERRUM_RS).read())
    checked = 0
    for m in re.finditer(r"(?:self\.inner\.|self\.comm\.|\.handle\(\)\.|fb2::(?:\w+::)?)(\w+)\(", src):
        name = m.group(1)
        if name in ("new", "clone") or name[0].isupper():      # constructors of this file's types, enum patterns
            continue
        assert name in fns, f"fb2_w4a16.rs calls {name}, which rust/fb2-w4a16-sys/src/lib.rs does not define"
        got = len(_args_at(src, m.end() - 1))
        assert got in fns[name], f"fb2_w4a16.rs calls {name} with {got} arguments, lib.rs takes {sorted(fns[name])}"
        checked += 1
    assert checked >= 25
    for variant in re.findall(r"fb2::Fb2Error::(\w+)", src):
        assert re.search(rf"\b{variant}\(String\)", open(LIB_RS).read()), variant


@pytest.mark.parametrize("path", [FFI_RS, LIB_RS, FERRUM_RS, EXAMPLE_RS])
def test_rust_sources_are_bracket_balanced(path):
    src = _strip(open(path).read())
    stack, pairs = [], {")": "(", "]": "[", "}": "{"}
    for c in src:
        if c in "([{":
            stack.append(c)
        elif c in pairs:
            assert stack and stack.pop() == pairs[c], f"unbalanced '{c}' in {path}"
    assert not stack


def test_integration_md_excerpts_match_the_declarations():
    """The `extern "C"` excerpts in INTEGRATION.md are for reading; they must still name declared functions with the
    declared number of arguments."""
    decls = _ffi_decls()
    src = _strip(open(os.path.join(ROOT, "INTEGRATION.md")).read())
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    seen = 0
    for m in re.finditer(r"pub fn (fb2_\w+)\(", src):
        assert m.group(1) in decls, m.group(1)
        assert len(_args_at(src, m.end() - 1)) == decls[m.group(1)], m.group(1)
        seen += 1
    assert seen >= 20


def test_example_calls_match_the_wrapper():
    """rust/fb2-w4a16-sys/examples/cuda_vs_cpu.rs (the reference's cuda_vs_cpu fixture driven from Rust)."""
    fns = _wrapper_fns()
    src = _strip(open(EXAMPLE_RS).read())
    for name, call in (("load_gptq", "GptqLinear::load_gptq("), ("forward_host", "lin.forward_host(")):
        i = src.index(call) + len(call) - 1
        assert len(_args_at(src, i)) in fns[name], name
