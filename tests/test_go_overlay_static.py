"""The cgo overlay (go/align/cuda_bridge.go) cannot be compiled here (no Go toolchain), so this is a
STATIC check of the drop-in boundary: every unexported method the reference's generated files declare
(tests/golden/go_align_methods.txt, regenerated from /root/reference by tests/golden/make_go_methods.py)
exists in the overlay with the same receiver, name and signature; AlignBatch exists on all six
types; every C symbol the overlay calls is declared in include/biogo_align.h and exported by the
library; brackets balance."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "go", "align", "cuda_bridge.go")
TYPES = ("SW", "NW", "Fitted", "SWAffine", "NWAffine", "FittedAffine")


def _go():
    return open(GO).read()


def _strip(src):
    """drop comments, the cgo preamble and string/rune literals (for bracket counting)"""
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    src = re.sub(r'"(\\.|[^"\\])*"', '""', src)
    src = re.sub(r"'(\\.|[^'\\])*'", "''", src)
    return src


def test_every_reference_method_is_replaced_with_the_same_signature():
    want = [ln.rstrip("\n").split("\t") for ln in open(os.path.join(ROOT, "tests", "golden", "go_align_methods.txt"))
            if ln.strip()]
    assert len(want) == 12
    src = _go()
    for f, recv, name, params, results in want:
        pat = r"^func \(a %s\) %s\(%s\) \(%s\) \{" % (re.escape(recv), re.escape(name), re.escape(params),
                                                      re.escape(results))
        assert re.search(pat, src, flags=re.M), f"{f}: func (a {recv}) {name}({params}) ({results}) missing"
    got = re.findall(r"^func \(a (\w+)\) (align\w+)\(", src, flags=re.M)
    assert sorted(got) == sorted((w[1], w[2]) for w in want)   # nothing extra, nothing twice


def test_reference_tree_agrees_with_the_fixture():
    ref = "/root/reference/align"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present on this box")
    import glob

    found = []
    for f in sorted(glob.glob(ref + "/*_letters.go") + glob.glob(ref + "/*_qletters.go")):
        for ln in open(f):
            m = re.match(r"^func \(a (\w+)\) (\w+)\((.*)\) \((.*)\) \{", ln)
            if m:
                found.append("\t".join((os.path.relpath(f, "/root/reference"),) + m.groups()))
    have = [ln.rstrip("\n") for ln in open(os.path.join(ROOT, "tests", "golden", "go_align_methods.txt")) if ln.strip()]
    assert found == have


def test_align_batch_on_all_six_types():
    src = _go()
    for t in TYPES:
        assert re.search(r"^func \(a %s\) AlignBatch\(refs, queries \[\]AlphabetSlicer\) \(\[\]\[\]feat\.Pair, \[\]error\) \{"
                         % t, src, flags=re.M), t
    # Fitted kinds skip the matrix size check (align/fitted_letters.go:71-78), the others keep it
    for t, kind, check in (("SW", "kindSW", "true"), ("NW", "kindNW", "true"), ("Fitted", "kindFitted", "false")):
        assert re.search(r"func \(a %s\) alignLetters.*\n\treturn scoring\{%s, false, Linear\(a\), 0, %s\}" % (t, kind, check), src)
        assert re.search(r"func \(a %sAffine\) alignLetters.*\n\treturn scoring\{%s, true, a\.Matrix, a\.GapOpen, %s\}"
                         % (t, kind, check), src)


def test_c_symbols_used_by_the_overlay_are_declared_and_exported():
    import ctypes as C

    src = _go()
    hdr = open(os.path.join(ROOT, "include", "biogo_align.h")).read()
    used = set(re.findall(r"\bC\.(ba_\w+)\(", src))
    assert {"ba_create", "ba_destroy", "ba_set_scoring", "ba_align_batch", "ba_free_result", "ba_alloc_host",
            "ba_free_host", "ba_create_multi", "ba_align_batch_multi", "ba_format_batch"} <= used
    from biogo_b200 import _lib

    lib = C.CDLL(_lib.DEFAULT_LIB) if os.path.exists(_lib.DEFAULT_LIB) else None
    for sym in sorted(used):
        assert re.search(r"\b%s\(" % sym, hdr), f"{sym} is not declared in biogo_align.h"
        if lib is not None:
            assert hasattr(lib, sym), sym
    for const in set(re.findall(r"\bC\.(BA_\w+)\b", src)):
        assert re.search(r"\b%s\b" % const, hdr), const
    for typ in set(re.findall(r"\bC\.(ba_\w+)\b(?!\()", src)) - used:
        assert re.search(r"\b%s\b" % typ, hdr), typ


def test_contexts_are_destroyed_explicitly_and_packing_is_pinned():
    src = _go()
    assert "sync.Pool" not in _strip(src)                    # the GC must never drop a ba_ctx silently
    assert "C.ba_destroy(" in src and "C.ba_destroy_multi(" in src and "func Shutdown()" in src
    assert "C.ba_alloc_host(" in src and "C.ba_free_host(" in src
    # a failed ba_create is reported, and ba_last_error is never called on a nil context
    m = re.search(r"if rc := C\.ba_create\(0, &c\); rc != C\.BA_OK \{(.*?)\n\t\}", src, flags=re.S)
    assert m and "if c != nil" in m.group(1)


def test_brackets_balance():
    s = _strip(_go())
    for a, b in ("()", "[]", "{}"):
        assert s.count(a) == s.count(b), (a, s.count(a), s.count(b))
    depth = 0
    for ch in s:
        depth += ch == "{"
        depth -= ch == "}"
        assert depth >= 0
