"""The Go provider ships as source (no Go toolchain in this image).  What CAN be checked here: every C name the cgo files use is
declared in include/fabgpu_ecdsa.h with the arity the call site passes, braces and parentheses balance, every file declares the
package its directory implies, and the patch for the reference's files touches files that exist in a reference checkout layout."""
import os
import re

from util import ROOT

GO = os.path.join(ROOT, "fabric-mod_b200", "go")


def _go_files():
    for d, _, fs in os.walk(GO):
        for f in fs:
            if f.endswith(".go"):
                yield os.path.join(d, f)


def _strip(src):
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    src = re.sub(r'"(\\.|[^"\\])*"', '""', src)
    src = re.sub(r"`[^`]*`", '""', src)
    return re.sub(r"'(\\.|[^'\\])'", "' '", src)


def _header_decls():
    txt = open(os.path.join(ROOT, "include", "fabgpu_ecdsa.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(fabgpu_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        args = [a for a in m.group(2).split(",") if a.strip() and a.strip() != "void"]
        decls[m.group(1)] = len(args)
    macros = set(re.findall(r"#define\s+(FABGPU_[A-Z0-9_]+)", txt)) | set(re.findall(r"\b(FABGPU_[A-Z0-9_]+)\s*=", txt))
    return decls, macros


def test_go_files_are_well_formed_and_in_their_packages():
    files = list(_go_files())
    assert len(files) >= 8
    for path in files:
        raw = open(path).read()
        body = _strip(raw.split("import \"C\"")[-1]) if "import \"C\"" in raw else _strip(raw)
        for a, b in ("{}", "()", "[]"):
            assert body.count(a) == body.count(b), (path, a)
        pkg = re.search(r"^package\s+(\w+)", _strip(raw), flags=re.M).group(1)
        assert pkg == os.path.basename(os.path.dirname(path)), path


def test_cgo_calls_match_the_header():
    decls, macros = _header_decls()
    used = 0
    for path in _go_files():
        src = _strip(open(path).read().split("import \"C\"")[-1])
        for m in re.finditer(r"\bC\.(fabgpu_[a-z0-9_]+)\s*\(", src):
            name = m.group(1)
            assert name in decls, (path, name)
            # count the call's top-level arguments
            i, depth, n_args, seen = m.end(), 1, 0, False
            while depth:
                c = src[i]
                if c in "([{":
                    depth += 1
                elif c in ")]}":
                    depth -= 1
                elif c == "," and depth == 1:
                    n_args += 1
                if not c.isspace() and depth:
                    seen = True
                i += 1
            n_args = n_args + 1 if seen else 0
            assert n_args == decls[name], (path, name, n_args, decls[name])
            used += 1
        for m in re.finditer(r"\bC\.(FABGPU_[A-Z0-9_]+)", src):
            assert m.group(1) in macros, (path, m.group(1))
    assert used >= 10


def test_patch_targets_reference_layout():
    patch = open(os.path.join(GO, "patches", "fabric-gpu-bccsp.patch")).read()
    targets = re.findall(r"^\+\+\+ b/(\S+)", patch, flags=re.M)
    assert targets, "the patch names no files"
    for t in targets:
        assert t.startswith(("bccsp/", "extensions/", "core/", "internal/", "sampleconfig/")), t
