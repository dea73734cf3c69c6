"""Static check of the ctypes boundary (no GPU): every call the Python host code makes into libcbops.so passes as many
arguments as include/cbops.h declares for that entry point.  ctypes does not check this (the functions have no argtypes: raw
pointers, sizes and a stream cross the boundary), and a missing argument would only show up on a GPU box as garbage."""
import ast
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "contrastboundary_b200")


def declared_arity():
    text = open(os.path.join(ROOT, "include", "cbops.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|float|size_t|void|unsigned long long|const char \*)\s*\*?\s*(cb_\w+|\w+_cuda_launcher)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        name, params = m.group(1), m.group(2).strip()
        out[name] = 0 if params in ("", "void") else params.count(",") + 1
    return out


def call_sites():
    """(file, line, function name, number of positional arguments) of L.call("cb_x", ...) and <lib>.cb_x(...) calls"""
    sites = []
    files = [os.path.join(PKG, f) for f in sorted(os.listdir(PKG))] + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    for d in ("tests", "tools"):
        files += [os.path.join(ROOT, d, f) for f in sorted(os.listdir(os.path.join(ROOT, d)))]
    for path in files:
        if not path.endswith(".py"):
            continue
        tree = ast.parse(open(path).read(), path)
        for node in ast.walk(tree):
            if not isinstance(node, ast.Call) or any(isinstance(a, ast.Starred) for a in node.args):
                continue
            f = node.func
            if isinstance(f, ast.Attribute) and f.attr == "call" and node.args and isinstance(node.args[0], ast.Constant) \
                    and isinstance(node.args[0].value, str) and node.args[0].value.startswith("cb_"):
                sites.append((path, node.lineno, node.args[0].value, len(node.args) - 1))
            elif isinstance(f, ast.Attribute) and f.attr.startswith("cb_") and not node.keywords:
                sites.append((path, node.lineno, f.attr, len(node.args)))
    return sites


def test_header_declares_what_python_calls_with_matching_arity():
    decl = declared_arity()
    assert len(decl) > 60, len(decl)
    sites = call_sites()
    assert len(sites) > 60, len(sites)
    bad = []
    for path, line, name, nargs in sites:
        if name not in decl:
            bad.append(f"{os.path.relpath(path, ROOT)}:{line}: {name} is not declared in include/cbops.h")
        elif decl[name] != nargs:
            bad.append(f"{os.path.relpath(path, ROOT)}:{line}: {name} called with {nargs} arguments, declared with {decl[name]}")
    assert not bad, "\n".join(bad)
