"""rust/hbmpc-sys (text only: no Rust toolchain here) must declare every entry point include/hbmpc_b200.h declares, with the same
number of arguments -- so the crate text cannot drift from the header unnoticed."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _decls(text, pattern):
    out = {}
    for m in re.finditer(pattern, text, flags=re.S):
        name, args = m.group(1), m.group(2)
        args = args.strip()
        out[name] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_rust_bindings_cover_the_header():
    hdr = open(os.path.join(ROOT, "include", "hbmpc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    rs = open(os.path.join(ROOT, "rust", "hbmpc-sys", "src", "lib.rs")).read()
    rs = re.sub(r"//.*", "", rs)
    c = _decls(hdr, r"\b(hbmpc_\w+)\s*\(([^;{]*?)\)\s*;")
    r = _decls(rs, r"pub fn (hbmpc_\w+)\s*\(([^;{]*?)\)\s*(?:->[^;]*)?;")
    assert c and set(c) == set(r), (sorted(set(c) - set(r)), sorted(set(r) - set(c)))
    assert {k: v for k, v in c.items() if r[k] != v} == {}
