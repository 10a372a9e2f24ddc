#!/usr/bin/env python
"""Static SASS instruction histogram per kernel of libhbmpc_b200.so (cuobjdump -sass), grouped by issue pipe.

usage: python tools/sass_hist.py [--lib path] [--kernels regex] [--json out.json]
fma-pipe classes: IMAD.WIDE* (32x32->64 multiply-add, half rate), other IMAD* (IMAD.X / IMAD.MOV / IMAD.IADD / IMAD.SHL ...), FP.
alu-pipe classes: IADD3*, LOP3, SHF, SEL, ISETP, PRMT, MOV, ...
"""
import argparse, collections, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def demangle(names):
    try:
        out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}

def classify(op):
    if op.startswith("IMAD.WIDE"): return "imad_wide"
    if op.startswith("IMAD"): return "imad_other"
    if op.startswith(("IADD3", "IADD", "LOP3", "SHF", "SEL", "ISETP", "PRMT", "MOV", "LEA", "IMNMX", "VIMNMX", "PLOP3", "POPC", "FLO", "BREV", "VIADD", "R2P", "P2R", "CS2R", "S2R", "IABS")): return "alu"
    if op.startswith(("LDS", "STS", "LDSM")): return "smem"
    if op.startswith(("LDG", "STG", "LDGSTS", "LD.", "ST.", "ATOM", "RED", "LDC", "ULDC", "LDL", "STL")): return "mem"
    if op.startswith(("BAR", "WARPSYNC", "DEPBAR", "BSSY", "BSYNC", "BRA", "EXIT", "CALL", "RET", "NOP", "YIELD", "ERRBAR", "MEMBAR", "LDGDEPBAR", "SHFL", "VOTE", "MATCH")): return "ctl"
    if op.startswith(("DFMA", "DADD", "DMUL", "FFMA", "FMUL", "FADD", "MUFU", "I2F", "F2I")): return "fp"
    return "other"

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "mpc-protocols_b200", "libhbmpc_b200.so"))
    ap.add_argument("--kernels", default=".")
    ap.add_argument("--json")
    a = ap.parse_args()
    sass = subprocess.run(["cuobjdump", "-sass", a.lib], capture_output=True, text=True).stdout
    cur, hist, ops = None, {}, {}
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1); hist[cur] = collections.Counter(); ops[cur] = collections.Counter(); continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and cur:
            op = m.group(1)
            hist[cur][classify(op)] += 1
            ops[cur][op.split(".")[0] + ("." + op.split(".")[1] if op.startswith("IMAD") and "." in op else "")] += 1
    names = demangle(list(hist))
    out = {}
    for k, h in hist.items():
        nm = names[k]
        if not re.search(a.kernels, nm): continue
        tot = sum(h.values())
        out[nm] = {"total": tot, **dict(h), "imad_detail": {o: c for o, c in ops[k].items() if o.startswith("IMAD")},
                   "fma_pipe_cycles_static": 4 * h["imad_wide"] + 2 * h["imad_other"], "alu_pipe_cycles_static": 2 * h["alu"]}
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)
    for nm, h in out.items():
        print(nm.split("(")[0], {k: v for k, v in h.items() if k != "imad_detail"}, h["imad_detail"])

if __name__ == "__main__":
    main()
