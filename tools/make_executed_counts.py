#!/usr/bin/env python
"""Builds profiles/r02_executed_counts.json -- the measured executed-instruction counts bench.py's roofline reads -- from the per-kernel
summaries tools/gpu_ncu_counts.sh leaves in gpurun_out/ (exec_ntt_TAG.json, exec_dense_TAG.json, exec_k4_TAG.json), and copies those
summaries to profiles/ as the evidence.   python tools/make_executed_counts.py <tag>"""
import json, os, re, shutil, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

def load(name):
    p = os.path.join(G, f"exec_{name}_{tag}.json")
    return json.load(open(p)) if os.path.exists(p) else []

def entry(e, src, what):
    c = e["classes"]
    return {"kernel": e["kernel"], "what": what, "items": e["items"], "imad_wide_thread_inst_per_item": e["imad_wide_thread_inst_per_item"],
            "thread_inst_per_item": e["thread_inst_per_item"], "imad_other_thread_inst_per_item": c.get("imad_other", {}).get("thread_inst", 0) / e["items"],
            "alu_thread_inst_per_item": c.get("alu", {}).get("thread_inst", 0) / e["items"], "launches": e.get("launches", 1), "source": src}

out = {}
for e in load("ntt"):
    m = re.search(r"ntt\w*_kernel<\(int\)(\d+), \(int\)(\d+)>", e["kernel"])
    if m and m.group(1) == "6":
        key = {"0": "gen", "1": "recon"}.get(m.group(2))
        if key and key not in out:
            out[key] = entry(e, f"profiles/{tag}_exec_ntt.json", "one launch at 2^20 items, n=64, t=21 (tools/ncu_ntt.py), ncu SourceCounters: sum of 'Thread Instructions Executed' over the IMAD.WIDE* SASS lines")
dense = [e for e in load("dense") if "matvec_kernel" in e["kernel"]]
if dense:
    e = max(dense, key=lambda e: e["imad_wide_thread_inst_per_item"])
    out["dense"] = entry(e, f"profiles/{tag}_exec_dense.json", "one launch at 2^20 chunks, 43 senders with flags (tools/ncu_ntt.py --what dense)")
k4 = load("k4")
if k4:
    out["k4"] = entry(k4[0], f"profiles/{tag}_exec_k4.json", "ALL kernels of one robust_interpolate_batch call, n=128, t=42, 2^17 codewords, e~U{0..42} (tools/ncu_ntt.py --what k4); per-kernel split in the source file")
for name in ("ntt", "dense", "k4"):
    src = os.path.join(G, f"exec_{name}_{tag}.json")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, f"{tag}_exec_{name}.json"))
json.dump(out, open(os.path.join(P, "r02_executed_counts.json"), "w"), indent=1)
print(json.dumps({k: round(v["imad_wide_thread_inst_per_item"], 1) for k, v in out.items()}))
