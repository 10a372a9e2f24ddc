#!/usr/bin/env python
"""Dynamic instruction counts per kernel from an `ncu --page source --csv` export (needs --import-source on and -lineinfo):
executed warp- and thread-level instructions per issue-pipe class (IMAD.WIDE / other IMAD / ALU / shared / global / control) and
the stall samples attributed to each class.  With --items N the thread-level IMAD.WIDE count is also given per item.
   python tools/ncu_exec_counts.py gpurun_out/src_TAG.csv [--items 1048576] [--json out.json] [--top 12]"""
import argparse, collections, csv, json, re, sys
sys.path.insert(0, __import__("os").path.dirname(__file__))
from sass_hist import classify

def parse(path):
    kernels, cur, hdr = [], None, None
    for row in csv.reader(open(path)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": []}
            kernels.append(cur)
            hdr = None
            continue
        if row[0] == "Address":
            hdr = row
            continue
        if cur is not None and hdr is not None and len(row) >= len(hdr) - 1:
            cur["rows"].append(dict(zip(hdr, row)))
    # `ncu --page source --csv` prints every launch twice (one section per view); keep one of each identical pair
    def sig(k):
        return (k["name"], len(k["rows"]), sum(int(r["Instructions Executed"] or 0) for r in k["rows"]))
    if len(kernels) % 2 == 0 and all(sig(kernels[i]) == sig(kernels[i + 1]) for i in range(0, len(kernels), 2)):
        kernels = kernels[::2]
    return kernels

def summarise(k, items=None, top=12):
    cls = collections.defaultdict(lambda: {"warp_inst": 0, "thread_inst": 0, "stall_samples": 0})
    stall_kinds = collections.Counter()
    rows = []
    for r in k["rows"]:
        src = r["Source"].strip()
        m = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
        if not m:
            continue
        op = m.group(1)
        c = classify(op)
        wi, ti, ss = int(r["Instructions Executed"] or 0), int(r["Thread Instructions Executed"] or 0), int(r["Warp Stall Sampling (All Samples)"] or 0)
        cls[c]["warp_inst"] += wi; cls[c]["thread_inst"] += ti; cls[c]["stall_samples"] += ss
        for key, v in r.items():
            if key.startswith("stall_") and "Not Issued" not in key and v not in ("", "0"):
                stall_kinds[key] += int(v)
        rows.append((ss, wi, src))
    tot_w = sum(v["warp_inst"] for v in cls.values())
    tot_s = sum(v["stall_samples"] for v in cls.values())
    out = {"kernel": k["name"], "warp_inst_total": tot_w, "classes": {c: dict(v) for c, v in cls.items()},
           "fma_pipe_warp_cycles": 4 * cls["imad_wide"]["warp_inst"] + 2 * cls["imad_other"]["warp_inst"] + 2 * cls["fp"]["warp_inst"],
           "alu_pipe_warp_cycles": 2 * cls["alu"]["warp_inst"], "stall_samples_total": tot_s,
           "stall_kinds": dict(stall_kinds.most_common(8)),
           "top_stall_instructions": [{"samples": s, "warp_inst": w, "sass": src} for s, w, src in sorted(rows, reverse=True)[:top]]}
    if items:
        out["items"] = items
        out["imad_wide_thread_inst_per_item"] = cls["imad_wide"]["thread_inst"] / items
        out["thread_inst_per_item"] = sum(v["thread_inst"] for v in cls.values()) / items
    return out

def aggregate(res, items):
    """per kernel name: launches and summed class counts; first entry: the grand total of the whole capture"""
    by = collections.OrderedDict()
    for r in res:
        name = re.sub(r"\(.*", "", r["kernel"])
        e = by.setdefault(name, {"kernel": name, "launches": 0, "warp_inst_total": 0, "classes": collections.defaultdict(lambda: collections.Counter()),
                                 "fma_pipe_warp_cycles": 0, "alu_pipe_warp_cycles": 0, "stall_samples_total": 0, "stall_kinds": collections.Counter(), "top_stall_instructions": []})
        e["launches"] += 1
        for k in ("warp_inst_total", "fma_pipe_warp_cycles", "alu_pipe_warp_cycles", "stall_samples_total"):
            e[k] += r[k]
        for c, v in r["classes"].items():
            e["classes"][c].update(v)
        e["stall_kinds"].update(r["stall_kinds"])
    out = []
    tot = {"kernel": "TOTAL (all kernels of the captured call)", "launches": 0, "warp_inst_total": 0, "classes": collections.defaultdict(lambda: collections.Counter()),
           "fma_pipe_warp_cycles": 0, "alu_pipe_warp_cycles": 0, "stall_samples_total": 0, "stall_kinds": collections.Counter(), "top_stall_instructions": []}
    for e in by.values():
        for k in ("launches", "warp_inst_total", "fma_pipe_warp_cycles", "alu_pipe_warp_cycles", "stall_samples_total"):
            tot[k] += e[k]
        for c, v in e["classes"].items():
            tot["classes"][c].update(v)
        tot["stall_kinds"].update(e["stall_kinds"])
    for e in [tot] + sorted(by.values(), key=lambda e: -e["fma_pipe_warp_cycles"]):
        e["classes"] = {c: dict(v) for c, v in e["classes"].items()}
        e["stall_kinds"] = dict(e["stall_kinds"].most_common(8))
        for c in ("imad_wide", "imad_other", "alu", "fp"):
            e["classes"].setdefault(c, {"warp_inst": 0, "thread_inst": 0, "stall_samples": 0})
        if items:
            e["items"] = items
            e["imad_wide_thread_inst_per_item"] = e["classes"]["imad_wide"]["thread_inst"] / items
            e["thread_inst_per_item"] = sum(v["thread_inst"] for v in e["classes"].values()) / items
        out.append(e)
    return out

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--items", type=int)
    ap.add_argument("--json")
    ap.add_argument("--top", type=int, default=12)
    ap.add_argument("--aggregate", action="store_true", help="sum all launches per kernel name and add a grand total (a multi-kernel call, e.g. K4)")
    a = ap.parse_args()
    res = [summarise(k, a.items, a.top) for k in parse(a.csv)]
    if a.aggregate:
        res = aggregate(res, a.items)
    if a.json:
        json.dump(res, open(a.json, "w"), indent=1)
    for r in res:
        print("==", r["kernel"], ("x%d" % r["launches"]) if "launches" in r else "", "warp inst", r["warp_inst_total"], "| WIDE/item", r.get("imad_wide_thread_inst_per_item"), "| inst/item", r.get("thread_inst_per_item"))
        for c, v in sorted(r["classes"].items(), key=lambda kv: -kv[1]["warp_inst"]):
            print(f"   {c:11s} warp_inst {v['warp_inst']:>12d} ({100*v['warp_inst']/max(r['warp_inst_total'],1):5.1f} %)  stall samples {v['stall_samples']:>8d} ({100*v['stall_samples']/max(r['stall_samples_total'],1):5.1f} %)")
        print("   fma pipe warp-cycles", r["fma_pipe_warp_cycles"], " alu pipe warp-cycles", r["alu_pipe_warp_cycles"], " stalls:", r["stall_kinds"])
        for t in r["top_stall_instructions"]:
            print(f"      {t['samples']:>7d} samples  {t['warp_inst']:>10d} x  {t['sass']}")

if __name__ == "__main__":
    main()
