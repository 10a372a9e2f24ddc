"""Prints the integer-pipe probe rates (G thread-instr/s) measured on this GPU: the three peak probes of hbmpc_measure_imad_peak and
the chain-latency table of hbmpc_measure_wide_chains (independent carry chains per thread x resident warps per sub-partition)."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hb = importlib.import_module("mpc-protocols_b200")
ctx = hb.Context(0)
out = {}
for v, name in [(0, "mad.lo.u32 (IMAD)"), (1, "IMAD.WIDE.U32.X carry chains (chain4w)"), (2, "DFMA")]:
    g, ms = ctx.measure_imad_peak(v)
    out[name] = {"ginst_per_s": g, "ms": ms}
tab = {}
for ch in (1, 2, 4, 8):
    tab[f"chains={ch}"] = {f"warps_per_smsp={w}": round(ctx.measure_wide_chains(ch, w), 1) for w in (1, 2, 3, 4, 6, 8, 12, 16)}
out["wide_chain_latency_table_ginst_per_s"] = tab
mm = {}
for ilp in (1, 2, 4):
    mm[f"ilp={ilp}"] = {f"warps_per_smsp={w}": round(ctx.measure_mont_mul(ilp, w), 2) for w in (1, 2, 4, 6, 8, 12, 16)}
out["mont_mul_gproducts_per_s"] = mm
print(json.dumps(out, indent=1))
