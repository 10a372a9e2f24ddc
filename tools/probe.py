"""Prints the integer-pipe probe rates (G thread-instr/s) measured on this GPU."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hb = importlib.import_module("mpc-protocols_b200")
ctx = hb.Context(0)
out = {}
for v, name in [(0, "mad.lo.u32"), (1, "mad.wide.u32"), (2, "chain4 (IMAD.WIDE.U32.X)")]:
    g, ms = ctx.measure_imad_peak(v)
    out[name] = {"ginst_per_s": g, "ms": ms}
print(json.dumps(out))
