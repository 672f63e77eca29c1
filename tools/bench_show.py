"""Print the headline numbers and the per-kernel table of a bench.py JSON line (file argument)."""
import json, sys
d = json.loads(open(sys.argv[1]).read())
print("value %.1f img/s  %.3f ms/step  e2e %.1f  clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("clocks")))
pk = d.get("per_kernel", {})
rows = sorted(((k, v) for k, v in pk.items() if isinstance(v, dict)), key=lambda kv: -kv[1]["ms_per_step"])
tot = sum(v["ms_per_step"] for _, v in rows)
for k, v in rows:
    if v["ms_per_step"] >= 0.05:
        print("  %7.3f ms %6.1f launches %5.1f%%  %s" % (v["ms_per_step"], v["launches_per_step"], 100 * v["ms_per_step"] / tot, k))
print("  sum %.3f, pass %.3f ms/step" % (tot, pk.get("_pass_ms_per_step", 0)))
