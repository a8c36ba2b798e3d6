"""Prints the headline numbers of a bench.py line read from stdin (last JSON line)."""
import json
import sys
j = json.loads([l for l in sys.stdin if l.startswith("{")][-1])
print(sys.argv[1] if len(sys.argv) > 1 else "", "value", round(j["value"], 4), "ms/step", round(j["ms_per_step"], 2), "e2e", round(j["e2e"]["value"], 4),
      {k: round(v, 2) for k, v in j["stage_ms_per_step"].items()})
for k in ("config4", "config4_fp16", "config4_fp16x3", "config5"):
    c = j.get(k)
    if c:
        print(" ", k, {x: (round(c[x], 2) if isinstance(c.get(x), float) else c.get(x)) for x in ("value", "ms_per_step", "ms", "fit_ms_rank0", "score_ms_rank0", "tflops_total", "error", "skipped") if x in c})
if j.get("fit"):
    print("  fit", j["fit"])
