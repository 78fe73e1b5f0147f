#!/usr/bin/env python
"""Developer tool: print the key fields of bench.py JSON lines found on stdin (ignores other output)."""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    try:
        d = json.loads(line)
    except Exception:
        continue
    r = d.get("roofline", {})
    print(tag, "gpus", d.get("n_gpus"), "ms/step %.4f" % d.get("ms_per_step", -1), "GF/s %.0f" % d.get("value", -1),
          "pass", {k: round(v, 4) for k, v in d.get("pass_ms", {}).items()},
          "stage", {k: round(v, 4) for k, v in d.get("stage_ms", {}).items()},
          "roof %.3f" % r.get("frac", -1), "e2e", round((d.get("e2e") or {}).get("value", -1)), "e2e_serial", round((d.get("e2e") or {}).get("serial_value", -1)), d.get("config", {}).get("t0"), d.get("config", {}).get("exchange"),
          d.get("clocks", {}).get("reasons"))
