"""Print the headline numbers of bench.py JSON lines read from stdin (development aid)."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l)
        r = d.get("roofline", {})
        print(tag, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "single_lane", round(d.get("single_lane", {}).get("value", 0), 1),
              "hot_ms", round(r.get("ms_per_launch", 0), 5), "frac", round(r.get("frac", 0), 4), r.get("segments_ms_per_eval"))
