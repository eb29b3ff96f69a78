"""Print value / e2e / roofline of bench.py JSON lines:  python tools/show_bench.py file.json [...]"""
import json
import sys

for p in sys.argv[1:]:
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(p, "unreadable:", e)
        continue
    e2e = d.get("e2e") or {}
    roof = d.get("roofline") or {}
    cpu = d.get("cpu_baseline") or {}
    print(f"{p}: value {d.get('value', 0):.1f} {d.get('unit')}  e2e {e2e.get('value', 0):.1f}  "
          f"roofline {roof.get('frac')}  cpu {cpu.get('value')}  clocks {d.get('clocks')}")
