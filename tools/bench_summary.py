"""Print the headline numbers and the per-category breakdown of a bench.py JSON line (last line starting with '{' of a log)."""
import json
import sys


def main():
    for path in sys.argv[1:]:
        lines = [x for x in open(path) if x.startswith("{")]
        if not lines:
            print(path, "no JSON line")
            continue
        d = json.loads(lines[-1])
        if "ms_per_step" not in d:
            print(path, d)
            continue
        r = d.get("roofline", {})
        print(f"{path}: {d['ms_per_step']:.2f} ms/step  {d['value']:.1f} {d['unit']}  e2e {d.get('e2e', {}).get('value')}  "
              f"clk {d.get('clocks', {}).get('sm_mhz')}  roofline {r.get('kernel')} frac {r.get('frac')}")
        for k, v in list(d.get("breakdown", {}).items())[:int(14)]:
            print(f"   {k:22s} {v['ms_per_step']:8.3f} ms  x{v['launches_per_step']}")


if __name__ == "__main__":
    try:
        main()
    except BrokenPipeError:
        pass
