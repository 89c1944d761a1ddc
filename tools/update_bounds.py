"""Rewrite the per-tap bound tables of tests/helpers.py from a measurement of tools/measure_taps.py:  bound = 1.5 x measured.

    python tools/update_bounds.py gpurun_out/taps.json
"""
import json
import re
import sys

path = "tests/helpers.py"
m = json.load(open(sys.argv[1]))
src = open(path).read()


def fix(table, key, val, txt):
    pat = re.compile(r'(    "%s": )([0-9.e+-]+)(,\s*# measured )([0-9.e+-]+)' % re.escape(key))
    new, n = pat.subn(lambda g: "%s%.2e%s%.2e" % (g.group(1), 1.5 * val, g.group(3), val), txt)
    assert n == 1, (table, key, n)
    return new


for k, v in m["taps"].items():
    src = fix("TAP_BOUNDS", k, v, src)
for k, v in m["scc_parts"].items():
    src = fix("SCC_PART_BOUNDS", k, v, src)
open(path, "w").write(src)
print("updated", len(m["taps"]) + len(m["scc_parts"]), "bounds")
