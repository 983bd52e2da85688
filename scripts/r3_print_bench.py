"""Headline fields of a bench.py JSON line: python scripts/r3_print_bench.py <file>."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])


def find(x, k, path=""):
    if isinstance(x, dict):
        for kk, v in x.items():
            if kk == k:
                print(path + "/" + kk, {a: b for a, b in v.items() if not isinstance(b, (dict, str))} if isinstance(v, dict) else v)
            find(v, k, path + "/" + kk)


print(d.get("metric"), d.get("value"), d.get("roofline", {}).get("frac"), "e2e", d.get("e2e", {}).get("value"), d.get("clocks", {}).get("reasons"))
for k in ("c1", "c2", "c2_dense", "c2x8", "c2x8_dense", "one_group_dense", "c5_scene", "c5_scenes_per_rank", "roialign"):
    find(d, k)
