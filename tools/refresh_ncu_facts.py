"""Refresh one precision block of profiles/ncu_facts_r02.json from a tools/summarize_ncu.py JSON.
   python tools/refresh_ncu_facts.py strict /tmp/ncu_strict.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"head": "head", "head_1x1": "head_1x1", "dcn64": "dcn", "dcn128": "dcn_128", "offconv64": "offset_convs",
        "conv128": "base_convs", "conv64": "base_convs_64", "conv512": "base_convs_512", "stem": "stem", "level0": "stem_level0",
        "head2_reduce": "head_1x1"}
prec, src = sys.argv[1], sys.argv[2]
path = os.path.join(ROOT, "profiles", "ncu_facts_r02.json")
facts = json.load(open(path))
new = json.load(open(src))
for name, row in new.items():
    key = KEYS.get(name)
    if key is None:
        continue
    old = facts[prec].get(key, {})
    facts[prec][key] = {"tensor_pipe_pct": round(row["tensor_pipe_pct"], 1), "issue_active_pct": round(row["issue_active_pct"], 1),
                        "lsu_wavefront_pct": round(row["lsu_wavefront_pct"], 1),
                        "dram_bytes_per_launch": int(row["dram_read"] + row["dram_write"]), "ncu_us": round(row["us"], 1),
                        "launch": old.get("launch", name), "kernel": row["kernel"]}
if len(sys.argv) > 3:
    facts["source"] = sys.argv[3]
json.dump(facts, open(path, "w"), indent=1)
print("updated", prec, sorted(new))
