"""Condense bench.py JSON lines (files given as arguments, else stdin) to the handful of numbers compared
between experiments."""
import sys, json, fileinput
for l in fileinput.input():
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l)
        r = d.get("roofline") or {}
        print(round(d["value"], 1), round(d["e2e"]["value"], 1), {k: round(v, 3) for k, v in r.get("share_of_step", {}).items()},
              "att_TF", round(r.get("achieved", 0), 1), "agree", round((d.get("parity") or {}).get("bf16_vs_fp32_match_set_agreement", 0), 4),
              "MHz", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
