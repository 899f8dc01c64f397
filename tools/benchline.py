"""Reads bench.py's JSON line on stdin and prints the few numbers worth watching."""
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d["value"]), "Ms/s", round(d["ms_per_step"], 3), "ms/step | e2e", round(d["e2e"]["value"]),
      {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()}, "| dec", round(d["decompress"]["value"]),
      "GB/s kernel", round(d["decompress"]["kernel_gbs"] or 0), d["clocks"])
