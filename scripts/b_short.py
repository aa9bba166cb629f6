import json, sys
d = json.loads(open(sys.argv[1]).read())
print(sys.argv[2], d["ms_per_step"], {k: round(v["avg_ms"], 4) for k, v in d["roofline"]["kernels"].items()})
