import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "steps", "clocks", "gpu_launches_per_step")})
print("e2e", d.get("e2e"))
r = d.get("roofline") or {}
print("roofline", {k: r.get(k) for k in ("kernel", "frac", "achieved", "traffic", "avg_launch_us")})
print("cpu", d.get("cpu_baseline"))
