import json, sys
d = json.load(open(sys.argv[1]))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("per_step_kernel_ms", d["per_step_kernel_ms"], "launches", d["launches_per_step"])
for k in d["kernels"][:n]:
    gbs = k["alg_bytes"] / 1e6 / max(k["ms_per_step"] / max(k["launches_per_step"], 1), 1e-9)
    print(f"{k['name']:26s} {str(k['args'])[:50]:50s} n={k['launches_per_step']:.0f} ms={k['ms_per_step']:.3f} {k['share']*100:4.1f}% algMB={k['alg_bytes']/1e6:6.0f} GB/s={gbs:5.0f}")
