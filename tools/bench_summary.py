import json, sys
d = json.load(open(sys.argv[1]))
print({k: round(d[k], 2) for k in ("value", "ms_per_step")}, "fwd TF/s", round(d["roofline"]["achieved"], 1), "frac", round(d["roofline"]["frac"], 3))
print({k: round(v["ms"], 3) for k, v in d["stages"].items()}, "e2e", round(d["e2e"]["value"], 1), d.get("clocks"))
import signal; signal.signal(signal.SIGPIPE, signal.SIG_DFL)
for l in d["layers"]:
    print(f'  {l["launch"]:14s} {l["ms"]:8.3f} ms {l["tflops"]:8.1f} TF/s')
if "cpu_baseline" in d: print(d["cpu_baseline"])
