"""print the interesting parts of a bench.py JSON line"""
import json, sys
j = json.load(open(sys.argv[1]))
for k in ("value", "ms_per_step", "serial_ms_per_step", "gpu_launches"):
    print(k, round(j[k], 3) if isinstance(j[k], float) else j[k])
print("e2e", round(j["e2e"]["value"], 1), "sync api ms", round(j["e2e"]["sync_api_ms_per_step"], 3))
if "sustained" in j: print("sustained", round(j["sustained"]["value"], 1), j["sustained"]["clocks"])
for other in ("clustered", "uniform"):
    if other in j: print(other, round(j[other]["value"], 1), round(j[other]["ms_per_step"], 3), j[other]["per_op_ms"])
print("roofline", [(e["op"], round(e["launch_ms"], 3), round(e["frac"], 3), round(e.get("fp32_frac_algorithmic", 0), 3)) for e in j["roofline"]["kernels"]], "step_frac", round(j["roofline"]["step_frac"], 3))
print("per_op", dict(list(j["roofline"]["per_op_ms"].items())[:12]))
ifr = j.get("interframe_latency_ms")
if ifr:
    if "error" in ifr: print(ifr)
    else:
        for b in [k for k in ifr if k.startswith("batch")]:
            print(b, [s["ms"] for s in ifr[b]["steps"]], "p50", ifr[b]["p50_ms"])
st = j.get("streaming")
if st:
    print("streaming", {k: st[k] for k in st if k != "note"})
print("cpu", j.get("cpu_baseline"))
if "extra" in j: print("extra", j["extra"])
