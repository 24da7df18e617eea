import sys,json
for ln in sys.stdin:
    ln=ln.strip()
    if not ln.startswith('{'): 
        print(ln); continue
    d=json.loads(ln)
    print(d["config"]["workload"][:40], "| Gp/s %.3f | ms %.3f |"%(d["value"], d["ms_per_step"]), d["phases_ms"], "| frac %.4f |"%d["roofline"]["frac"], d["stats"], d.get("e2e",{}).get("value"), d.get("cpu_baseline",{}).get("value"))
