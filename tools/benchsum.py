import sys,json
for ln in sys.stdin:
    ln=ln.strip()
    if not ln.startswith('{'): 
        print(ln[:200]); continue
    d=json.loads(ln)
    print("%s value=%.4g kernel=%.4g frac=%.3f phases=%s e2e=%.4g launches=%d"%(d["config"]["workload"][:5], d["value"], d["roofline"]["kernel_node_updates_per_s"], d["roofline"]["frac"], {k:round(v,3) for k,v in d["phase_ms_per_step"].items()}, d["e2e"]["value"] if d.get("e2e") else 0, d["gpu_launches"]))
