import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'N=%d'%d["n_gpus"], 'value %.1f it/s'%d["value"], '%.3f ms/step'%d["ms_per_step"], "e2e %.1f"%d["e2e"]["value"], 'launches', d["gpu_launches"], 'roof', d["roofline"] and (d["roofline"]["kernel"], round(d["roofline"]["frac"],3))); print('   ', {k:round(v["ms_total"]/10,3) for k,v in d["kernel_groups_ms_per_10_iterations"].items()})
    except Exception as e: print("ERR",f,e)
