import json
r=json.load(open("gpurun_out/step_breakdown.json"))
print({k:(round(v,2) if isinstance(v,float) else v) for k,v in r.items() if k not in("ops","gemm_shapes")})
for k,v in list(r["ops"].items())[:3]: print("%-22s %9.3f ms %5d calls" % (k, v["ms"], v["calls"]))
for g in r["gemm_shapes"][:8]: print(g)
