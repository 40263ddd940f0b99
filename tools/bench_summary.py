import json,sys
j=json.load(open(sys.argv[1]))
print(sys.argv[1], "ms/step %.1f e2e %.1f"%(j["ms_per_step"], j["e2e"]["ms_per_step"]), {k:round(v,3) for k,v in j["phases_s"].items()})
r=j["roofline"]; print("  dominant:", r["kernel"][:40], "ms %.1f frac %.3f ach %.1f %s"%(r["ms_per_step"], r["frac"], r["achieved"], r["unit"]), "fp64eq", r.get("fp64_equivalent_tflops"))
for k,v in j["kernels"].items():
    print("  ", k[:44], {kk:(round(vv,2) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("ms_per_step","launches_per_step","achieved","frac","gbs","fp64_equivalent_tflops")})
