# final-state check of the round (one B200): build + smoke, full GPU suite, default bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | grep -v "^Dilations" | tail -8
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py > gpurun_out/r2_final_n1.json 2> gpurun_out/r2_final_n1.err
tail -c 600 gpurun_out/r2_final_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_final_n1.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"],d["clocks"])
print("roof",{k:d["roofline"].get(k) for k in ("bound","frac","hbm_frac","launch_ms","achieved","tensor_frac_algorithmic","issued_fp16_frac_of_sustained_peak")})
print("cpu", d["cpu_baseline"])
for k,v in d["configs"].items():
    if "error" in v: print(k,"ERROR",v["error"]); continue
    if k=="shipped": print(k, {n[:24]:(c["msamples_per_s"],c["msamples_per_s_fp32"]) for n,c in v.get("checkpoints",{}).items()})
    elif k=="cfg5": print(k,{q:{z:v[q].get(z) for z in ("samples_per_s","us_per_chunk")} for q in v if q.startswith("chunk")})
    else: print(k,{z:v.get(z) for z in ("samples_per_s","e2e_samples_per_s","ms_per_step")}, v.get("e2e"), {z:v["roofline"].get(z) for z in ("bound","frac")} if "roofline" in v else None)
PY
