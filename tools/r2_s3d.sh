# dev: session-3 final-state evidence (one B200)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/s3d_tests.log
python bench.py > gpurun_out/r2_final_n1.json 2> gpurun_out/r2_final_n1.err
python tools/ckpt_bench.py --json gpurun_out/s3_ckpt.json 2>&1 | grep -v "^Dilations" > gpurun_out/s3_ckpt.log
NCU="ncu --set full --import-source on --clock-control none -f"
$NCU -k regex:ring_block -s 6 -c 2 -o gpurun_out/r2_ring_gcn64 python tools/ckpt_bench.py --only GCN_3 --iters 2 > gpurun_out/r2_ncu_e.log 2>&1
$NCU -k regex:ring_block -s 14 -c 3 -o gpurun_out/r2_ring_passes python tools/ckpt_bench.py --only TCN_99_egfx --iters 2 > gpurun_out/r2_ncu_f.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_shipped.csv python tools/ckpt_bench.py --iters 1 > /dev/null 2>&1
cat gpurun_out/s3d_tests.log gpurun_out/s3_ckpt.log; tail -c 1500 gpurun_out/r2_final_n1.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_final_n1.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"],d["clocks"])
print("roof",{k:d["roofline"].get(k) for k in ("bound","frac","hbm_frac","launch_ms","achieved")})
for k,v in d["configs"].items():
    if k=="shipped": print(k, {n:(c["msamples_per_s"],c["msamples_per_s_fp32"]) for n,c in v.get("checkpoints",{}).items()} or v)
    elif k=="cfg5": print(k,{q:{z:v[q].get(z) for z in ("samples_per_s","us_per_chunk")} for q in v if q.startswith("chunk")}, v.get("error"))
    else: print(k,{z:v.get(z) for z in ("samples_per_s","e2e_samples_per_s","ms_per_step","error")})
PY
ls -la gpurun_out/*.ncu-rep
