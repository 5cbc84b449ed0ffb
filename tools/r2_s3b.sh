mkdir -p gpurun_out
for v in "65536 1" "65536 0" "1024 1"; do
set -- $v
NASR_STREAM_GRAPH=0 NASR_SMALL_GATHER=$2 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/stream_${1}_g$2.csv python tools/stream_probe.py $1 6 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/stream_${1}_g$2.csv")) if len(r)>10 and r[0].isdigit()]
# last chunk's launches: take the last 13 rows
last=rows[-13:]
tot=0
for r in last:
    name=r[4][:60]; dur=float(r[-1].replace(',',''))/1e3; tot+=dur
    print(f"  {name:60s} grid {r[7]:>14s} {dur:8.2f} us")
print("chunk $1 small_gather=$2: sum of last 13 kernels", round(tot,1), "us")
PY
done
