cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
for w in detection_abc_kleene_1Mx100 detection_kleene_all_1Mx100; do
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 40 --csv --log-file gpurun_out/r2/launches_$w.csv python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2/ncu_$w.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/r2/launches_$w.csv')))
for i,r in enumerate(rows):
    if r and r[0]=='ID': h=r; start=i+1; break
ki=h.index('Kernel Name'); vi=h.index('Metric Value'); mi=h.index('Metric Name'); ii=h.index('ID')
d={}
for r in rows[start:]:
    if len(r)>vi: d.setdefault(r[ii],{'k':r[ki][:70]})[r[mi]]=r[vi]
ids=sorted(d,key=int)[-9:]
for i in ids: print(d[i])
PY
done
