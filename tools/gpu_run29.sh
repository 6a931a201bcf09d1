cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2b
N=${1:-2}
for B in 10 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --blocks $B > gpurun_out/r2b/scale_N${N}_B$B.json 2> gpurun_out/r2b/scale_N${N}_B$B.err || tail -20 gpurun_out/r2b/scale_N${N}_B$B.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2b/scale_N${N}_B$B.json')); print('N=$N blocks=$B ms',round(d['ms_per_step'],3),'value %.3e'%d['value'],'k1',round(d['roofline']['kernel_ms'],3),'parity',d['cpu_baseline'].get('parity_on_sample'),d['exchange'])
PY
done
