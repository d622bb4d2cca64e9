#!/usr/bin/env bash
# Round-2 evidence run on 8 B200s of one box: C5 (n = 50M, p = 8192, Type-I, predictive mean + variance at 100 000 rows) and C3 at N = 8;
# the 2-GPU parity tests (torch.distributed and the library's own NCCL communicator).
set -u
O=gpurun_out/r02_final8c
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.txt 2>&1
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q > $O/pytest_multi_gpu.log 2>&1
echo "pytest multi rc=$?" >> $O/pytest_multi_gpu.log
tail -2 $O/pytest_multi_gpu.log
run() {  # name, N, args...
  name=$1; N=$2; shift 2
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > $O/$name.json 2> $O/$name.err
  echo "$name rc=$?"
}
run bench_c5_n8 8 --config C5 --steps 2 --warmup 3
run bench_c3_n8 8 --steps 3 --warmup 3
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_final8c/bench*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'unparsable',e); continue
    print(f, 'value',round(j['value'],4),'e2e',j['e2e'] and round(j['e2e']['value'],4),'ms',round(j['ms_per_step'],1), j['clocks'].get('sm_mhz'))
    print('   check', json.dumps(j.get('check'))[:700])
    print('   predict', json.dumps(j.get('predict'))[:400])
PY
