#!/usr/bin/env bash
# GPU session 13 (round 2): CTA-pair GEMM (cta_group::2, 256x128 tiles) against the default at the round-2 digit counts.
set -u
O=gpurun_out/r02_s13
mkdir -p $O
for g in int8 int8x2 int8 int8x2; do
  timeout 600 python bench.py --rows 1e6 --steps 3 --warmup 1 --no-e2e --no-cpu-baseline --oracle-rows 0 --no-peaks --no-check --gemm $g > $O/${g}_$RANDOM.json 2> $O/err.txt
  echo "$g rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_s13/*.json')):
    j=json.loads(open(f).read().strip().splitlines()[-1])
    k={r['slot']:r for r in j['roofline']['kernels']}
    print(f.split('/')[-1],'ms',round(j['ms_per_step'],1),'gram',round(k['k_gram']['ms_total'],1),round(k['k_gram']['issued_int8_tops']),'z',round(k['k_zgemm']['ms_total'],1),round(k['k_zgemm']['issued_int8_tops']),'clk',j['clocks']['sm_mhz'],j['clocks']['power_w_median'])
PY
