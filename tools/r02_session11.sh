#!/usr/bin/env bash
# GPU session 11 (round 2): ncu of the two slab builders, the tail and the Gram GEMM at two slab budgets; slab-budget sweep of pass 1.
set -u
O=gpurun_out/r02_s11
mkdir -p $O
for mb in 4096 1024 512 256; do
  timeout 600 python bench.py --rows 1e6 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --oracle-rows 0 --no-peaks --no-check --slab-mb $mb > $O/slab_$mb.json 2> $O/slab_$mb.err
  echo "slab $mb rc=$?"
done
python - <<'PY'
import json
for mb in (4096,1024,512,256):
    j=json.loads(open('gpurun_out/r02_s11/slab_%d.json'%mb).read().strip().splitlines()[-1])
    k={r['slot']:r for r in j['roofline']['kernels']}
    print(mb,'ms',round(j['ms_per_step'],1),'gram',k['k_gram']['launches'],round(k['k_gram']['ms_total'],1),round(k['k_gram']['issued_int8_tops']),'build_t',round(k['k_build_phi_t']['ms_total'],1),'clk',j['clocks']['sm_mhz'],j['clocks']['power_w_median'])
PY
CMD="python bench.py --rows 3e5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-check --no-peaks"
ncu --set full --clock-control none --import-source on -k 'regex:k_build_phi' -s 16 -c 4 -o $O/builders $CMD > $O/ncu1.log 2>&1; echo "ncu1 rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:k_contract_tail' -s 12 -c 1 -o $O/tail $CMD > $O/ncu2.log 2>&1; echo "ncu2 rc=$?"
ncu --set full --clock-control none -k 'regex:k_ozaki' -s 16 -c 2 -o $O/gram4096 $CMD > $O/ncu3.log 2>&1; echo "ncu3 rc=$?"
ncu --set full --clock-control none -k 'regex:k_ozaki' -s 44 -c 2 -o $O/gram512 $CMD --slab-mb 512 > $O/ncu4.log 2>&1; echo "ncu4 rc=$?"
for f in builders tail gram4096 gram512; do ncu -i $O/$f.ncu-rep --page raw --csv > $O/${f}_raw.csv 2>/dev/null; done
ncu -i $O/builders.ncu-rep --page source --csv --kernel-name 'regex:k_build_phi<' > $O/source_build.csv 2>/dev/null
ncu -i $O/builders.ncu-rep --page source --csv --kernel-name 'regex:k_build_phi_t' > $O/source_build_t.csv 2>/dev/null
ncu -i $O/tail.ncu-rep --page source --csv > $O/source_tail.csv 2>/dev/null
du -sm $O
sz=$(du -sm $O | cut -f1); if [ "$sz" -gt 55 ]; then rm -f $O/builders.ncu-rep; fi
