#!/usr/bin/env bash
# GPU session 5 (round 2): ncu --set full of one launch of every hot kernel at the default digits (6, 4); summaries exported on the box.
set -u
O=gpurun_out/r02_s5
mkdir -p $O
CMD="python bench.py --rows 3e5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-check --no-peaks"
$CMD > $O/plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain.log; exit 1; }
K='regex:k_build_phi|k_contract_rows|k_tables|k_ozaki|k_slot_hi'
ncu --set full --clock-control none --import-source on -k "$K" -s 34 -c 4 -o $O/pass1 $CMD > $O/ncu1.log 2>&1; echo "ncu1 rc=$?"
ncu --set full --clock-control none --import-source on -k "$K" -s 44 -c 3 -o $O/pass2 $CMD > $O/ncu2.log 2>&1; echo "ncu2 rc=$?"
for f in pass1 pass2; do
  ncu -i $O/$f.ncu-rep --page raw --csv > $O/${f}_raw.csv 2>/dev/null
  ncu -i $O/$f.ncu-rep --page source --csv --kernel-name regex:k_contract_rows > $O/${f}_source_contract.csv 2>/dev/null
done
ncu -i $O/pass1.ncu-rep --page source --csv --kernel-name regex:k_build_phi_t > $O/pass1_source_build_t.csv 2>/dev/null
ncu -i $O/pass2.ncu-rep --page source --csv --kernel-name regex:k_build_phi > $O/pass2_source_build.csv 2>/dev/null
ls -la $O
du -sm $O
sz=$(du -sm $O | cut -f1)
if [ "$sz" -gt 55 ]; then rm -f $O/pass1.ncu-rep; fi
sz=$(du -sm $O | cut -f1)
if [ "$sz" -gt 55 ]; then rm -f $O/pass2.ncu-rep; fi
du -sm $O
