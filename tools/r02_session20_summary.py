import json
j = json.loads(open('gpurun_out/r02_s20/bench_default.json').read().strip().splitlines()[-1])
print('value', j['value'], 'e2e', j['e2e'], 'ms', j['ms_per_step'], 'launches', j['gpu_launches'], 'clocks', j['clocks'])
print('roofline', {k: j['roofline'][k] for k in ('bound', 'achieved', 'peak', 'unit', 'frac')}, j['roofline']['traffic'])
for k in j['roofline']['kernels']:
    print('    ', k['slot'], k['launches'], round(k['ms_total'] / j['steps'], 1), round(k['share_of_step'], 4), k.get('issued_int8_tops'))
print('cpu', json.dumps(j['cpu_baseline'])[:300])
print('check', json.dumps(j['check'])[:900])
j = json.loads(open('gpurun_out/r02_s20/c5_sample.json').read().strip().splitlines()[-1])
k = {r['slot']: r for r in j['roofline']['kernels']}
print('C5 sample ms', round(j['ms_per_step'], 1), 'gram', round(k['k_gram']['issued_int8_tops']), json.dumps(j['check']['int8_vs_fp64_full_n'])[:300])
