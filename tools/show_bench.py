"""Print the interesting parts of a bench.py JSON line.  usage: python tools/show_bench.py gpurun_out/bench_x.json"""
import json, sys
d = None
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
print('value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'ms', round(d['ms_per_step'], 2), 'roof', round(d['roofline']['frac'], 4), 'n_gpus', d['n_gpus'])
for k, v in d['kernels'].items():
    print(' ', k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a != 'vs_sdpa'})
for r in d['kernels'].get('attention', {}).get('vs_sdpa', []):
    print('   sdpa', r['shape'], round(r['ours_ms'], 3), round(r['sdpa_cudnn_ms'], 3), round(r['speedup_vs_best_sdpa'], 2))
cl = d['clip_loss']
print('loss', round(cl['value'], 3), 'frac', round(cl['roofline']['frac'], 3), 'parity', cl.get('parity_ok'), {k: round(v['ms_per_step'], 3) for k, v in cl['kernels'].items()})
rg = d.get('reference_gpu') or {}
if 'tower' in rg:
    print('ref_gpu tower', round(rg['tower']['value'], 1), 'ours/ref', round(rg['tower'].get('ours_over_reference', 0), 2), 'loss ms', rg.get('clip_loss', {}).get('value'))
ew = d.get('extra_workloads', {})
print('latency', json.dumps(ew.get('small_batch_latency')))
for k in ('b16_384_fwd_bwd', 'h14_train_step'):
    w = ew.get(k)
    if w:
        print(k, round(w['ms_per_step'], 1), 'frac', round(w['roofline']['frac'], 3), {a: round(b['ms_per_step'], 1) for a, b in w['kernels'].items()})
print('errors', d.get('leg_errors'), 'clocks', d['clocks'])
print('cpu', d['cpu_baseline'])
