import json, sys
l = [x for x in open(sys.argv[1]) if x.startswith('{')]
d = json.loads(l[-1])
k = d.pop('kernels', None) or {}
print({a: d[a] for a in ('value', 'ms_per_step', 'tflops_algorithmic', 'gpu_launches')}, 'e2e', d['e2e']['value'], 'clocks', d.get('clocks'))
print('roofline', d.get('roofline')); print('attn', d.get('roofline_attention')); print('cpu', d.get('cpu_baseline'))
tot = sum(v['ms_per_launch'] * v['launches_per_step'] for v in k.values())
for n, v in sorted(k.items(), key=lambda kv: -kv[1]['ms_per_launch'] * kv[1]['launches_per_step']):
    print(f"{n:22s} {v['ms_per_launch']:8.3f} ms x{v['launches_per_step']}  {100*v['ms_per_launch']*v['launches_per_step']/tot:5.1f}%  {v['tflops'] and round(v['tflops'],1)} TF/s")
print('sum of kernels', round(tot, 3), 'ms')
