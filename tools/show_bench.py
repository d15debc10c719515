import json, sys
d = json.load(open(sys.argv[1]))
print('value %.2fM  ms %.4f  frac %.3f  e2e %.2fM  trained %.2fM' % (d['value'] / 1e6, d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'] / 1e6, d['trained_like']['value'] / 1e6))
print('mlp', d.get('mlp'))
print('fp32', d['roofline'].get('fp32_pipe'), 'cpu', d.get('cpu_baseline', {}).get('value'))
