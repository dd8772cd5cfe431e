"""Short readout of a bench.py JSON line: python tools/bench_summary.py gpurun_out/x.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'n_gpus')}, d.get('clocks'))
e = d.get('e2e') or {}
print('e2e', e.get('value'), 'GB/s sustained', e.get('h2d_GBps_sustained_all_ranks'), e.get('pinned_h2d_copy_GBps'))
r = d['roofline']
print('roofline', r['kernel'], round(r['frac'], 4), {k: (round(v['GBps']), round(v['share'], 3)) for k, v in r['per_kernel'].items()})
for k in ('full_extract_rcnn', 'full_extract_rcnn_topk100'):
    x = d.get(k)
    if not x:
        continue
    print(k, x.get('error') or (round(x['frames_per_s']), round(x['e2e']['value']), x['stage_ms_per_batch'], round(x['roofline']['frac'], 3),
                               x.get('dense_engine', {}).get('layer_shapes_on_tcgen05')))
a = d.get('azure')
if a:
    print('azure', a.get('error') or (round(a['frames_per_s']), round(a['e2e']['value']) if a.get('e2e') else None, a['roofline']['kernel'], round(a['roofline']['frac'], 3)))
print('cpu_baseline', d.get('cpu_baseline', {}) and {k: d['cpu_baseline'][k] for k in ('value', 'cores', 'kind')})
for k in ('features_regimes', 'prep_with_invalid_pixels', 'tracking_branch', 'secondary_error'):
    if k in d:
        print(k, json.dumps(d[k])[:600])
