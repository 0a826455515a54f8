#!/bin/bash
# Run under gpurun:  tools/gpu_check.sh <tag> [full]   -> gpurun_out/<tag>_*
# GPU parity tests, bench (both arms optional), ncu launch list of one bench pass, optional ncu --set full capture.
tag=${1:-run}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
tail -3 gpurun_out/${tag}_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/${tag}_bench.json'))
    print('encode b64  %.1f us  frac %.3f  e2e %.0f img/s' % (d['ms_per_step']*1e3, d['roofline']['frac'], d['e2e']['value']))
    s = d['stages']
    print('encode b256 %.1f us (single %.1f)  frac %.3f' % (s['encode_b256']['ms_per_step']*1e3, s['encode_b256'].get('single_stream_ms_per_step', 0)*1e3, s['encode_b256']['roofline']['frac']))
    for k in ('ssd512_b128', 'crowded'):
        for kk, v in (s.get(k) or {}).items():
            if isinstance(v, dict) and 'value' in v: print('%s.%s %.0f img/s frac %.3f' % (k, kk, v['value'], v['roofline']['frac']))
    if s.get('postprocess'):
        print('post b256   %.1f us  frac %.3f  e2e %.0f img/s' % (s['postprocess']['ms_per_step']*1e3, s['postprocess']['roofline']['frac'], s['postprocess']['e2e']['value']))
    print('cpu', d.get('cpu_baseline'))
except Exception as e:
    print('bench parse failed', e)
    print(open('gpurun_out/${tag}_bench.err').read()[-2000:])
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/${tag}_launches.csv')) if len(r) > 10 and r[0].isdigit()]
t = collections.defaultdict(list)
for r in rows:
    t[r[4].split('(')[0][:60] + ' grid' + r[8]].append(float(r[-1]) / 1e3)
for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1])):
    print('%-90s n=%3d  avg %.1f us' % (k, len(v), sum(v) / len(v)))
PY
if [ "$2" = "full" ]; then
  ncu --set full --clock-control none --import-source on -k regex:match_encode -s 2 -c 1 -o gpurun_out/${tag}_enc python tools/prof.py --stage encode --batch 256 --iters 3 > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k regex:match_encode -s 2 -c 1 -o gpurun_out/${tag}_enc64 python tools/prof.py --stage encode --batch 64 --iters 3 > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k 'regex:scatter|pivot|select_topk|nms_kernel|tpfp' -s 6 -c 6 -o gpurun_out/${tag}_post python tools/prof.py --stage post --batch 256 --iters 2 > /dev/null 2>&1
fi
