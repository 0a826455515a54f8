#!/bin/bash
# Run under gpurun: tools/post_ab.sh <tag> -- post-process timing (events + ncu per-kernel durations) of libronk.so and every _variants/*.so
tag=${1:-pab}
mkdir -p gpurun_out
lib=ron_tensorflow_b200/libronk.so
cp $lib /tmp/base.so
for so in /tmp/base.so ron_tensorflow_b200/_variants/*.so; do
  [ -f "$so" ] || continue
  v=$(basename $so .so)
  [ "$so" != /tmp/base.so ] && cp $so $lib
  echo "== $v" | tee -a gpurun_out/${tag}.txt
  python tools/post_time.py --batch 256 2>&1 | tail -1 | tee -a gpurun_out/${tag}.txt
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_$v.csv python tools/prof.py --stage post --batch 256 --iters 3 > /dev/null 2>&1
  python - <<PY | tee -a gpurun_out/${tag}.txt
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/${tag}_$v.csv') if not l.startswith('==')))
h = rows[0]; k = h.index('Kernel Name'); vi = h.index('Metric Value')
d = collections.OrderedDict()
for r in rows[1:]:
    if len(r) > vi and r[k].startswith('ronk::'):
        d.setdefault(r[k].split('(')[0], []).append(float(r[vi].replace(',', '')) / 1e3)
for n, v in d.items():
    print('   %-40s n=%d  last %.1f us  min %.1f' % (n, len(v), v[-1], min(v)))
PY
done
cp /tmp/base.so $lib
