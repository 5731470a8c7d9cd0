for args in "" "" "--rad-chunk 28" "--rad-chunk 112" "--exc-chunk 334" "--exc-chunk 1000" "--no-graph" "--faithful"; do
  python bench.py --steps 600 --warmup 10 --no-cpu $args 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$args', '| value %.2fM e2e %.2fM ms/step %.4f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step']), {k: round(v,4) for k,v in d['kernel_ms'].items()}, 'rad frac %.3f exc frac %.3f' % (d['roofline']['frac'], d['roofline']['excitation']['frac']), d['clocks'])
"
done
