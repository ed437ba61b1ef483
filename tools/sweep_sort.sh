for opt in "sort.rank=0" "sort.rank=1" "sort.rank=0 --opt sort.ctas_per_sm=1" "sort.impl=1"; do echo "== $opt"; python tools/ops_bench.py --ops orderby --scale 0.25 --reps 2 --opt $opt | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], d.get('rows'), 'total_ms', round(d.get('total_ms',0),3), 'kernel_ms', round(d.get('kernel_ms',0),3), 'ok', d.get('check_ok'), d.get('error',''))"; done
