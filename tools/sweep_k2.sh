for pt in 256 512; do for sl in 8 16 32; do echo "== part.threads=$pt lut_slice=${sl}MB"; python tools/ops_bench.py --ops groupby,join --join-scale 0.25 --reps 2 --opt part.threads=$pt --opt join.lut_slice_bytes=$((sl*1048576)) | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], d.get('rows'), 'total_ms', round(d.get('total_ms',0),3), 'kernel_ms', round(d.get('kernel_ms',0),3), 'ok', d.get('check_ok'), d.get('error',''))"; done; done
