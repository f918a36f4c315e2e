cd /root/repo
for r in 1 2; do for lib in build/alt/lib_*.so; do
SRCNN_B200_LIB=$PWD/$lib python bench.py --steps 50 --warmup 5 --no-cpu 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$lib','value %.0f A %.4f B %.4f C %.4f e2e %.0f'%(d['value'],d['stages']['colour_bicubic_ms'],d['roofline']['kernel_ms'],d['stages']['merge_ms'],d['e2e']['value']))"
done; done
