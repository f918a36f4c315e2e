cd /root/repo
line() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f ms %.4f A %.4f B %.4f C %.4f'%(d['value'],d['ms_per_step'],d['stages']['colour_bicubic_ms'],d['roofline']['kernel_ms'],d['stages']['merge_ms']))"; }
for cfg in "4 0" "5 0" "6 0" "8 0" "6 1"; do set -- $cfg
echo "cfg2 ISR=$1 TMA=$2: $(SRCNN_KA_ISR=$1 SRCNN_KA_TMA=$2 python bench.py --steps 50 --warmup 5 --no-cpu 2>/dev/null | line)"
done
for c in cfg3 cfg5; do
for cfg in "8 0" "8 1" "16 0" "16 1" "6 0"; do set -- $cfg
echo "$c ISR=$1 TMA=$2: $(SRCNN_KA_ISR=$1 SRCNN_KA_TMA=$2 python bench.py --config $c --steps 3 --warmup 3 --no-cpu 2>/dev/null | line)"
done
echo "$c KA_INT=0: $(SRCNN_KA_INT=0 python bench.py --config $c --steps 3 --warmup 3 --no-cpu 2>/dev/null | line)"
done
