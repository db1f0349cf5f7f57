for d in 0 1 2 3 4 8 12; do echo "== debug $d"; PO2_WT_DEBUG=$d timeout 120 python tools/bench_wgrad.py --compute 2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try:
        r=json.loads(l)
        if 'r56' in r['layer']: print(r['layer'], 'po2 %.2f us'%r['po2_us'])
    except Exception: pass
"; done
