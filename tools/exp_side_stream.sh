for v in "0 " "1 " "1 10=3" "1 10=2" "0 10=3"; do set -- $v; echo "== side=$1 debug=$2"; DM_WGRAD_STREAM=$1 DM_DEBUG=$2 DM_BENCH_FAST=1 python bench.py --steps 6 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['ms_per_step'],2),'clk',d['clocks']['sm_mhz'], 'wgrad', round(d['roofline']['wgrad']['achieved']))
    elif 'Error' in l or 'error' in l: print(l.strip())
"; done
