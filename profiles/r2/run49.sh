for v in 1 0; do
D3B_ALPHA_BRANCH=$v timeout 300 python bench.py --workload c5 --steps 40 --warmup 5 --headline-only 2>/dev/null | python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('alpha_branch=$v c5',round(d['value'],2),round(d['ms_per_step']*1e3,1))"
done
