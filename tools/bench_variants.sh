for v in "A=1" "BENCH_PENDING=100" "BENCH_NO_SAMPLER=1" "BENCH_NO_SAMPLER=1 BENCH_PENDING=100"; do
  env $v python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['ms_per_step'],3), round(d['value']/1e8,3), round(d['e2e']['value']/1e8,3))"
done
