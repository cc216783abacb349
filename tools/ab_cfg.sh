for c in c3 c5; do for v in fp32 tf32x3; do for ch in 1 0; do
  timeout 300 python bench.py --no-cpu --steps 300 --warmup 5 --config $c --precision $v --chain $ch 2>&1 | python -c "import json,sys; L=[x for x in sys.stdin if x.startswith('{')]; d=json.loads(L[-1]) if L else None; print('$c $v chain=$ch', (round(d['ms_per_step']*1e3,1), round(d['value']), d['kernels_per_step']) if d else 'FAILED')"
done; done; done
