for v in fp32 tf32x3 tf32; do
  python bench.py --no-cpu --steps 3000 --precision $v 2>/dev/null | python -c "import json,sys; d=json.loads([x for x in sys.stdin if x.startswith('{')][-1]); print('$v', round(d['ms_per_step']*1e3,2), round(d['value']), round(d['e2e']['value']), d['losses'])"
done
