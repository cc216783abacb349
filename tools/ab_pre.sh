for v in pre nopre; do
  if [ $v = nopre ]; then export D4PG_NO_PRE=1; else unset D4PG_NO_PRE; fi
  python bench.py --no-cpu --steps 3000 2>/dev/null | python -c "import json,sys; d=json.loads([x for x in sys.stdin if x.startswith('{')][-1]); print('$v', round(d['ms_per_step']*1e3,2), round(d['value']), round(d['e2e']['value']))"
done
