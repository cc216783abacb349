for v in 0 1 2; do
  D4PG_PDL=$v python bench.py --no-cpu --steps 3000 2>/dev/null | python -c "import json,sys; d=json.loads([x for x in sys.stdin if x.startswith('{')][-1]); print('pdl=$v', round(d['ms_per_step']*1e3,2), round(d['value']))"
done
