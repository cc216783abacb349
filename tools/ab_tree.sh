for i in 1 2; do
  for v in fast slow; do
    if [ $v = slow ]; then export D4PG_TREE_SLOW=1; else unset D4PG_TREE_SLOW; fi
    python bench.py --no-cpu --steps 3000 2>/dev/null | python -c "import json,sys; d=json.loads([x for x in sys.stdin if x.startswith('{')][-1]); print('$v', round(d['ms_per_step']*1e3,2), round(d['value']))"
  done
done
