for v in 1 0; do
  D4PG_COMM_PEER=$v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$v bench.py --gpus 2 --steps 3000 --warmup 20 2>/dev/null | python -c "import json,sys; L=[x for x in sys.stdin if x.startswith('{')]; d=json.loads(L[-1]) if L else None; print('peer=$v', (round(d['ms_per_step']*1e3,2), round(d['value']), d['kernels_per_step'], d['losses']) if d else 'FAILED')"
done
