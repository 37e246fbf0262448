for hc in 148 222 296; do
  BPC_HOST_CHUNK=$hc python bench.py --steps 3 --warmup 3 --no-extras --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('host_chunk $hc', round(d['value']), round(d['e2e']['value']))"
done
