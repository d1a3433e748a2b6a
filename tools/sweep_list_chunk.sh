for ch in 2 4 8 16 64; do PF_CUDA_LIST_CHUNK=$ch python bench.py --workload c5_batch_512 --steps 20 --warmup 5 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
p=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('chunk=$ch', 'e2e ms', round(p['e2e']['ms_per_step'],4), 'dev ms', round(p['ms_per_step'],4), 'h2d', p['e2e']['h2d_bytes_per_step'], 'd2h', p['e2e']['d2h_bytes_per_step'])"; done
