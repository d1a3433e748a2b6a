for rows in 3 4 5 6; do for w in c2_textured_1080p c3_phong_4k c1_gears_800x600; do
  PF_CUDA_BIN_ROWS=$rows python bench.py --workload $w --steps 20 --warmup 5 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
p=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=p['roofline']
print('rows=$rows', '$w', 'ms', round(p['ms_per_step'],4), 'raster', round(r['kernel_ms'],4), 'front', round(r['frontend_kernels_ms'],4), 'e2e', round(p['e2e']['ms_per_step'],4))"
done; done
