python -m pytest tests -m gpu -x -q 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3_phong_4k --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/b.log 2>&1
grep k_setup gpurun_out/launches_c3.csv | tail -2 | cut -d, -f5,15
python bench.py --workload c3_phong_4k --steps 10 --warmup 5 --no-extra --no-cpu-baseline 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('C3 dev %.4f raster %.4f front %.4f'%(j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frontend_kernels_ms']))"
python bench.py --steps 30 --warmup 5 --no-extra --no-cpu-baseline 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('C2 dev %.4f raster %.4f front %.4f'%(j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frontend_kernels_ms']))"
