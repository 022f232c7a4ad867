set -x
mkdir -p gpurun_out
python -m pytest tests/test_dcm.py tests/test_materials.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/c1_pytest.log
python bench.py --steps 3 --warmup 1 --no-kirchhoff --no-dcm --no-rejection --no-cpu-baseline --config5-wave-n 20000 --config5-wave-m 65536 > gpurun_out/c1_bench_a.json 2> gpurun_out/c1_bench_a.err
python bench.py --steps 3 --warmup 1 --no-kirchhoff --no-rejection --no-cpu-baseline --config5-wave-n 20000 --config5-wave-m 65536 > gpurun_out/c1_bench_b.json 2> gpurun_out/c1_bench_b.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c1_launches_c5.csv python bench.py --steps 1 --warmup 1 --nrays 20000000 --no-kirchhoff --no-dcm --no-rejection --no-cpu-baseline --config5-wave-n 20000 --config5-wave-m 65536 > gpurun_out/c1_ncu.log 2>&1
nvidia-smi --query-gpu=memory.used,memory.total --format=csv >> gpurun_out/c1_ncu.log
