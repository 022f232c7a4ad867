# Run under gpurun at the end of a round: full GPU test suite, smoke(), both bench arms, ncu launch lists + full-set captures
# (tools/gpu_profile.sh); the results are copied from gpurun_out/ into profiles/rNN_* (see DESIGN.md section 5).
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02_final.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02_final.log 2>&1
python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_ref_final.json 2> gpurun_out/bench_r02_ref_final.err
bash tools/gpu_profile.sh > gpurun_out/profile_final.log 2>&1
