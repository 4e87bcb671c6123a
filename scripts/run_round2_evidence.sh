# The round's evidence in one call (1 GPU): full GPU test suite, bench line of both arms, ncu launch list, ncu --set full of the hot kernels.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; tail -3 gpurun_out/f_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; tail -c 400 gpurun_out/f_bench.json
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; tail -c 600 gpurun_out/f_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/f_launches.csv python scripts/prof_step.py 10000000 2 > gpurun_out/f_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_count_mask2|k_fill3|k_expand_rows" -s 3 -c 3 -o gpurun_out/f_full python scripts/prof_step.py 10000000 2 > gpurun_out/f_f.log 2>&1
