set -x
python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/f_pytest.log
python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
python bench.py --ig --steps 10 --warmup 3 > gpurun_out/f_ig.json 2>/dev/null
python tools/kernel_bench.py > gpurun_out/f_kernel_bench.txt 2>&1
python tools/gemm_bench.py 20 --cublas > gpurun_out/f_gemm_bench.txt 2>&1
python tools/config_bench.py > gpurun_out/f_config_bench.txt 2>&1
python tools/graph_step.py > gpurun_out/f_graph_step.txt 2>&1
python bench.py --step-only --steps 8 --warmup 3 > gpurun_out/f_plain_step.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/f_launches.csv python bench.py --step-only --steps 8 --warmup 3 > gpurun_out/f_ncu.log 2>&1
tail -2 gpurun_out/f_pytest.log; cat gpurun_out/f_bench.json | cut -c1-400
