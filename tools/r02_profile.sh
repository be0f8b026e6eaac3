# Round-2 evidence run (one gpurun call, 1 GPU): the bench line, the launch list of the training step, and ncu --set full
# captures of the step's top kernels as the step launches them.  Reports land in gpurun_out/ (scratch); the summaries
# under profiles/ are made from them with tools/launch_summary.py and tools/ncu_summary.py.
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err || exit 1
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
python bench.py --step-only --steps 4 --warmup 2 > gpurun_out/n_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --step-only --steps 4 --warmup 2 > gpurun_out/n_launches.log 2>&1
for k in "gemm_tc_kernel<256, 1, 1, 2, 0" mha_bwd_tma add_ln_bwd_pipe; do
  tag=$(echo "$k" | tr -c 'a-zA-Z0-9' '_' | cut -c1-24)
  ncu --set full --clock-control none --import-source on -k "regex:$k" -s 12 -c 1 -f -o gpurun_out/r02_step_$tag \
      python bench.py --step-only --steps 4 --warmup 2 > gpurun_out/n_$tag.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
