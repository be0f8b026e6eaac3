# ncu --set full of single launches of the step's top non-GEMM kernels, as the training step launches them
# (after the same command has run plain).  Reports go to gpurun_out/ (scratch); summaries are made with tools/ncu_summary.py.
set -x
python bench.py --step-only --steps 4 --warmup 2 > gpurun_out/n_plain.log 2>&1 || exit 1
for k in mha_bwd_tma add_ln_bwd_pipe embed_bwd_pipe; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/r01_step_$k \
      python bench.py --step-only --steps 4 --warmup 2 > gpurun_out/n_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
