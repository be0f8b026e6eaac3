# N independent single-GPU runs at the same time (no communication at all) against one run alone: separates what a
# multi-GPU job loses to sharing the chassis (power, host) from what the gradient exchange costs.
N=${1:-4}
python bench.py --step-only --steps 150 --warmup 10 2>/dev/null | tail -1 | cut -c1-160
for i in $(seq 0 $((N-1))); do
  CUDA_VISIBLE_DEVICES=$i python bench.py --step-only --steps 150 --warmup 10 2>/dev/null | tail -1 | cut -c1-160 > /tmp/conc_$i.log &
done
wait
for i in $(seq 0 $((N-1))); do echo "concurrent gpu $i: $(cat /tmp/conc_$i.log)"; done
for i in $(seq 0 $((N-1))); do echo "alone gpu $i: $(CUDA_VISIBLE_DEVICES=$i python bench.py --step-only --steps 150 --warmup 10 2>/dev/null | tail -1 | cut -c60-160)"; done
