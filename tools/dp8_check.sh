N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
NVLS_CHECK_NO_TIMING=1 timeout 300 $T tools/nvls_check.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | head -12
echo "== nvls (auto)"; timeout 300 $T bench.py --gpus $N --steps 80 --warmup 10 --step-only 2>&1 | grep '"metric"' | cut -c1-220
echo "== nccl";        MMER_DP_MODE=nccl timeout 300 $T bench.py --gpus $N --steps 80 --warmup 10 --step-only 2>&1 | grep '"metric"' | cut -c1-220
echo "== 1 gpu";       python bench.py --steps 80 --warmup 10 --step-only 2>&1 | grep '"metric"' | cut -c1-220
