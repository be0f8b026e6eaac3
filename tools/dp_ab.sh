run() { echo "== $1"; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 100 --warmup 10 --step-only 2>&1 | grep '"metric"' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"; }
run single_gpu_ref_n1 X=1
python bench.py --step-only --steps 100 --warmup 10 2>&1 | tail -1 | cut -c1-200
run overlap_default X=1
run overlap_reserve4_ctas4 MMER_DEBUG=9=4 NCCL_MAX_CTAS=4
run overlap_reserve8_ctas8 MMER_DEBUG=9=8 NCCL_MAX_CTAS=8
run overlap_ctas4_noreserve NCCL_MAX_CTAS=4
run no_overlap MMER_DP_OVERLAP=0
run overlap_default_again X=1
