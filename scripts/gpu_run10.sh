python -m pytest tests -m gpu -x -q -k "topk" 2>&1 | tail -3
python scripts/topk_multi.py 250000 1024 32 2>&1 | tail -1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/topk_multi.py 250000 1024 32 2>&1 | tail -1
