mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
( $TR scripts/configs_multi.py c4 4000 8000 1024 && $TR scripts/configs_multi.py c5 16000 1024 32 \
  && $TR scripts/configs_multi.py c4 50000 100000 1024 && timeout 300 $TR scripts/configs_multi.py c5 1000000 1024 32 ) > gpurun_out/configs_n8.log 2>&1
grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$" gpurun_out/configs_n8.log | tail -20
