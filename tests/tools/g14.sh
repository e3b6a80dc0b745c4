timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/tools/mgpu_check.py > gpurun_out/mg.log 2>&1
echo rc=$?; tail -40 gpurun_out/mg.log
