export MGPU_TRACE=1 MGPU_CASES=1 MALLOC_CHECK_=3 LD_PRELOAD=$PWD/tests/tools/dbg/abrt.so FITSNE_SHARDED_SYNC=1
timeout -k 5 45 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/tools/mgpu_check.py > gpurun_out/mg_case1.log 2>&1; echo "rc=$?" >> gpurun_out/mg_case1.log
grep -v "^W1017\|^\*\*\*\*\|^  File\|^    " gpurun_out/mg_case1.log | head -60
