export MGPU_TRACE=1 MALLOC_CHECK_=3 LD_PRELOAD=$PWD/tests/tools/dbg/abrt.so FITSNE_SHARDED_SYNC=1
run() { name=$1; shift
  ( env "$@" timeout -k 5 40 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT tests/tools/mgpu_check.py > gpurun_out/mg2_$name.log 2>&1; echo "rc=$?" >> gpurun_out/mg2_$name.log )
  echo "== $name"; grep -E "rank 0\] closed|dims=|MGPU_OK|free|malloc|corrupt|rc=|\.so|python\(" gpurun_out/mg2_$name.log | head -40
}
PORT=29517 run full MGPU_CASES=0,1
PORT=29518 run twice1d MGPU_CASES=1,1
PORT=29519 run nosingle MGPU_CASES=0,1 MGPU_SKIP_SINGLE=1
