export MALLOC_CHECK_=3
run() { name=$1; shift
  ( env "$@" timeout -k 5 45 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT tests/tools/mgpu_debug.py > gpurun_out/dbg_$name.log 2>&1; echo "rc=$?" >> gpurun_out/dbg_$name.log )
  echo "== $name"; grep -E "^it |DEBUG_DONE|free|malloc|corrupt|rc=|backtrace|fitsne|Error" gpurun_out/dbg_$name.log | head -30
}
PORT=29517 run d1 LD_PRELOAD=$PWD/tests/tools/dbg/abrt.so
PORT=29518 run d1_noreorder SINGLE_FLAGS=4
