# 2 GPUs, tightly bounded: which build / mode of the sharded path fails in mgpu_check, and with what message
export MGPU_TRACE=1
run() { # name, env...
  name=$1; shift
  ( env "$@" timeout -k 5 55 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT tests/tools/mgpu_check.py > gpurun_out/mg_$name.log 2>&1; echo "rc=$?" >> gpurun_out/mg_$name.log )
  echo "== $name: $(grep -c MGPU_OK gpurun_out/mg_$name.log) ok, $(tail -1 gpurun_out/mg_$name.log)"; grep -E "rel-L2|CUDA error|illegal|NCCL|what\(\)|Error" gpurun_out/mg_$name.log | head -8
}
PORT=29517 run old FITSNE_LIB=$PWD/fit-sne_b200/lib/old/libfitsne_b200.so
PORT=29518 run new_sync FITSNE_SHARDED_SYNC=1
PORT=29519 run new_batched FITSNE_TRACE=0
PORT=29520 run new_sync_even FITSNE_SHARDED_SYNC=1 MGPU_N=200000
