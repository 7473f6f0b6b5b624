N=$1
if [ "$N" = "1" ]; then
  ( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2f_n1.json 2> gpurun_out/bench_r2f_n1.err ) 2> gpurun_out/bench_r2f_n1.time; echo rc=$?; tail -3 gpurun_out/bench_r2f_n1.time
else
  ( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2f_n$N.json 2> gpurun_out/bench_r2f_n$N.err ) 2> gpurun_out/bench_r2f_n$N.time; echo rc=$?; tail -3 gpurun_out/bench_r2f_n$N.time
fi
