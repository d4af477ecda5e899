N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --blobs 0 > gpurun_out/bench_r02q_n$N.json 2> gpurun_out/bench_r02q_n$N.err; echo "exit $?"
tail -2 gpurun_out/bench_r02q_n$N.err | cut -c1-300
