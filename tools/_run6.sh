python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --blobs 0 > gpurun_out/bench_r02k_n2.json 2> gpurun_out/bench_r02k_n2.err; echo "exit $?"
tail -5 gpurun_out/bench_r02k_n2.err | cut -c1-400
wc -c gpurun_out/bench_r02k_n2.json
