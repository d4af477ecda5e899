python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 > gpurun_out/bench_r02q_n8.json 2> gpurun_out/bench_r02q_n8.err; echo "exit $?"
python bench.py --gpus 8 --single-process --blobs 0 > gpurun_out/bench_r02q_n8_sp.json 2> gpurun_out/bench_r02q_n8_sp.err; echo "exit $?"
