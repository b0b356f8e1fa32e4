mkdir -p gpurun_out/r2s
nproc; nvidia-smi topo -m | head -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 tools/pcie8.py 2>/dev/null | tee gpurun_out/r2s/pcie8_n8.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2s/bench_n8.json 2> gpurun_out/r2s/bench_n8.err
tail -2 gpurun_out/r2s/bench_n8.err
python tools/parse_bench.py < gpurun_out/r2s/bench_n8.json | head -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2s/bench_n4.json 2> gpurun_out/r2s/bench_n4.err
python tools/parse_bench.py < gpurun_out/r2s/bench_n4.json | head -3
