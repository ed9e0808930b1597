TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
$TR bench.py --gpus 2 > gpurun_out/bench_sib_2gpu_p2p.json 2> gpurun_out/bench_sib_2gpu_p2p.err; tail -c 1200 gpurun_out/bench_sib_2gpu_p2p.json; tail -3 gpurun_out/bench_sib_2gpu_p2p.err
$TR bench.py --gpus 2 --no-p2p > gpurun_out/bench_sib_2gpu_nccl.json 2> gpurun_out/bench_sib_2gpu_nccl.err; tail -c 600 gpurun_out/bench_sib_2gpu_nccl.json
$TR bench.py --gpus 2 --workload gyroid --steps 10 > gpurun_out/bench_gy946_2gpu.json 2> gpurun_out/bench_gy946_2gpu.err; tail -c 1500 gpurun_out/bench_gy946_2gpu.json; tail -3 gpurun_out/bench_gy946_2gpu.err
