# 1 -> 8 GPUs of one box, same code, back to back (what the driver does at round end): N=1 and N=8 bench lines + C5 at N=8.
out=gpurun_out
tag=${1:-r01s}
set -x
timeout 200 python bench.py --no-cpu > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; echo rc=$?
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 > $out/${tag}_bench_n8.json 2> $out/${tag}_bench_n8.err; echo rc=$?
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --workload c5 > $out/${tag}_c5_n8.json 2> $out/${tag}_c5_n8.err; echo rc=$?
nproc > $out/${tag}_host.txt; free -g >> $out/${tag}_host.txt
tail -c 300 $out/${tag}_bench_n8.err; tail -c 300 $out/${tag}_c5_n8.err
