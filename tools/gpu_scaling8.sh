# 1 -> N GPUs of one box, same code, back to back (what the driver does at round end): the N=1 and N=8 bench lines of this
# repo's arm and the N=8 reference arm (rank 0 alone: one single-thread reference process per host core).
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_scaling8.sh r02p 8'
out=gpurun_out
tag=${1:-r02p}
n=${2:-8}
set -x
timeout 240 python bench.py --no-cpu --no-sublegs > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; echo rc=$?
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $n --no-sublegs > $out/${tag}_bench_n$n.json 2> $out/${tag}_bench_n$n.err; echo rc=$?
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $n --steps 20 --warmup 5 > $out/${tag}_ref_n$n.json 2> $out/${tag}_ref_n$n.err; echo rc=$?
nproc > $out/${tag}_host.txt; free -g >> $out/${tag}_host.txt
tail -c 300 $out/${tag}_bench_n$n.err; tail -c 300 $out/${tag}_ref_n$n.err
