set -x
out=gpurun_out
timeout 300 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -k "sweep or multi_stream" > $out/r01e_tests2.log 2>&1; tail -3 $out/r01e_tests2.log
timeout 600 python bench.py > $out/r01e_bench_n1.json 2> $out/r01e_bench_n1.err; echo rc=$?
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $out/r01e_bench_n2.json 2> $out/r01e_bench_n2.err; echo rc=$?
timeout 600 python bench.py --workload c5 > $out/r01e_c5_n1.json 2> $out/r01e_c5_n1.err; echo rc=$?
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c5 > $out/r01e_c5_n2.json 2> $out/r01e_c5_n2.err; echo rc=$?
tail -c 800 $out/r01e_c5_n2.err
nvidia-smi topo -m > $out/r01e_topo.txt 2>&1
for f in /sys/bus/pci/devices/*/numa_node; do d=$(dirname $f); if [ "$(cat $d/vendor)" = "0x10de" ]; then echo $d $(cat $f) $(cat $d/class); fi; done >> $out/r01e_topo.txt
lscpu | grep -i "numa\|socket\|^CPU(s)" >> $out/r01e_topo.txt
