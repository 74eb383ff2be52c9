cp uav-airvision_b200/lib/libavb.so /tmp/libavb_orig.so
for B in 4 6 8; do
  cp gpurun_variants/libavb_b$B.so uav-airvision_b200/lib/libavb.so
  timeout 300 python bench.py --steps 30 --no-cpu > gpurun_out/r01h_var_b$B.json 2> gpurun_out/r01h_var_b$B.err; echo "B=$B rc=$?"
done
cp /tmp/libavb_orig.so uav-airvision_b200/lib/libavb.so
