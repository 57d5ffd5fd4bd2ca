#!/bin/bash
# round 2 / session 25: full gpu tests; evidence for the new trunk kernels (ncu), per-class DRAM traffic of one step, layer table
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -5
timeout 900 python tools/bench_trunk.py gpurun_out/r02_trunk_layers.md > gpurun_out/s25_trunk.log 2>&1; tail -3 gpurun_out/s25_trunk.log
for k in stem_conv maxpool3x3s2 bn_fold2_fwd; do
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$k" -c 1 -o gpurun_out/r02_step_$k -f python tools/profile_step.py > gpurun_out/s25_ncu_$k.log 2>&1; echo "ncu $k exit $?"
done
# trunk GEMMs inside the step: the 20th..22nd gemm_kmajor launches of the step are layer-2 convolutions; the last ones trunk dgrads
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_kmajor -s 12 -c 1 -o gpurun_out/r02_step_trunk_conv_fwd -f python tools/profile_step.py > gpurun_out/s25_ncu_tf.log 2>&1; echo "ncu trunk fwd exit $?"
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_kmajor -s 340 -c 1 -o gpurun_out/r02_step_trunk_conv_dgrad -f python tools/profile_step.py > gpurun_out/s25_ncu_td.log 2>&1; echo "ncu trunk dgrad exit $?"
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_mnmajor -s 150 -c 1 -o gpurun_out/r02_step_trunk_conv_wgrad -f python tools/profile_step.py > gpurun_out/s25_ncu_tw.log 2>&1; echo "ncu trunk wgrad exit $?"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/step_traffic.csv python tools/profile_step.py > gpurun_out/s25_ncu_step.log 2>&1; echo "ncu step exit $?"; wc -l gpurun_out/step_traffic.csv
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s25_bench.json 2> gpurun_out/s25_bench.err; tail -c 2500 gpurun_out/s25_bench.json
ls -la gpurun_out/r02_step_*.ncu-rep | awk '{print $5, $9}' | tail -8
