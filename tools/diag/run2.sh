set -x
mkdir -p gpurun_out
timeout 300 python tools/diag/seal_diag.py > gpurun_out/seal_diag.log 2>&1; tail -40 gpurun_out/seal_diag.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py::test_seal_key_directory_round_trip > gpurun_out/r02_gpu_tests_rest.log 2>&1; tail -5 gpurun_out/r02_gpu_tests_rest.log
bash tools/resnet_knobs.sh > gpurun_out/resnet_knobs.log 2>&1; cat gpurun_out/resnet_knobs.log
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_batch_l3 -f python tools/batch_ncu_probe.py 3 32 rotate > gpurun_out/ncu_batch_l3.log 2>&1; grep "us per op" gpurun_out/ncu_batch_l3.log
HEVM_FUSED=0 timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_solo_l3 -f python tools/ks_probe.py 3 > gpurun_out/ncu_solo_l3.log 2>&1; tail -2 gpurun_out/ncu_solo_l3.log
