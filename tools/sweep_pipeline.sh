export SONAR_BENCH_NO_CLOCKS=1
python tools/wcfg_bench.py 2>&1 | tail -8
python -m pytest tests/test_gpu_wavelet.py tests/test_gpu_properties.py tests/test_torch_ops.py -x -q 2>&1 | tail -4
ncu --set full --clock-control none --import-source on -k regex:wcfg_strip -s 3 -c 1 -o gpurun_out/r02_c4_wcfg_strip3 python tools/wcfg_one.py > gpurun_out/r2_ncu7.log 2>&1; ncu -i gpurun_out/r02_c4_wcfg_strip3.ncu-rep --page source --csv > gpurun_out/r02_c4_wcfg_strip3_src.csv 2>/dev/null; ncu -i gpurun_out/r02_c4_wcfg_strip3.ncu-rep --page details > gpurun_out/r02_c4_wcfg_strip3_details.txt
