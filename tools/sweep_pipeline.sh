export SONAR_BENCH_NO_CLOCKS=1
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/c5_noise_probe.py 2>&1 | head -8
SONAR_BENCH_ITEMS=1 python tools/c5_job_probe.py 2>&1 | tail -1
python tools/c5_job_probe.py 2>&1 | tail -1
python tools/step_probe.py 2>&1 | tail -3
