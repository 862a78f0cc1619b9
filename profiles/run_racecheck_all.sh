mkdir -p gpurun_out
NV_COMPUTE_SANITIZER_MAX_RACECHECK_HAZARDS=1000000 timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --log-file gpurun_out/sanitizer_racecheck_all.log \
  python -m pytest -x -q -m gpu -p no:cacheprovider tests/test_glm_gpu.py -k "(normal or ordered or bernoulli) and not full_size and not 50021 and not 200003 and not fuzz" > gpurun_out/sanitizer_pytest_race.log 2>&1
echo rc=$?; tail -2 gpurun_out/sanitizer_pytest_race.log
grep -oE "[a-z_]+\.(cu|cuh):[0-9]+" gpurun_out/sanitizer_racecheck_all.log | sort | uniq -c
tail -2 gpurun_out/sanitizer_racecheck_all.log
