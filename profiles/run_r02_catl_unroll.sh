mkdir -p gpurun_out
out=gpurun_out/r02_time_categorical_lpmf_unroll.txt; : > $out
run() { echo "== $*" | tee -a $out; env "$@" timeout 300 python profiles/time_categorical_lpmf.py 2>&1 | grep '"lin_var": true' | cut -c1-84 | tee -a $out; }
run MATH_B200_LIB=math_b200/lib/libstanmath_cuda_u2.so
run X=4
run MATH_B200_LIB=math_b200/lib/libstanmath_cuda_u8.so
run MATH_B200_LIB=math_b200/lib/libstanmath_cuda_u8.so SMC_CATL_W=16 SMC_CATL_S=2
