K=k_skin_profile_tol_x2
timeout 400 ncu --set full --clock-control none -k "regex:^${K}\$" -s 1 -c 1 -f -o gpurun_out/r02_prof_$K python bench.py --steps 2 --warmup 3 --no-cpu --no-cpp-driver --e2e-steps 1 --e2e-samples 1048576 > /dev/null 2>&1
ncu -i gpurun_out/r02_prof_$K.ncu-rep --page raw --csv > gpurun_out/r02_prof_$K.raw.csv 2>/dev/null
rm -f gpurun_out/r02_prof_$K.ncu-rep; ls -la gpurun_out/r02_prof_$K.raw.csv
