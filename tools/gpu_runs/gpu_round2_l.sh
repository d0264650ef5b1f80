#!/bin/bash
mkdir -p gpurun_out
python tools/tol_flag_hunt.py 24 > gpurun_out/r02l_flag_hunt.log 2>&1; echo "hunt rc=$?"; grep -v "mismatches 0" gpurun_out/r02l_flag_hunt.log | tail -3
bash tools/final_profile.sh r02
