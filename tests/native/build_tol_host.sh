#!/bin/sh
# TEST INFRASTRUCTURE: builds tests/native/libtol_host.so (plain) and libtol_host_ulp.so (every emulated MUFU result
# moved by a pseudo-random -1/0/+1 ulp) from the __host__ __device__ tolerance-policy header rls_tol.cuh.
set -e
cd "$(dirname "$0")"
FLAGS="-O2 -std=c++14 -mfma -ffp-contract=fast -fno-fast-math -fopenmp -fPIC -shared -x c++"
g++ $FLAGS -o libtol_host.so tol_host.cpp
g++ $FLAGS -DRLS_TOL_EMULATE_ULP -o libtol_host_ulp.so tol_host.cpp
