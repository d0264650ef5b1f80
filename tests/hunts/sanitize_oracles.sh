#!/bin/bash
# TEST INFRASTRUCTURE (the `host` leg of tools/sanitize.sh): both CPU oracles built with -fsanitize=address,undefined and
# the oracle pinning + physics tests run on them.  CPU only.  Log: gpurun_out/<tag>_sanitizer_host_asan_ubsan.txt.
TAG=${1:-r02}
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
cd "$ROOT"
mkdir -p gpurun_out
LOG=gpurun_out/${TAG}_sanitizer_host_asan_ubsan.txt
: > $LOG
SAN="-fsanitize=address,undefined -fno-omit-frame-pointer -g"
mkdir -p build/san
# same sources, same pinned FP flags (oracle/Makefile), into build/san/ (the normal libraries are left alone)
FP="-O1 -ffp-contract=off -fno-fast-math -fopenmp -fPIC -DNDEBUG"
gcc -std=c11 $FP $SAN -shared -o build/san/librls_oracle.so oracle/rls_oracle.c -lm >> $LOG 2>&1
if [ -f /root/reference/src/rlGgx.h ]; then
  g++ -std=c++14 -D_LINUX $FP $SAN -Ioracle/shim -I/root/reference/src -Ioracle -w -shared -o build/san/librls_ref.so \
      oracle/ref_driver.cpp oracle/shim/shim_stubs.cpp /root/reference/src/rlUtil.cpp /root/reference/src/rlGgx.cpp \
      /root/reference/src/rlSss.cpp -lm >> $LOG 2>&1
fi
ASAN_LIB=$(gcc -print-file-name=libasan.so)
echo "### pytest tests/test_oracle_pinning.py tests/test_oracle_physics.py on the sanitized oracles" >> $LOG
RLS_ORACLE_DIR=$ROOT/build/san LD_PRELOAD=$ASAN_LIB ASAN_OPTIONS=detect_leaks=0:halt_on_error=0 UBSAN_OPTIONS=print_stacktrace=1 \
    python -m pytest -q tests/test_oracle_pinning.py tests/test_oracle_physics.py >> $LOG 2>&1; echo "rc=$?" >> $LOG
echo "asan/ubsan: $(grep -c 'runtime error' $LOG) UBSan reports, $(grep -c 'ERROR: AddressSanitizer' $LOG) ASan reports; $(tail -3 $LOG | tr '\n' ' ')"
