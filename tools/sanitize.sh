#!/bin/bash
# Sanitizer evidence.
#   GPU box:   tools/sanitize.sh gpu [tag]   compute-sanitizer memcheck / racecheck / initcheck / synccheck over the C++ host
#              driver (every fused kernel under the three arithmetic policies, the host-staged pipeline with its
#              three copy streams, the compact-frame decode, the sweep + re-run lists) and over the experiments library's
#              shared-memory kernels (TMA / mbarrier pipe, CTA-level lobe partition) through pytest at reduced sizes.
#   anywhere:  tools/sanitize.sh host [tag]  = tests/hunts/sanitize_oracles.sh: ASan/UBSan builds of both oracles + the pinning and
#              physics tests on them (CPU only).
# Logs: gpurun_out/<tag>_sanitizer_*.txt (copy the summaries to profiles/).
MODE=${1:-gpu}
TAG=${2:-r02}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
cd "$ROOT"
mkdir -p gpurun_out
if [ "$MODE" = gpu ]; then
  D=rlshaders_b200/host/rls_driver
  CS="compute-sanitizer --error-exitcode 9 --print-limit 20"
  for TOOL in memcheck racecheck initcheck synccheck; do
    LOG=gpurun_out/${TAG}_sanitizer_${TOOL}.txt
    : > $LOG
    run() { echo "### $*" >> $LOG; timeout 900 $CS --tool $TOOL "$@" >> $LOG 2>&1; echo "rc=$?" >> $LOG; }
    for W in ggx dielectric disney skin dielectric_q disney_q; do run $D $W 17; done                      # host-staged path, default policy
    for P in fast exact tolerant; do
      for W in dielectric disney skin; do run $D --gpus 1 --policy $P --reps 1 $W 17; done
      run $D --gpus 1 --policy $P --reps 1 sweep 6
    done
    # the experiments library (shared-memory permutation with two barriers, TMA bulk copies + mbarrier) and the compact
    # frames, through pytest at reduced sizes
    # (initcheck: without the TMA-staged kernel -- it reports that kernel's cp.async.bulk shared->global stores as never
    #  having initialised the output arrays, 20 reports per array, and then crawls; memcheck covers that kernel)
    SKIP="not packed"; [ $TOOL = initcheck ] && SKIP="not packed and not tma"
    echo "### pytest experiments + compact frames ($SKIP)" >> $LOG
    RLS_TEST_N=16384 timeout 900 $CS --tool $TOOL --target-processes all python -m pytest -q -x tests/test_experiments.py \
        "tests/test_gpu_parity.py::test_compact_frame_host_forms" -k "$SKIP" >> $LOG 2>&1; echo "rc=$?" >> $LOG
    echo "$TOOL: $(grep -c '^rc=0' $LOG) runs clean, $(grep -c '^rc=[1-9]' $LOG) with findings; $(grep -h 'ERROR SUMMARY' $LOG | sort | uniq -c | tr '\n' ';')"
  done
  # the C++ host driver itself under ASan + UBSan (host side of the staging pipeline, arenas, multi-device bookkeeping)
  LOG=gpurun_out/${TAG}_sanitizer_driver_asan_ubsan.txt
  : > $LOG
  g++ -O1 -g -std=c++14 -fsanitize=address,undefined -fno-omit-frame-pointer -o build/rls_driver_asan rlshaders_b200/host/rls_driver.cpp \
      -Lrlshaders_b200 -lrls_b200 -Wl,-rpath,$ROOT/rlshaders_b200 >> $LOG 2>&1
  export ASAN_OPTIONS=protect_shadow_gap=0:detect_leaks=0:halt_on_error=0 UBSAN_OPTIONS=print_stacktrace=1
  for W in ggx dielectric disney skin; do echo "### $W 18" >> $LOG; timeout 600 build/rls_driver_asan $W 18 >> $LOG 2>&1; echo "rc=$?" >> $LOG; done
  for P in fast tolerant; do
    for W in dielectric disney skin "sweep 6"; do echo "### --gpus 1 --policy $P $W" >> $LOG; timeout 600 build/rls_driver_asan --gpus 1 --policy $P --reps 2 $W 18 >> $LOG 2>&1; echo "rc=$?" >> $LOG; done
  done
  echo "driver asan/ubsan: $(grep -c 'runtime error' $LOG) UBSan reports, $(grep -c 'ERROR: AddressSanitizer' $LOG) ASan reports, $(grep -c '^rc=0' $LOG) runs clean, $(grep -c '^rc=[1-9]' $LOG) failed"
else
  exec bash "$ROOT/tests/hunts/sanitize_oracles.sh" "$TAG"
fi
