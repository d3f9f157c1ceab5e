#!/bin/bash
# Runs the CPU test suite of the oracle on an AddressSanitizer + UndefinedBehaviorSanitizer build of it.
#   bash scripts/oracle_sanitize.sh [pytest args]
# python itself is not instrumented, so libasan is preloaded and leak checking (python "leaks" by design) is off.
set -e
cd "$(dirname "$0")/.."
make -C oracle liboracle_asan.so
ASAN=$(/usr/bin/g++ -print-file-name=libasan.so)
UBSAN=$(/usr/bin/g++ -print-file-name=libubsan.so)
ORACLE_SANITIZE=1 LD_PRELOAD="$ASAN:$UBSAN" ASAN_OPTIONS=detect_leaks=0:abort_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
    python -m pytest tests/test_oracle_cloud.py tests/test_oracle_pipeline.py tests/test_oracle_smallmat.py tests/test_host_logic.py -q -m "not gpu" -p no:cacheprovider "$@"
