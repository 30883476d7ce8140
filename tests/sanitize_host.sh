#!/bin/bash
# Host-side sanitizer runs (CPU only): the shared host logic of form_b200/host/form/*.hpp (Estimator,
# fixed-lag smoother, key-scanner, trace, SE(3) / dense algebra) and the CPU oracle, compiled with
# ASan + UBSan and with TSan, driven by the CPU test files.  compute-sanitizer (the device side) is
# closed on the GPU pool of this project (profiles/r03/compute_sanitizer_closed.txt).
#   tests/sanitize_host.sh [outdir]     -> prints a summary; the log of the round is profiles/r05/host_sanitizers.txt
set -e
cd "$(dirname "$0")/.."
OUT=${1:-/tmp/form_sanitize}
mkdir -p "$OUT"
SRCS="oracle_extract.cpp oracle_map.cpp oracle_factor.cpp oracle_hotpath.cpp oracle_capi.cpp oracle_pipeline_capi.cpp oracle_smoother_capi.cpp"
# -Bsymbolic as in oracle/Makefile: the instrumented copies of the header-only host logic must be the ones that
# run (pytest loads the product's libformhost.so globally, whose uninstrumented copies would otherwise interpose)
FLAGS="-O1 -g -std=c++17 -fPIC -ffp-contract=off -pthread -DFORM_HOTPATH_INJECTED_ONLY -fno-omit-frame-pointer -I../include -I../form_b200/host -I. -shared -Wl,-Bsymbolic"
(cd oracle && g++ $FLAGS -fsanitize=address,undefined -o "$OUT/liboracle_asan.so" $SRCS)
(cd oracle && g++ $FLAGS -fsanitize=thread -o "$OUT/liboracle_tsan.so" $SRCS)
if [ -d /root/reference/form ]; then
  (cd oracle/ref && g++ -O1 -g -std=c++17 -fPIC -ffp-contract=off -w -shared -Wl,-Bsymbolic -fsanitize=address,undefined -fno-omit-frame-pointer \
     -I../shim -I/root/reference -I../../include -I.. -I../../form_b200/host ref_capi.cpp ref_estimator_capi.cpp \
     /root/reference/form/feature/factor.cpp /root/reference/form/mapping/keyscanner.cpp \
     /root/reference/form/optimization/constraints.cpp /root/reference/form/form.cpp ../oracle_extract.cpp -o "$OUT/libformref_asan.so")
  export FORM_REF_LIB="$OUT/libformref_asan.so"   # FORM's own code over the stand-ins, instrumented too
fi
echo "== ASan + UBSan (halt_on_error=1): host logic + oracle (+ FORM's own code over the stand-ins) through the CPU tests"
LD_PRELOAD=$(g++ -print-file-name=libasan.so):$(g++ -print-file-name=libubsan.so) \
  ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
  FORM_ORACLE_LIB="$OUT/liboracle_asan.so" \
  python -m pytest tests/test_pipeline_cpu.py tests/test_smoother_independent.py tests/test_keyscanner_pin.py \
    tests/test_golden.py tests/test_oracle_extract.py tests/test_oracle_map_factor.py tests/test_reference_pins.py \
    tests/test_reference_pipeline.py -q -m "not gpu" -p no:cacheprovider 2>&1 | tail -2
echo "== TSan: the pipeline with 4 worker threads (tests that vary the thread count)"
LD_PRELOAD=$(g++ -print-file-name=libtsan.so) TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0" \
  FORM_ORACLE_LIB="$OUT/liboracle_tsan.so" \
  python -m pytest tests/test_pipeline_cpu.py tests/test_golden.py -q -s -m "not gpu" -p no:cacheprovider > "$OUT/tsan.log" 2>&1 || true
tail -1 "$OUT/tsan.log"
echo "ThreadSanitizer reports: $(grep -c 'WARNING: ThreadSanitizer' "$OUT/tsan.log" || true)"
echo "== TSan: the threaded CPU leg of bench.py (one single-threaded estimator + replay per host core)"
LD_PRELOAD=$(g++ -print-file-name=libtsan.so) TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0" \
  FORM_ORACLE_LIB="$OUT/liboracle_tsan.so" \
  python bench.py --impl reference --sensor vlp-16 --steps 3 --warmup 3 --preroll 0 > "$OUT/tsan_bench.log" 2>&1 || true
echo "bench line printed: $(grep -c '"impl": "reference"' "$OUT/tsan_bench.log")   ThreadSanitizer reports: $(grep -c 'WARNING: ThreadSanitizer' "$OUT/tsan_bench.log" || true)"
