#!/bin/bash
# AddressSanitizer + UndefinedBehaviorSanitizer over the kernel bodies (csrc/wq_kernels.cuh, wq_align.cuh, wq_bias.cuh: __host__ __device__
# code driven single-threaded by tests/emu/wq_emu.cpp) on the CPU.  compute-sanitizer is closed on this GPU pool (profiles/r02_sanitizer.md),
# so this is the memory-safety evidence for the code the kernels share with the harness; races are not covered by it.
set -e
cd "$(dirname "$0")/.."
SO=tests/emu/libwq_emu.so
cp $SO /tmp/libwq_emu.so.keep 2>/dev/null || true
/usr/bin/g++ -O1 -g -std=c++17 -fPIC -shared -fsanitize=address,undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer -Wno-unknown-pragmas \
    -ffp-contract=off -o $SO tests/emu/wq_emu.cpp
touch $SO
echo "instrumented: $(nm -D $SO | grep -c __asan_) __asan_ symbols, $(nm -D $SO | grep -c __ubsan_) __ubsan_ symbols in $SO"
ASAN_OPTIONS=detect_leaks=0:abort_on_error=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1 LD_PRELOAD=$(/usr/bin/gcc -print-file-name=libasan.so):$(/usr/bin/gcc -print-file-name=libubsan.so) \
    python -m pytest tests/test_kernel_logic_cpu.py "tests/test_bias.py::test_kernel_bodies_on_cpu" -q -x -p no:cacheprovider 2>&1 | tail -15
rc=${PIPESTATUS[0]}
rm -f $SO
cp /tmp/libwq_emu.so.keep $SO 2>/dev/null || true
# the host stages (class building, class shares, assign_ambig!, event construction and cuts, SAM formatter, index re-layout) and the phases of the
# cluster PSI kernel run block by block on heap buffers of exactly the shared-memory size the host computes: tests/emu/wq_host_emu.cu
HSO=tests/emu/libwq_host_emu.so
cp $HSO /tmp/libwq_host_emu.so.keep 2>/dev/null || true
nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -fmad=false -Xcompiler -fPIC,-fsanitize=address,-fsanitize=undefined,-fno-sanitize-recover=undefined,-fno-omit-frame-pointer \
    -shared -o $HSO tests/emu/wq_host_emu.cu -lasan -lubsan -lz
touch $HSO
echo "instrumented: $(nm -D $HSO | grep -c __asan_) __asan_ symbols, $(nm -D $HSO | grep -c __ubsan_) __ubsan_ symbols in $HSO"
ASAN_OPTIONS=detect_leaks=0:abort_on_error=0:halt_on_error=1:protect_shadow_gap=0 UBSAN_OPTIONS=print_stacktrace=1 LD_PRELOAD=$(/usr/bin/gcc -print-file-name=libasan.so):$(/usr/bin/gcc -print-file-name=libubsan.so) \
    python -m pytest tests/test_host_stages_cpu.py tests/test_events_flat.py tests/test_psi_cluster_cpu.py tests/test_inflate_cpu.py "tests/test_bias.py::test_host_stages_with_bias_on_cpu" "tests/test_sam.py::test_sam_formatter_on_cpu" -q -x -p no:cacheprovider 2>&1 | tail -15
rc2=${PIPESTATUS[0]}
rm -f $HSO
cp /tmp/libwq_host_emu.so.keep $HSO 2>/dev/null || true
exit $((rc + rc2))
