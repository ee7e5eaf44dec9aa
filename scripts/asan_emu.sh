# usage: bash scripts/asan_emu.sh   (here, no GPU) -- the kernels' memory safety under AddressSanitizer + UBSan: builds the host emulation of
# the library sources with -fsanitize=address,undefined and runs the emulation test-suites and scripts/sanitize_small.py --emu on it. Every "device"
# table is a host allocation in the emulation, so an out-of-bounds access of a kernel aborts at the instruction that makes it.
set -e
cd "$(dirname "$0")/.."
make -C tests/emu libmctsb200_emu_asan.so > /dev/null
export MCTSB200_EMU_LIB=$PWD/tests/emu/libmctsb200_emu_asan.so
ASAN=$(${ASAN_CXX:-/usr/bin/g++} -print-file-name=libasan.so)
export LD_PRELOAD=$(readlink -f "$ASAN") ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1
python -m pytest tests/test_host_emu_parity.py tests/test_api_mirror.py tests/test_distributed_gloo.py tests/test_fuzz_parity.py -x -q -m "not gpu" -p no:cacheprovider
python scripts/sanitize_small.py --emu "$MCTSB200_EMU_LIB"
