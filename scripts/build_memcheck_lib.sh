# Build pyfe3d_b200/lib/variants/memcheck/libpyfe3d_b200.so = the library with -DPF3_MEMCHECK_BUILD (common.cuh: prop_index),
# then rebuild the default library.  scripts/gpu_sanitizer.sh runs compute-sanitizer against that variant.
set -e
cd "$(dirname "$0")/.."
PF3_EXTRA_NVCC_FLAGS="-DPF3_MEMCHECK_BUILD" python -c "from pyfe3d_b200 import build as B; B.build_lib(force=True, verbose=False)"
mkdir -p pyfe3d_b200/lib/variants/memcheck
cp pyfe3d_b200/lib/libpyfe3d_b200.so pyfe3d_b200/lib/variants/memcheck/
python -c "from pyfe3d_b200 import build as B; B.build_lib(force=True, verbose=False)"
echo "built pyfe3d_b200/lib/variants/memcheck/libpyfe3d_b200.so; default library rebuilt"
