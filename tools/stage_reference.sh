# Stage the reference's source tree under baseline/_ref (git-ignored, travels with gpurun) so that
# tests/test_gpu_reference_drivers.py can run the UNMODIFIED Main.py drivers on the GPU box:
#     bash tools/stage_reference.sh /root/reference && gpurun -- 'python -m pytest tests/test_gpu_reference_drivers.py -m gpu -q -s'
# Remove baseline/_ref afterwards; nothing else in the repo reads it.
set -e
src=${1:-/root/reference}
mkdir -p baseline/_ref
cp "$src"/*.py baseline/_ref/
ls baseline/_ref
