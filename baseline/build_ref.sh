#!/usr/bin/env bash
# Stage the UNMODIFIED reference (Felix-Petersen/gendr) into the git-ignored baseline/_ref/
# and build ONLY its generalized_renderer CUDA extension for sm_100a, so that the reference's
# own CUDA kernels can be run next to ours on the GPU box (parity + "reference_cuda" timing).
# Nothing from the reference is committed: baseline/_ref/ is listed in .gitignore.
# Usage: baseline/build_ref.sh [/root/reference]
set -euo pipefail
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/_ref"
if [ ! -d "$REF/gendr" ]; then echo "reference not found at $REF (expected on the build container only)"; exit 0; fi
rm -rf "$DST"; mkdir -p "$DST"
cp -r "$REF/gendr" "$DST/gendr"
mkdir -p "$DST/data"; cp "$REF/experiments/data/sphere_642.obj" "$DST/data/"
chmod -R u+w "$DST"
cd "$DST"
# a two-extension setup script (the reference's setup.py builds four; the renderer is the hot path, the voxelizer is
# SURVEY 8(f) row 4 -- the parity oracle for gendr_voxelize)
cat > setup_renderer_only.py <<'PY'
from setuptools import setup
from torch.utils.cpp_extension import BuildExtension, CUDAExtension
setup(name='gendr_ref_renderer',
      ext_modules=[CUDAExtension('gendr.cuda.generalized_renderer',
                                 ['gendr/cuda/generalized_renderer_cuda.cpp',
                                  'gendr/cuda/generalized_renderer_cuda_kernel.cu']),
                   CUDAExtension('gendr.cuda.voxelization',
                                 ['gendr/cuda/voxelization_cuda.cpp',
                                  'gendr/cuda/voxelization_cuda_kernel.cu'])],
      cmdclass={'build_ext': BuildExtension})
PY
TORCH_CUDA_ARCH_LIST="10.0a" MAX_JOBS=8 python setup_renderer_only.py build_ext --inplace > build.log 2>&1 || { tail -30 build.log; exit 1; }
# the two asset-I/O extensions are not built: empty stand-ins so `import gendr` works
for m in load_textures create_texture_image; do
  [ -f gendr/cuda/$m.py ] || echo "# stand-in: extension not built (off the hot path)" > gendr/cuda/$m.py
done
ls -la gendr/cuda/*.so
