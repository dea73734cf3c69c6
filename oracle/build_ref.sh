#!/usr/bin/env bash
# oracle/build_ref.sh — TEST INFRASTRUCTURE.
# Compiles the REFERENCE's own sources, unmodified and where they lie under /root/reference,
# into oracle/_ref/ (git-ignored; travels to the GPU box with the gpurun snapshot):
#   oracle/_ref/pointops_cuda.so  the reference's pointops CUDA extension for sm_100a
#                                 (pytorch/lib/pointops/src/**; flags of its setup.py:16-31: nvcc -O2)
#                                 — the GPU oracle and the "stock pointops" baseline B1.
#   oracle/_ref/libref_cpu.so     reference TF-side C++ cores (compile_op.sh:26-31 flags: -std=c++11 -O2)
#   oracle/_ref/libref_cpy.so     reference CPython-flavour grid_subsampling core
# The only shim on an include path is an EMPTY THC/THC.h (removed from torch >= 1.11; unused).
# No reference source is copied into the repo.  Does nothing if /root/reference is absent.
set -euo pipefail
cd "$(dirname "$0")"
REF=${REF_ROOT:-/root/reference}
if [ ! -d "$REF" ]; then echo "[build_ref] $REF absent — keeping prebuilt oracle/_ref"; exit 0; fi
OUT=_ref
mkdir -p $OUT/obj $OUT/stub/THC
: > $OUT/stub/THC/THC.h

# ---- CPU cores -------------------------------------------------------------------------------
R=$REF/tensorflow/ops/tf_custom_ops
if [ ! -f $OUT/libref_cpu.so ] || [ ref_shim/ref_cpu_shim.cpp -nt $OUT/libref_cpu.so ]; then
  g++ -std=c++11 -O2 -fPIC -shared -I$R -include cstring ref_shim/ref_cpu_shim.cpp \
      $R/tf_neighbors/neighbors/neighbors.cpp $R/tf_subsampling/grid_subsampling/grid_subsampling.cpp \
      $R/cpp_utils/cloud/cloud.cpp -o $OUT/libref_cpu.so
fi
W=$REF/tensorflow/ops/cpp_wrappers
if [ ! -f $OUT/libref_cpy.so ] || [ ref_shim/ref_cpy_shim.cpp -nt $OUT/libref_cpy.so ]; then
  g++ -std=c++11 -O2 -fPIC -shared -I$W -include cstring ref_shim/ref_cpy_shim.cpp \
      $W/cpp_subsampling/grid_subsampling/grid_subsampling.cpp $W/cpp_utils/cloud/cloud.cpp -o $OUT/libref_cpy.so
fi

# ---- GPU pointops ----------------------------------------------------------------------------
if [ ! -f $OUT/pointops_cuda.so ]; then
  P=$REF/pytorch/lib/pointops/src
  PY=${PYTHON:-python}
  TI=$($PY -c "import torch.utils.cpp_extension as c; print(' '.join('-I'+p for p in c.include_paths()))")
  PYINC=$($PY -c "import sysconfig;print(sysconfig.get_paths()['include'])")
  TL=$($PY -c "import torch,os;print(os.path.join(os.path.dirname(torch.__file__),'lib'))")
  ABI=$($PY -c "import torch;print(int(torch._C._GLIBCXX_USE_CXX11_ABI))")
  for k in knnquery sampling grouping interpolation subtraction aggregation; do
    nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a $TI -I$PYINC -Xcompiler -fPIC \
         -D_GLIBCXX_USE_CXX11_ABI=$ABI -c $P/$k/${k}_cuda_kernel.cu -o $OUT/obj/${k}_k.o &
    g++ -O2 -std=c++17 -DTORCH_EXTENSION_NAME=pointops_cuda -DTORCH_API_INCLUDE_EXTENSION_H -I$OUT/stub $TI -I$PYINC \
         -D_GLIBCXX_USE_CXX11_ABI=$ABI -I/usr/local/cuda/include -fPIC -c $P/$k/${k}_cuda.cpp -o $OUT/obj/${k}_h.o &
  done
  g++ -O2 -std=c++17 -DTORCH_EXTENSION_NAME=pointops_cuda -DTORCH_API_INCLUDE_EXTENSION_H -I$OUT/stub $TI -I$PYINC \
      -D_GLIBCXX_USE_CXX11_ABI=$ABI -I/usr/local/cuda/include -fPIC -c $P/pointops_api.cpp -o $OUT/obj/api_h.o &
  wait
  g++ -shared -o $OUT/pointops_cuda.so $OUT/obj/*_h.o $OUT/obj/*_k.o -L$TL -ltorch -ltorch_cpu -ltorch_python -lc10 \
      -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$TL
fi

# ---- drop-in proof: the reference's UNMODIFIED pybind glue linked against libcbops.so ---------------------
# (the ten *_cuda_launcher symbols come from contrastboundary_b200/csrc/compat.cu instead of the reference's .cu files)
CB=../contrastboundary_b200
if [ -f $CB/libcbops.so ] && { [ ! -f $OUT/dropin/pointops_cuda.so ] || [ $CB/csrc/compat.cu -nt $OUT/dropin/pointops_cuda.so ]; }; then
  PY=${PYTHON:-python}
  TL=$($PY -c "import torch,os;print(os.path.join(os.path.dirname(torch.__file__),'lib'))")
  mkdir -p $OUT/dropin
  g++ -shared -o $OUT/dropin/pointops_cuda.so $OUT/obj/*_h.o -L$CB -l:libcbops.so -L$TL -ltorch -ltorch_cpu -ltorch_python -lc10 \
      -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$TL -Wl,-rpath,'$ORIGIN/../../../contrastboundary_b200'
fi

# ---- the reference's Python model code, for tests that run it on top of this repo's operators -------------
# copied (not committed: baseline/_ref is git-ignored) because /root/reference does not exist on the GPU box
B=../baseline/_ref/pytorch
if [ ! -f $B/model/blocks.py ]; then
  mkdir -p $B/lib/pointops/functions $B/util
  cp -r $REF/pytorch/model $B/
  cp $REF/pytorch/lib/pointops/functions/*.py $B/lib/pointops/functions/
  cp $REF/pytorch/util/config.py $REF/pytorch/util/voxelize.py $REF/pytorch/util/data_util.py $B/util/
  chmod -R u+w $B
  touch $B/lib/__init__.py $B/lib/pointops/__init__.py $B/util/__init__.py
fi
echo "[build_ref] ok: $(ls $OUT/*.so $OUT/dropin/*.so 2>/dev/null | tr '\n' ' ')"
