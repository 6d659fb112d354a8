# builds library variants: tools/dev/variants.sh name "-DX=..." ...
NVCC=/usr/local/cuda/bin/nvcc
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  $NVCC -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off -cudart static $flags -shared -o build/variants/$name.so libiqo_b200/csrc/plan.cpp libiqo_b200/csrc/kernels.cu libiqo_b200/csrc/capi.cu libiqo_b200/csrc/classes.cpp &
done
wait
ls -la build/variants
