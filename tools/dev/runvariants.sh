# runs the bench pair for every library variant on this box (restores the product library at the end)
cp libiqo_b200/lib/libiqo_cuda.so build/variants/_orig.so
for v in "$@"; do cp build/variants/$v.so libiqo_b200/lib/libiqo_cuda.so; echo "variant $v"; bash tools/dev/runbench.sh; done
cp build/variants/_orig.so libiqo_b200/lib/libiqo_cuda.so
