#!/usr/bin/env bash
# TEST INFRASTRUCTURE: builds the reference's own CPU implementation of the GNN hot path into oracle/_ref/.
#
# Sources are compiled from where they lie under $REF (/root/reference, read-only): a scratch copy is made
# under $SCRATCH (default /tmp) only because four one-line patches are needed for the CPU path to build
# and run; no reference source is written into this repository — only binaries land in
# oracle/_ref/ (git-ignored, shipped to the GPU box with the snapshot).
#   patch 1  include/gnn/global.h:61,63   drop `#define ENABLE_GPU` / `#define USE_GGNN` (selects the CPU twins)
#   patch 2  src/gnn/net.cpp:150-154      drop the debug printf loop + exit(0) left in load_data
#   patch 3  src/gnn/net.cpp:620          drop `template class Model<GGNN_layer>;` (GGNN has no CPU twin)
#   patch 4  -include cstdint -include unistd.h (GCC 13 no longer provides uint8_t/getpid transitively)
#   patch 5  train.cpp: call print_timers() before main returns (defined at train.cpp:60 but never called)
# Third-party: cblas_sgemm comes from OpenBLAS 0.3.15 (the LP64 build bundled with opencv_python_headless in
# this image; the reference does not pin a BLAS version, src/gnn/Makefile:11,53). Boost is replaced by the
# std-based stand-ins in oracle/shims/boost (only dropout uses it; every config runs dropout 0).
# Flags follow src/gnn/Makefile:10,23 (-O3 -fopenmp -std=c++11, no -march=native).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
SCRATCH="${SCRATCH:-/tmp/gai_ref_build}"
if [ ! -d "$REF/src/gnn" ]; then
  echo "build_ref.sh: $REF not present; keeping prebuilt oracle/_ref as is"; exit 0
fi
PY="${PYTHON:-python}"
OBL_DIR="$($PY - <<'EOF'
import glob, site
for p in site.getsitepackages():
    g = glob.glob(p + "/opencv_python_headless.libs/libopenblasp-r0-*.so")
    if g:
        print(g[0]); break
EOF
)"
if [ -z "$OBL_DIR" ]; then echo "build_ref.sh: no OpenBLAS found (unbuildable)"; exit 1; fi
OBL_LIBDIR="$(dirname "$OBL_DIR")"
rm -rf "$SCRATCH"; mkdir -p "$SCRATCH" "$OUT"
cp -r "$REF/include" "$SCRATCH/include"
mkdir -p "$SCRATCH/src"
cp -r "$REF/src/gnn" "$REF/src/layers" "$REF/src/utilities" "$REF/src/common" "$REF/src/partitioner" "$SCRATCH/src/"
R="$SCRATCH"
sed -i 's/^#define ENABLE_GPU$/\/\/&/; s/^#define USE_GGNN$/\/\/&/' "$R/include/gnn/global.h"
sed -i '150,154d' "$R/src/gnn/net.cpp"
sed -i '/template class Model<GGNN_layer>;/d' "$R/src/gnn/net.cpp"
grep -q 'exit(0)' "$R/src/gnn/net.cpp" && { echo "patch 2 failed"; exit 1; }
sed -i 's/^  std::cout << "Test accuracy: ".*$/&\n  print_timers();/' "$R/src/gnn/train.cpp"

FL="-fopenmp -pthread -O3 -std=c++11 -w -fPIC -include cstdint -include unistd.h"
INC="-I$HERE/shims -I$R/include -I$R/include/gnn -I$R/include/layers -I$R/include/utils"
SRCS="$R/src/layers/softmax_loss_layer.cpp $R/src/layers/sigmoid_loss_layer.cpp $R/src/layers/l2norm_layer.cpp $R/src/layers/dense_layer.cpp \
 $R/src/gnn/gconv/gcn_layer.cpp $R/src/gnn/gconv/gcn_aggregator.cpp $R/src/gnn/gconv/sage_layer.cpp $R/src/gnn/gconv/sage_aggregator.cpp \
 $R/src/gnn/gconv/gat_layer.cpp $R/src/gnn/gconv/gat_aggregator.cpp \
 $R/src/gnn/graph_conv_layer.cpp $R/src/gnn/lgraph.cpp $R/src/gnn/reader.cpp $R/src/gnn/loss_layer.cpp $R/src/gnn/sampler.cpp \
 $R/src/utilities/random.cpp $R/src/utilities/math_functions.cpp $R/src/utilities/optimizer.cpp"
LINK="$OBL_DIR -Wl,--disable-new-dtags,-rpath,$OBL_LIBDIR"

build_arch() {  # $1 = name, $2 = macro
  local od="$SCRATCH/obj_$1"; mkdir -p "$od"
  local objs=""
  for s in $SRCS $R/src/gnn/net.cpp; do
    local o="$od/$(basename "$s" .cpp).o"
    g++ -c $FL $2 $INC "$s" -o "$o" &
    objs="$objs $o"
  done
  wait
  g++ $FL $2 $INC "$R/src/gnn/train.cpp" $objs $LINK -o "$OUT/cpu_train_$1"
  echo "$objs"
}
OBJS_GCN="$(build_arch gcn "")"
build_arch sage "-DUSE_SAGE" >/dev/null
build_arch gat "-DUSE_GAT" >/dev/null
# harness: reference objects (arch-macro-free TUs + net.o, whose template instantiations cover all three layer types)
g++ -shared $FL -fno-access-control $INC "$HERE/ref_harness.cpp" $OBJS_GCN $LINK -o "$OUT/libref_gnn.so"

# partitioner (no external deps): src/common/{graph,VertexSet}.cc + src/partitioner/graph_partition.cc
if [ -f "$HERE/ref_part_harness.cpp" ]; then
  g++ -shared -O3 -fopenmp -std=c++17 -w -fPIC -fno-access-control -I"$R/include" \
    "$R/src/common/VertexSet.cc" "$R/src/common/graph.cc" "$R/src/partitioner/graph_partition.cc" \
    "$HERE/ref_part_harness.cpp" -o "$OUT/libref_part.so"
fi
# text -> CSR converter (src/converters/converter.cc + src/common/{graph,VertexSet}.cc, no external deps) behind oracle/ref_conv_harness.cpp.
#   patch 6  converter.cc: `new Graph(nv, ne)` with two uint64_t arguments is ambiguous between GraphT(vidType, eidType) and
#            GraphT(bool, bool) under g++ 13 (4 call sites) -> explicit casts to the intended (vidType, eidType) overload
if [ -f "$HERE/ref_conv_harness.cpp" ]; then
  mkdir -p "$R/src/converters"
  cp "$REF/src/converters/converter.cc" "$REF/src/converters/converter.h" "$R/src/converters/"
  sed -i 's/new Graph(nv, ne)/new Graph((vidType)nv, (eidType)ne)/' "$R/src/converters/converter.cc"
  g++ -O2 -fopenmp -std=c++17 -w -I"$R/include" -I"$R/src/converters" \
    "$R/src/common/VertexSet.cc" "$R/src/common/graph.cc" "$R/src/converters/converter.cc" "$HERE/ref_conv_harness.cpp" -o "$OUT/ref_convert"
fi
rm -rf "$SCRATCH"
ls -la "$OUT"
