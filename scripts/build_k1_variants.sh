#!/bin/bash
# A/B builds of K1's row update (development aid): libnph_<name>.so under nanopolish_b200/csrc/build/variants/
set -e
cd "$(dirname "$0")/../nanopolish_b200/csrc"
SRCS="nph_api.cu hmm_schedule.cu hmm_forward.cu hmm_forward_w4.cu hmm_forward_w8.cu hmm_forward_w16.cu hmm_forward_w32.cu hmm_forward_w32c.cu hmm_viterbi.cu eventalign_chain.cu abea.cu event_detect.cu squiggle_prep.cu load_from_raw.cu methylation.cu variants.cu"
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC"
build() {   # name, extra flags
  local name=$1; shift
  mkdir -p build/v_$name build/variants
  for f in $SRCS; do
    case $f in hmm_forward*|hmm_schedule.cu) echo "nvcc $FLAGS $* -c $f -o build/v_$name/${f%.cu}.o";; *) echo "true";; esac
  done | xargs -P 8 -I{} sh -c "{}"
  objs=""
  for f in $SRCS; do
    case $f in hmm_forward*|hmm_schedule.cu) objs="$objs build/v_$name/${f%.cu}.o";; *) objs="$objs build/${f%.cu}.o";; esac
  done
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o build/variants/libnph_$name.so $objs
  echo "built $name"
}
if [ "$1" = "codes" ]; then
  build cur
  build nocodes -DNPH_NO_CODES
  exit 0
fi
build scalar_lea   -DNPH_PACKED_F32X2=0 -DNPH_EVEN_CELL_COST=88.0f -DNPH_LSUM_LEA
build scalar_imad  -DNPH_PACKED_F32X2=0 -DNPH_EVEN_CELL_COST=88.0f
build packed_all
build packed_all_lea -DNPH_LSUM_LEA
build packed_arith -DNPH_PACKED_LSUM=0
build packed_lsum  -DNPH_PACKED_ARITH=0
build packed_all_c9 -DNPH_EVEN_CELL_COST=88.0f
