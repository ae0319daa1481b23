#!/bin/bash
# Debug build of the C-ABI library with -DTQ_CONV_TRACE: the conv kernel's producer, MMA issuer and two epilogue groups
# account their cycles (clock64) per wait / phase and CTAs 0 and 100 print them at the end of every launch.
# Output: tools/bin/libtq_b200_trace.so (git-ignored); run anything with TQ_B200_LIB=tools/bin/libtq_b200_trace.so.
set -e
cd "$(dirname "$0")/../term_quantization_b200/csrc"
B=/tmp/tq_trace_build; mkdir -p $B ../../tools/bin
for f in tq_capi tq_encode tq_calib tq_gemm tq_pool tq_dw; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I../../include -I. \
      --expt-relaxed-constexpr -cudart static -DTQ_CONV_TRACE -c $f.cu -o $B/$f.o &
done
wait
for f in tq_capi tq_encode tq_calib tq_gemm tq_pool tq_dw; do test -f $B/$f.o; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../../tools/bin/libtq_b200_trace.so $B/*.o -lcuda
