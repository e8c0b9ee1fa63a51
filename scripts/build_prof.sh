#!/bin/sh
# instrumented build of the library (per-phase clock64 accounting in the QP kernel); not the product build
cd "$(dirname "$0")/../multiagent_planning_b200/csrc" && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC -shared -diag-suppress 177 -DDMPC_PROF -DDMPC_SINGLE_TU ${PROF_EXTRA} -o ../libdmpc_b200_prof.so dmpc_b200.cu model_tables.cpp
