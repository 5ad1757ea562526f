#!/bin/bash
# `ncu --set full` of the first launch of every kernel of the library (tools/kernel_tour.py)
RX='^(condense_|counts_kernel|csr_reduce|element_dofs|enforce_|entity_|facet_|finalize_kernel|head_flags|interface_|local_|make_keys|mapping_|mesh_|p1_|p1tet_|qp_reduce|rows_|scan_|slot_of|spmv_|tabulate_|vec_reduce)'
mkdir -p gpurun_out
timeout ${2:-300} ncu --set full --clock-control none --kernel-id "::regex:$RX:1" -f \
  -o gpurun_out/prof_r2_tour python tools/kernel_tour.py ${1:-41} > gpurun_out/tour.log 2>&1
tail -3 gpurun_out/tour.log
ls -la gpurun_out/prof_r2_tour.ncu-rep
