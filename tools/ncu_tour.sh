#!/bin/bash
# `ncu --set full` of the first launch of every kernel of the library (tools/kernel_tour.py)
RX='^(condense_|counts_kernel|csr_reduce|element_dofs|enforce_|entity_|facet_|finalize_kernel|head_flags|interface_|local_|make_keys|mapping_|mesh_|p1_|p1tet_|qp_reduce|rows_|scan_|slot_of|spmv_|tabulate_|vec_reduce)'
mkdir -p gpurun_out
timeout ${2:-300} ncu --set full --clock-control none --kernel-id "::regex:$RX:1" -f \
  -o gpurun_out/prof_r2_tour python tools/kernel_tour.py ${1:-41} > gpurun_out/tour.log 2>&1
tail -3 gpurun_out/tour.log
ls -la gpurun_out/prof_r2_tour.ncu-rep
# the table is made on the box: gpurun copies back at most 64 MiB and a full tour report exceeds
# that (75 MB with the round-2 kernels) - the report itself only travels when it is small enough
python tools/ncu_kernel_table.py gpurun_out/prof_r2_tour.ncu-rep gpurun_out/r2_ncu_all_kernels.md \
  "ncu evidence for every kernel of the library (one B200, tools/kernel_tour.py ${1:-41})"
if [ "$(stat -c %s gpurun_out/prof_r2_tour.ncu-rep)" -gt 50000000 ]; then rm gpurun_out/prof_r2_tour.ncu-rep; fi
