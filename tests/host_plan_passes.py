"""Test infrastructure: run the two device-only plan passes of the fused path on the CPU.

``csrc/skb_p1_plan.cu`` holds two kernels that rewrite the tile records in place (bank spreading
of the P2 lists, conflict-free tile-local vertex ids).  Both are scalar per-thread code - one
thread per half-group / per tile, no warp intrinsics, no shared memory - so the *same source* can
be compiled by g++: this module cuts the kernels out of the .cu file (everything except the
``extern "C"`` launchers), puts a few-line shim in front (``__global__`` -> nothing,
``blockIdx`` / ``threadIdx`` as plain variables), adds host drivers that loop over the thread
ids, and builds ``oracle/_build/libp1_plan_host.so``.  Nothing here is product code and nothing
in the product calls it; the CPU tests use it to check that a plan is still a correct plan after
the two passes (tests/test_fused_plan_cpu.py).  ``p1_combine_kernel`` - the second kernel of the
warm step, a scalar grid-stride loop in csrc/skb_p1_fused.cu - is cut out and compiled the same
way (``combine``), so the emulation's last stage runs the shipped source too.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "scikit-fem_b200", "csrc", "skb_p1_plan.cu")
SRC_FUSED = os.path.join(ROOT, "scikit-fem_b200", "csrc", "skb_p1_fused.cu")
OUT_DIR = os.path.join(ROOT, "oracle", "_build")
LIB = os.path.join(OUT_DIR, "libp1_plan_host.so")

SHIM = r"""
#include <cstdint>
#include <cstring>
#include <algorithm>
#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__ __restrict
using std::min;
using std::max;
struct idx3 { unsigned x, y, z; };
static idx3 blockIdx = {0, 0, 0}, threadIdx = {0, 0, 0}, blockDim = {1, 1, 1},
            gridDim = {1, 1, 1};
"""

DRIVERS = r"""
extern "C" void host_plan_spread(uint16_t *rec16, const int64_t *grp_pos, const int32_t *grp_len,
                                 int64_t ngroups, int zero_base) {
  for (int64_t t = 0; t < 2 * ngroups; ++t) {
    blockIdx.x = (unsigned)t;
    skb::p1_plan_spread_kernel(rec16, grp_pos, grp_len, ngroups, zero_base);
  }
}
extern "C" void host_plan_renumber(unsigned char *rec, const uint64_t *rec_start, int ntiles,
                                   int tile_elems) {
  for (int t = 0; t < ntiles; ++t) {
    blockIdx.x = (unsigned)t;
    skb::p1_plan_renumber_kernel(rec, rec_start, ntiles, tile_elems);
  }
}
extern "C" void host_p1_combine(const double *scratch, const uint32_t *sptr, const uint32_t *gslot,
                                const uint32_t *gslot2, int64_t nshared, double *csr_data) {
  blockIdx.x = 0;                       // grid-stride loop: one "thread" walks every slot
  skb::p1_combine_kernel(scratch, sptr, gslot, gslot2, nshared, csr_data);
}
"""


def _host_source():
    src = open(SRC).read()
    src = src.replace('#include "skb_common.cuh"', "")
    # drop the extern "C" launchers (<<<...>>> is not C++): from the keyword to the closing
    # brace in column 0
    src = re.sub(r'extern "C"[^\n]*\n(?:.*\n)*?\}\n', "", src)
    # the second kernel of the warm step, cut out of the fused kernel's file (the fused kernel
    # itself is TMA / mbarrier / cp.async PTX and stays GPU-only)
    fused = open(SRC_FUSED).read()
    m = re.search(r'__global__ void __launch_bounds__\(256\)\np1_combine_kernel\((?:.*\n)*?\}\n',
                  fused)
    combine = "namespace skb {\n" + m.group(0) + "}\n"
    return SHIM + src + combine + DRIVERS


def build():
    os.makedirs(OUT_DIR, exist_ok=True)
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= max(
            os.path.getmtime(SRC), os.path.getmtime(SRC_FUSED), os.path.getmtime(__file__)):
        return LIB
    cpp = os.path.join(OUT_DIR, "p1_plan_host.cpp")
    with open(cpp, "w") as f:
        f.write(_host_source())
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas",
                    "-o", LIB, cpp], check=True, capture_output=True)
    return LIB


def apply(fp, T, spread=True, renumber=True):
    """Run the plan passes in place on the records of a plan built on CPU tensors."""
    lib = C.CDLL(build())
    rec = fp.rec.numpy()                       # int32 view of the record blob, shared memory
    rs = np.ascontiguousarray(fp.rec_start.numpy().astype(np.uint64))
    if spread:
        pos, length = [], []
        for tile in range(fp.ntiles):
            base = int(rs[tile])
            hdr = [int(v) & 0xFFFFFFFF for v in rec[base // 4: base // 4 + 8]]
            ngroups, off_grp, off_ids = hdr[1], hdr[3], hdr[5]
            grp = rec[(base + off_grp) // 4: (base + off_grp) // 4 + ngroups].view(np.uint32)
            for g in grp:
                pos.append((base + off_ids) // 2 + int(g & 0xFFFF) * 32)
                length.append(int(g >> 16))
        pos = np.asarray(pos, dtype=np.int64)
        length = np.asarray(length, dtype=np.int32)
        lib.host_plan_spread(C.c_void_p(rec.ctypes.data), C.c_void_p(pos.ctypes.data),
                             C.c_void_p(length.ctypes.data), C.c_int64(len(pos)),
                             C.c_int(10 * T))
    if renumber:
        lib.host_plan_renumber(C.c_void_p(rec.ctypes.data), C.c_void_p(rs.ctypes.data),
                               C.c_int(fp.ntiles), C.c_int(T))
    return fp


def apply2(fp, T, spread=True, renumber=True):
    """Same passes on a second-generation plan (skfem_b200/fused2.py) built with
    ``defer_finalize=True``: bank spreading works on the column-major ELL array ``fp._ell``,
    vertex renumbering on the records (whose header keeps nverts / off_verts in place)."""
    lib = C.CDLL(build())
    if spread and fp._ids_base16 is not None:
        ell = fp._ell.numpy()
        pos = np.ascontiguousarray(fp._gcum.numpy()[:-1].astype(np.int64))
        length = np.ascontiguousarray(fp._grp_len.numpy().astype(np.int32))
        lib.host_plan_spread(C.c_void_p(ell.ctypes.data), C.c_void_p(pos.ctypes.data),
                             C.c_void_p(length.ctypes.data), C.c_int64(len(pos)),
                             C.c_int(10 * T))
    if renumber:
        rec = fp.rec.numpy()
        rs = np.ascontiguousarray(fp.rec_start.numpy().astype(np.uint64))
        lib.host_plan_renumber(C.c_void_p(rec.ctypes.data), C.c_void_p(rs.ctypes.data),
                               C.c_int(fp.ntiles), C.c_int(T))
    return fp


def combine(fp, scratch, csr):
    """p1_combine_kernel (csrc/skb_p1_fused.cu) on the host: adds the per-tile partials in
    ``scratch`` into ``csr`` (both float64 numpy arrays) following the plan's sptr / gslot."""
    lib = C.CDLL(build())
    sptr, gslot, gslot2 = (np.ascontiguousarray(x.numpy()) for x in (fp.sptr, fp.gslot, fp.gslot2))
    lib.host_p1_combine(C.c_void_p(scratch.ctypes.data), C.c_void_p(sptr.ctypes.data),
                        C.c_void_p(gslot.ctypes.data), C.c_void_p(gslot2.ctypes.data),
                        C.c_int64(fp.nshared), C.c_void_p(csr.ctypes.data))
    return csr
