"""Test infrastructure: the FacetBasis kernels (csrc/skb_facet.cu) on the host - scalar
grid-stride kernels (one thread per facet / per facet quadrature point), compiled by g++ from
the shipped source with the shim of tests/host_local.py.  Not product code."""
import ctypes as C
import os
import subprocess

import host_local

SRC = os.path.join(host_local.CSRC, "skb_facet.cu")
LIB = os.path.join(host_local.OUT_DIR, "libfacet_host.so")

DRIVERS = r"""
using namespace skb;
extern "C" void host_facet_geometry(const skb_space_t *s, const int32_t *facets, int64_t nft,
                                    const int32_t *find, const int32_t *tind, const int32_t *tind_n,
                                    const int32_t *lfacet, int64_t nf, const double *Xb,
                                    const double *Wb, int nqp, double *x, double *Y, double *dx,
                                    double *nrm, double *detabs) {
  if (s->dim == 2)
    facet_geometry_kernel<2>(*s, facets, nft, find, tind, tind_n, lfacet, nf, Xb, Wb, nqp, x, Y, dx,
                             nrm, detabs);
  else
    facet_geometry_kernel<3>(*s, facets, nft, find, tind, tind_n, lfacet, nf, Xb, Wb, nqp, x, Y, dx,
                             nrm, detabs);
}
extern "C" void host_facet_basis(const skb_space_t *s, const int32_t *tind, int64_t nf, int nqp,
                                 const double *Y, const double *coef, const int32_t *expo,
                                 const int32_t *nterm, int b, double *value, double *grad) {
  if (s->dim == 2) facet_basis_kernel<2>(*s, tind, nf, nqp, Y, coef, expo, nterm, b, value, grad);
  else facet_basis_kernel<3>(*s, tind, nf, nqp, Y, coef, expo, nterm, b, value, grad);
}
"""


def _host_source():
    body = open(SRC).read().split('#include "skb_common.cuh"', 1)[1]
    body = body[:body.index('extern "C"')]
    open_ns = body.count("namespace skb {") - body.count("}  // namespace skb")
    return host_local.PRELUDE % {"hdr": host_local.HDR} + body + "}\n" * open_ns + DRIVERS


def build():
    os.makedirs(host_local.OUT_DIR, exist_ok=True)
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= max(
            os.path.getmtime(SRC), os.path.getmtime(host_local.HDR), os.path.getmtime(__file__),
            os.path.getmtime(host_local.__file__)):
        return LIB
    cpp = os.path.join(host_local.OUT_DIR, "facet_host.cpp")
    with open(cpp, "w") as f:
        f.write(_host_source())
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC",
                    "-Wno-unknown-pragmas", "-I", cuda_inc, "-o", LIB, cpp],
                   check=True, capture_output=True)
    return LIB


def lib():
    return C.CDLL(build())
