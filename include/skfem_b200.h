/* skfem_b200.h -- C ABI of the B200-native finite element assembly engine.
 *
 * Drop-in boundary for the scikit-fem (v12.0.1) assembly hot path
 *     CellBasis(mesh, elem) -> BilinearForm/LinearForm._assemble -> COOData -> CSR
 * The reference is pure Python and has no FFI; the seams these entry points
 * replace are the Python call signatures listed in SURVEY.md section 8(b).
 * Each entry point names the reference code (paths relative to
 * /root/reference/skfem/) whose arithmetic it reproduces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns and allocates every buffer (inputs, outputs, scratch);
 *     nothing is retained between calls; the only process-wide state is the
 *     launch counter, the profiling switches and the skb_sm_reserve knob;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*);
 *     calls return without synchronising unless documented otherwise;
 *   - return value: 0 on success, a positive cudaError_t, or a negative
 *     SKB_E* code for argument errors.  Nothing throws.
 *   - arithmetic is IEEE-754 binary64, round-to-nearest, NO fused
 *     multiply-add contraction, in the reference's operation order
 *     (SURVEY.md Appendix A), so element-local data is bit-identical to
 *     numpy's and the value-dependent CSR pattern matches scipy's.
 */
#ifndef SKFEM_B200_H
#define SKFEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKB_OK 0
#define SKB_EINVAL (-1)      /* bad argument / unsupported combination   */
#define SKB_ETOOBIG (-2)     /* tables do not fit on-chip / index overflow */
#define SKB_EZERODET (-3)    /* "Zero Jacobian determinant"               */

/* mapping kinds */
#define SKB_MAP_AFFINE 0     /* mapping/mapping_affine.py (tri, tet)      */
#define SKB_MAP_ISO_HEX1 1   /* mapping/mapping_isoparametric.py on MeshHex1 */

/* library bilinear forms (models/poisson.py:7-19, models/elasticity.py:35-53) */
#define SKB_FORM_LAPLACE 0          /* dot(grad(u), grad(v))              */
#define SKB_FORM_MASS 1             /* u * v   (vector: dot(u, v))        */
#define SKB_FORM_VECTOR_LAPLACE 2   /* ddot(grad(u), grad(v))             */
#define SKB_FORM_ELASTICITY 3       /* ddot(C(sym_grad(u)), sym_grad(v)); params = {Lambda, 2.*Mu} */
/* library linear forms (models/poisson.py:22-24) */
#define SKB_LFORM_UNIT_LOAD 0       /* v */

/* A finite element space on a mesh: Mesh.p/.t (mesh/mesh.py:27-28,558-559),
 * the reference tables of Element.lbasis at the quadrature points
 * (element/element_h1.py:10-24) and the rule itself (quadrature.py:12-77). */
typedef struct skb_space {
  int32_t dim;        /* 2 | 3                                              */
  int32_t nnodes;     /* rows of t: 3 (tri), 4 (tet), 8 (hex)               */
  int32_t mapping;    /* SKB_MAP_*                                          */
  int32_t nbs;        /* scalar basis functions per element                 */
  int32_t ncomp;      /* 1, or dim for ElementVector (element_vector.py:36-48): local index = ncomp*b + n */
  int32_t nqp;        /* quadrature points per element                      */
  int64_t npts;       /* p.shape[1] (leading dimension of p)                */
  int64_t nel_total;  /* t.shape[1]                                         */
  const double *p;    /* (dim, npts)  float64, coordinate-major             */
  const int32_t *t;   /* (nnodes, nel_total) int32, node-major              */
  const int32_t *tind;/* optional element subset (CellBasis(elements=...)), NULL = all */
  int64_t nel;        /* elements to process (= nel_total when tind==NULL)  */
  const double *phi;  /* (nbs, nqp)        lbasis values                    */
  const double *dphi; /* (nbs, dim, nqp)   lbasis reference gradients       */
  const double *W;    /* (nqp,)            quadrature weights               */
  const double *mdphi;/* (nnodes, dim, nqp) mapping-element gradients, ISO only */
  const double *mphi; /* (nnodes, nqp)      mapping-element values, ISO only (global coords) */
  const double *X;    /* (dim, nqp) quadrature points, needed by skb_tabulate's x output (affine) */
} skb_space_t;

/* ---- element-local data: replaces CellBasis.__init__ + Form._assemble ----
 * (assembly/basis/cell_basis.py:94-106; assembly/form/bilinear_form.py:58-128,
 *  150-151; linear_form.py:18-49).  Geometry (mapping_affine.py:55-131 /
 * mapping_isoparametric.py:112-226), push-forward (element_h1.py:17), the
 * integrand and the numpy-pairwise quadrature sum are fused; nothing of shape
 * (.., nel, nqp) is materialised.
 * out_local: (Nbu, Nbv, nel) float64, C order == COOData.data before
 * flatten (bilinear_form.py:121); Nbu = Nbv = nbs*ncomp.                   */
int skb_local_bilinear(const skb_space_t *space, int form, const double *params_host,
                       double *out_local, void *stream);
/* The same element-local data, bit for bit, written element-major: out_local_em is
 * (nel, Nbv, Nbu), i.e. out_local_em[e][i][j] == out_local[j][i][e] - the layout
 * skb_csr_reduce_em gathers from on warm re-assembly (the entries a CSR row takes from one
 * element are then consecutive in memory).  Affine meshes at the default rules of P1 / P2
 * (compile-time rule sizes); SKB_EINVAL otherwise - callers then use skb_local_bilinear.   */
int skb_local_bilinear_em(const skb_space_t *space, int form, const double *params_host,
                          double *out_local_em, void *stream);
/* out_local: (Nbv, nel) float64 (linear_form.py:41-44).                    */
int skb_local_linear(const skb_space_t *space, int form, const double *params_host,
                     double *out_local, void *stream);

/* ElementHex2 (27 functions) on trilinear hexahedra at the default 7^3 tensor rule, laplace |
 * mass (models/poisson.py:7-19 over element_hex2.py:1255-1260, mapping_isoparametric.py:
 * 112-226, quadrature.py:63-77): the same element-local data as skb_local_bilinear, computed
 * by sum factorisation (csrc/skb_hex_sf.cu) - one axis of the tensor-product rule at a time,
 * 63 k instead of 750 k multiply-adds per element.  Value-level parity (rtol 1e-12).  Host
 * tables: qstride[3] stride of each axis' 1-D index in the point index of `space`;
 * pp[4][9][nq] products of the 1-D quadratic functions (type bit 0 / 1: derivative on the
 * first / second factor; row 3 i + j; nodes 0, 1/2, 1); g[2][nq] = 1 - x, x; bnode[27] =
 * a + 3 b + 9 c (1-D node of each basis function per axis); vtx[8] local vertex at corner
 * 4 a + 2 b + c.  Returns SKB_EINVAL for anything but nq = 7 / 27 functions (callers then use
 * skb_local_bilinear), SKB_EZERODET like the reference raises (synchronises the stream).
 * element_major != 0: out_local is written as (nel, Nbv, Nbu) instead - 729 consecutive values
 * per element, the layout skb_csr_reduce_em gathers from (warm re-assembly; the plan builders
 * and COOData read the reference layout).                                                   */
int skb_local_hex_sumfact(const skb_space_t *space, int form, int32_t nq,
                          const int32_t *qstride_host, const double *pp_host,
                          const double *g_host, const uint8_t *bnode_host,
                          const uint8_t *vtx_host, int32_t element_major, double *out_local,
                          void *stream);

/* ---- sparsity plan: replaces COOData._assemble_scipy_csr's structure ----
 * (assembly/form/coo_data.py:27-36 -> scipy coo_matrix.eliminate_zeros +
 * tocsr: coo_tocsr, csr_sort_indices, csr_sum_duplicates).
 * COO entry k = (j*Nbv + i)*nel + e has row = dofs_v[i*nel+e],
 * col = dofs_u[j*nel+e] (bilinear_form.py:88-91).  Entries whose local value
 * is == 0.0 are dropped when drop_zeros != 0 (coo_data.py:35), which makes the
 * pattern value dependent.
 *
 * Step 1 (symbolic): stable device radix sort of (row*Ncols+col) keys, head
 * flags and their scan.  Scratch is caller-provided:
 *   keys_a/keys_b  uint64[ncoo], vals_a/vals_b uint32[ncoo], slot uint32[ncoo],
 *   tmp: skb_plan_scratch_bytes(ncoo) bytes.
 * Writes counts_host[0] = nnz (CSR slots), counts_host[1] = nkeep (surviving
 * triplets), counts_host[2] = 0|1 (sorted keys/vals ended up in the *_a | *_b
 * buffers) after synchronising the stream.  counts_host: int64[3].
 * Step 2 (finalize): fills caller-allocated indptr int32[nrows+1],
 * indices int32[nnz], segptr uint32[nnz+1], perm uint32[nkeep] where
 * perm[segptr[s] .. segptr[s+1]) are the COO entries of slot s in stable COO
 * order (entry-major, element-minor).                                       */
int64_t skb_plan_scratch_bytes(int64_t ncoo);
int skb_plan_symbolic(const int32_t *dofs_v, const int32_t *dofs_u, int32_t nbv, int32_t nbu,
                      int64_t nel, int64_t nrows, int64_t ncols,
                      const double *local_or_null, int drop_zeros,
                      uint64_t *keys_a, uint64_t *keys_b, uint32_t *vals_a, uint32_t *vals_b,
                      uint32_t *slot, void *tmp, int64_t tmp_bytes,
                      int64_t *counts_host, void *stream);
int skb_plan_finalize(int64_t ncoo, int64_t nrows, int64_t ncols, int64_t nnz, int64_t nkeep,
                      const uint64_t *keys_sorted, const uint32_t *vals_sorted, const uint32_t *slot,
                      int32_t *indptr, int32_t *indices, uint32_t *segptr, uint32_t *perm,
                      void *stream);
/* The same plan without a global sort (csrc/skb_plan_rows.cu): row r only receives entries
 * from the elements containing DOF r, so the triplets are bucketed by row (incidence lists
 * built with integer atomics) and every row is sorted by (col, k) on its own in shared memory
 * - one warp per row (in registers up to 128 surviving entries, in shared memory up to 512),
 * one CTA per row up to 8192.  Outputs are bit for
 * bit those of the two steps above.  Scratch (caller-provided, device): mask uint32[Nbv*nel],
 * rc uint64[nrows], sums uint32[4096], incstart / candstart uint32[nrows+1], cursor
 * uint32[2*nrows], nuniq uint32[nrows], inc_words uint32[2*counts[1]], dofs_ut int32[Nbu*nel],
 * ucol / uoff uint32[counts[0]],
 * flag int32[3].
 * skb_plan_rows_count: counts_host[0] = nkeep, counts_host[1] = incidences kept (stream
 * synchronised).  skb_plan_rows_sort: perm, indptr, *nnz_host (stream synchronised).
 * skb_plan_rows_emit: indices, segptr.  SKB_ETOOBIG (Nbu > 32, a row with more than 8192
 * entries, more than 2^24 rows ...): use the radix-sort path.                              */
int skb_plan_rows_count(const int32_t *dofs_v, int32_t nbv, int32_t nbu, int64_t nel,
                        int64_t nrows, const double *local_or_null, int drop_zeros,
                        uint32_t *mask, unsigned long long *rc, uint32_t *sums,
                        uint32_t *incstart, uint32_t *candstart, int64_t *counts_host,
                        void *stream);
int skb_plan_rows_sort(const int32_t *dofs_v, const int32_t *dofs_u, int32_t nbv, int32_t nbu,
                       int64_t nel, int64_t nrows, const uint32_t *mask,
                       const uint32_t *incstart, const uint32_t *candstart, uint32_t *cursor,
                       uint32_t *inc_words, int32_t *dofs_ut, uint32_t *sums, uint32_t *perm,
                       uint32_t *ucol,
                       uint32_t *uoff, uint32_t *nuniq, int32_t *indptr, int32_t *flag,
                       int64_t *nnz_host, void *stream);
int skb_plan_rows_emit(int64_t nrows, int64_t nnz, int64_t nkeep, const uint32_t *candstart,
                       const int32_t *indptr, const uint32_t *ucol, const uint32_t *uoff,
                       int32_t *indices, uint32_t *segptr, void *stream);
/* Mesh.build_entities (mesh/mesh.py:1065-1082: np.sort + np.unique(axis=1) over the vertex
 * tuples of all local edges / facets) through the same row-bucket machinery: the sorted unique
 * edges are the CSR pattern whose incidences are (local vertex i, element e) with the
 * surviving columns {j adjacent to i : vertex j > vertex i}; triangular facets use row = id of
 * the edge of the two smallest vertices, col = the largest vertex.  skb_entity_masks forms the
 * masks (bit j of mask[i*nel+e] <=> j in adj_host[i] and tu[j][e] > vmax[i][e]); pass them to
 * skb_plan_rows_count with local == NULL and drop_zeros == 2.  skb_plan_slot_of_entry inverts
 * perm / segptr (slot[k] = CSR slot of surviving COO entry k), which gives t2e / t2f.       */
/* MeshTet.init_tensor / MeshHex.init_tensor (mesh/mesh_tet_1.py:326-393,
 * mesh/mesh_hex_1.py:97-155) on the device (csrc/skb_mesh.cu): x, y, z sorted device
 * coordinates; corner_host[ntypes][nnodes] the cell corners (0..7) of every element type (six
 * Kuhn tetrahedra, or the hexahedron); p double[3][npx*npy*npz], t int32[nnodes][ntypes*ncells],
 * element order type-major.  Bit for bit the host generator's arrays.                      */
int skb_mesh_tensor(const double *x, const double *y, const double *z, int32_t npx, int32_t npy,
                    int32_t npz, int32_t ntypes, int32_t nnodes, const int32_t *corner_host,
                    double *p, int32_t *t, void *stream);
/* Dofs.__init__ (assembly/dofs.py:264-334), element_dofs on the device: row r of the result is
 * add_r + mul_r * src_r[e]; rows_dev = nrows records {int64 src (device pointer to an int32 row
 * of t / t2e / t2f, 0 = the element index), int32 mul, int32 add}; out int32[nrows][nel].   */
int skb_element_dofs(const void *rows_dev, int32_t nrows, int64_t nel, int32_t *out,
                     void *stream);
int skb_entity_masks(const int32_t *tu, int32_t nbu, int32_t nbv, int64_t nel,
                     const int32_t *vmax, const uint32_t *adj_host, uint32_t *mask, void *stream);
int skb_plan_slot_of_entry(const uint32_t *segptr, const uint32_t *perm, int64_t nnz,
                           int32_t *slot, void *stream);

/* ---- numeric phase: replaces csr_sum_duplicates (coo_data.py:36) ----
 * data[s] = sum of local[perm[k]], k in [segptr[s], segptr[s+1]), added
 * sequentially in that fixed order: deterministic, no float atomics.       */
int skb_csr_reduce(const double *local, const uint32_t *perm, const uint32_t *segptr,
                   int64_t nnz, double *data, void *stream);
/* The same sums, in the same order, over element-major local data local_em (nel, Nbv, Nbu):
 * COO entry k = (j*Nbv + i)*nel + e is read at e*Nbu*Nbv + i*Nbu + j, so the entries a CSR
 * row takes from one element are consecutive in memory (whole 32-byte sectors are used; in
 * the reference layout they are nel doubles apart).  perm / segptr are the plan's, unchanged. */
int skb_csr_reduce_em(const double *local_em, int64_t nel, int32_t nbu, int32_t nbv,
                      const uint32_t *perm, const uint32_t *segptr, int64_t nnz, double *data,
                      void *stream);
/* LinearForm scatter, replaces COOData.toarray 1-tensor branch / scipy
 * coo_todense (coo_data.py:102-108): vec[r] = sequential sum in COO order. */
int skb_vec_reduce(const double *local, const uint32_t *perm, const uint32_t *segptr,
                   const int32_t *indptr, int64_t nrows, double *vec, void *stream);

/* ---- fused P1 path (headline): ElementTetP1 Laplace straight to CSR values ----
 * One pass replaces CellBasis.__init__ (cell_basis.py:94-106),
 * BilinearForm._assemble (bilinear_form.py:58-128,150-151) and the value part of
 * COOData._assemble_scipy_csr (coo_data.py:27-36): local matrices are formed
 * in registers (bit-identical to numpy), staged in shared memory and reduced
 * per CSR slot inside the CTA; they never reach HBM.  The plan is built once
 * per (mesh, pattern) by skfem_b200/fused.py: elements are cut into tiles of
 * tile_elems and every tile owns one contiguous record in `rec` (header, tile-
 * local connectivity, vertex list, slot groups, slot targets, sliced-ELL
 * contribution indices; layout documented in fused.py), rec_start[t] being its
 * byte offset (multiple of 16).  The kernel is persistent and warp-specialised:
 * tile_elems compute threads (one element each) + reduce_threads reduce
 * threads per CTA; records stream through a ring of `ring` (4|5) shared
 * buffers of rec_cap bytes by TMA bulk copies, vertex coordinates are gathered
 * with cp.async; vcap = most vertices in a tile.  (tile_elems, reduce_threads)
 * in {(128,96) (256,128) (256,224) (256,256) (384,224) (384,352) (512,128) (512,256)
 * (512,384) (512,480) (768,224)}, one compute thread per element, or (512,736) (512,608)
 * [256 compute threads, two elements each] (512,640) [128 compute threads, four each];
 * one more warp per CTA issues the TMA copies.
 * skb_p1_fused_smem_bytes gives the dynamic shared memory a configuration
 * needs (must be <= 227 KB).  tame != 0 asserts that every vertex coordinate is
 * 0 or within [2^-60, 2^60] in magnitude, which lets the kernel use its
 * shared-reciprocal exact division without per-element range checks (tame == 0
 * selects plain IEEE division).  tame == 2 (opt-in, tile 512 / 480 reduce threads)
 * selects the fast arithmetic: fused multiply-adds and one reciprocal per element,
 * values within a few ulp per term of the reference order (inside the rtol 1e-12
 * bar for CSR values; element-local data no longer bit-identical).
 * w = the common quadrature weight (all weights of
 * the rule must be equal), nqp = number of quadrature points.  skb_p1_combine
 * adds, in tile order, the per-tile partial sums of CSR slots touched by more
 * than one tile.  No float atomics: bit-reproducible.                        */
/* profiling / test switches, default 0.  For skb_p1tet_laplace_fused: bit0 skips
 * the local-matrix phase, bit1 the slot-reduction phase (results meaningless),
 * bit2 prints per-role cycle counts of block 0.  bit3: skb_local_bilinear uses
 * the dense, uncached reference kernel instead of the cached/sparse one.    */
void skb_debug_flags(int flags);
int64_t skb_p1_fused_smem_bytes(int32_t tile_elems, int32_t ring, int32_t rec_cap, int32_t vcap);
int skb_p1tet_laplace_fused(const double *p, int64_t npts, const void *rec,
                            const uint64_t *rec_start, int32_t ntiles, int32_t tile_elems,
                            int32_t reduce_threads, int32_t ring, int32_t rec_cap, int32_t vcap,
                            int32_t tame, double w, int32_t nqp, double *csr_data,
                            double *scratch, void *stream);
/* The Laplace local matrix is bitwise symmetric, so only canonical slots
 * (row <= col) are reduced; gslot2[k] is the mirror CSR slot (col,row) that
 * receives the same sum (== gslot[k] on the diagonal).                       */
int skb_p1_combine(const double *scratch, const uint32_t *sptr, const uint32_t *gslot,
                   const uint32_t *gslot2, int64_t nshared, double *csr_data, void *stream);

/* ---- fused P1 path, second generation (csrc/skb_p1_fused2.cu, plan: skfem_b200/fused2.py) ----
 * Same replacement as above.  Elements are ordered by a k-d tree into super-tiles (compact
 * boxes of tiles_per_super consecutive tiles, the last one possibly shorter); a CTA (tile_elems
 * compute threads, one element each, + one producer warp; several CTAs per SM) walks whole
 * super-tiles: per tile P1 (local matrices
 * in registers -> shared memory) and P2 (fixed-order per-slot sums) accumulate into a pool of
 * pool_cap accumulators in shared memory, one per canonical CSR slot of the super-tile; after
 * the last tile the pool is flushed in CSR order through the flush table fl (uint32 pairs
 * {CSR slot, or bit31 | scratch position, or 0xFFFFFFFF for a padding entry; mirror slot or
 * 0xFFFFFFFF}, st_fl0[s] = first entry of super-tile s, always even; fetched by TMA).  Only slots shared between super-tiles go through scratch + skb_p1_combine2.
 * mode: 0 any coordinates / any equal-weight rule (IEEE division), 1 nqp == 4 and coordinates
 * 0 or within [2^-60, 2^60] (shared reciprocal + Markstein corrections == IEEE division),
 * 2 additionally coordinates within [2^-28, 2^28] and w within [2^-20, 1] (the quadrature sum
 * of 4 equal terms as one exact scaling; bit-identical), 3 opt-in FMA arithmetic.
 * p may differ from the coordinates the plan was built with (re-assembly on a moved mesh,
 * docs/examples/ex10.py-style loops): the kernel compares the zero mask of every local matrix
 * with the plan's and sets bit 0 of *flag when one differs - the value-dependent pattern
 * (coo_data.py:35) may then have changed and the caller must re-plan.
 * nz_out != NULL (plan time): only the local matrices are formed; the zero mask (bit k set <=>
 * entry k of the row-major upper triangle is nonzero) of element e of tile t is stored at
 * nz_out[t * tile_elems + e], from which the plan builder fills the expected masks.
 * ctas_per_sm: 0 = as many as fit.  No float atomics: bit-reproducible.                     */
int64_t skb_p1_fused2_smem_bytes(int32_t tile_elems, int32_t ring, int32_t rec_cap,
                                 int32_t vcap, int32_t pool_cap);
int skb_p1tet_laplace_fused2(const double *p, int64_t npts, const void *rec,
                             const uint64_t *rec_start, const int64_t *st_fl0, const void *fl,
                             int32_t nst, int32_t ntiles, int32_t tiles_per_super,
                             int32_t tile_elems, int32_t ring, int32_t rec_cap, int32_t vcap,
                             int32_t pool_cap, int32_t ctas_per_sm, int32_t mode, double w,
                             int32_t nqp, double *csr_data, double *scratch, int32_t *flag,
                             uint16_t *nz_out, void *stream);
/* The same pipeline for the mass form u * v (models/poisson.py:17-19) on ElementTetP1 with the
 * 4-point rule: tab_host = phi[4][4] (basis function x quadrature point) followed by W[4]
 * (host doubles).  Plan and records are built as for the Laplace form, from the mass matrix'
 * own (full graph) pattern; entries follow numpy's order  sum_q (phi_j phi_i) * (|det| W_q). */
int skb_p1tet_mass_fused2(const double *tab_host, const double *p, int64_t npts, const void *rec,
                          const uint64_t *rec_start, const int64_t *st_fl0, const void *fl,
                          int32_t nst, int32_t ntiles, int32_t tiles_per_super,
                          int32_t tile_elems, int32_t ring, int32_t rec_cap, int32_t vcap,
                          int32_t pool_cap, int32_t ctas_per_sm, double *csr_data,
                          double *scratch, int32_t *flag, uint16_t *nz_out, void *stream);
int skb_p1_combine2(const double *scratch, const uint32_t *sptr, const uint32_t *gslot,
                    const uint32_t *gslot2, int64_t nshared, double *csr_data, void *stream);

/* ---- multi-GPU interface rows: the device steps around the NCCL
 * all-to-all-v that replaces PETSc's MATIS -> mpiaij row addition (coo_data.py:151-170).
 * skb_pack_interface: out[i] = vals[slots[i]] (values of the CSR slots whose row a peer owns,
 * destination-major).  skb_unpack_add_interface: data[pos[i]] += recv[i] for one received
 * segment; targets within a segment are distinct and the caller issues the segments in
 * source-rank order on one stream, so the sums are deterministic (no atomics).            */
int skb_pack_interface(const double *vals, const int64_t *slots, int64_t n, double *out,
                       void *stream);
int skb_unpack_add_interface(double *data, const int64_t *pos, const double *recv, int64_t n,
                             void *stream);

/* ---- materialised basis for traced (user-defined) forms -----------------
 * grad: (dim, nel, nqp) of scalar basis function b (element_h1.py:17);
 * dx: (nel, nqp) (cell_basis.py:104-105); x: (dim, nel, nqp)
 * (mapping_affine.py:183-193 / mapping_isoparametric.py:52-58,170-171);
 * detabs: (nel, nqp) |detDF|.  Any output pointer may be NULL.             */
int skb_tabulate(const skb_space_t *space, int b, double *grad, double *dx, double *x,
                 double *detabs, void *stream);
/* Mapping.DF / invDF / detDF (mapping/mapping.py:6-114; mapping_affine.py:205-232,
 * mapping_isoparametric.py:173-226) at the space's quadrature points: DF, invDF
 * (dim, dim, nel, nqp), det (nel, nqp) signed.  Any output pointer may be NULL.  Hexahedra:
 * returns SKB_EZERODET like the reference raises (synchronises the stream).                 */
int skb_mapping(const skb_space_t *space, double *DF, double *invDF, double *det, void *stream);
/* out[e] = sum over q of integrand[e,q]*dx[e,q] (bilinear_form.py:150-151) in
 * numpy's order: pairwise (numpy pairwise_sum) when the product array is
 * C-ordered, plain left-to-right (sequential != 0) when it is Fortran-ordered
 * (the caller tracks the layout numpy would have produced).                 */
int skb_qp_reduce(const double *integrand, const double *dx, int64_t nel, int32_t nqp,
                  int sequential, double *out, void *stream);

/* ---- FacetBasis on affine meshes (assembly/basis/facet_basis.py:76-116) ---
 * Quadrature on `nf` mesh facets.  facets: (dim, nfacets_total) int32 vertex
 * indices; find[nf] facet indices; tind[nf] / tind_normals[nf] the elements the
 * trace / the normal are taken from (f2t[side, find] / f2t[0, find]);
 * lfacet[nf] the local index of each facet in tind_normals (row of t2f);
 * Xb (dim-1, nqp), Wb (nqp) the rule on the reference facet.  Outputs, any may
 * be NULL:  x (dim, nf, nqp) = G(Xb) (mapping_affine.py:234-246);  Y (dim, nf,
 * nqp) = invF(x, tind) (:195-203);  dx (nf, nqp) = |detB| Wb (:170-181, facet_
 * basis.py:114-115);  normals (dim, nf, nqp) (:248-281);  detabs (nf, nqp).   */
int skb_facet_geometry(const skb_space_t *space, const int32_t *facets, int64_t nfacets_total,
                       const int32_t *find, const int32_t *tind, const int32_t *tind_normals,
                       const int32_t *lfacet, int64_t nf, const double *Xb, const double *Wb,
                       int32_t nqp, double *x, double *Y, double *dx, double *normals,
                       double *detabs, void *stream);
/* Scalar basis function b at the per-facet local points Y: value (nf, nqp)
 * and, if grad != NULL, the pushed-forward gradient (dim, nf, nqp)
 * (ElementH1.gbasis, element/element_h1.py:10-18, with lbasis evaluated from
 * monomial tables: for function b and component c in {phi, d/dx0, ...}
 * poly_nterm[b*(1+dim)+c] terms, each poly_coef[(b*(1+dim)+c)*12 + k] times the
 * monomial whose exponents are the bytes of poly_expo[...]; terms are summed
 * left to right).                                                            */
int skb_facet_basis(const skb_space_t *space, const int32_t *tind, int64_t nf, int32_t nqp,
                    const double *Y, const double *poly_coef, const int32_t *poly_expo,
                    const int32_t *poly_nterm, int32_t b, double *value, double *grad,
                    void *stream);

/* ---- boundary conditions on the device CSR ------------
 * skfem.utils.enforce (utils.py:327-400): for every r in D[nD] the stored
 * entries of row r become 0.0 (they stay in the pattern) and the diagonal
 * entry becomes diag; *missing_diag (device int32, caller-zeroed) is set to 1
 * if a row of D has no stored diagonal (the reference would insert one).    */
int skb_csr_enforce(const int32_t *indptr, const int32_t *indices, double *data,
                    const int32_t *D, int64_t nD, double diag, int32_t *missing_diag,
                    void *stream);
/* skfem.utils.condense (utils.py:462-603) for ascending index sets: rows I[nI],
 * columns with colmap[col] >= 0 (colmap: int32[ncols], new column index or -1).
 * Pass 1 writes the per-row entry counts; the caller scans them into
 * new_indptr[nI+1]; pass 2 writes new_indices / new_data (= A[I][:, I]) and, if
 * bout != NULL, bout[k] = b[I[k]] - sum over dropped columns c of a*x[c], added
 * left to right from 0.0 (scipy csr_matvec order: bit-identical).            */
int skb_csr_condense_count(const int32_t *indptr, const int32_t *indices, const int32_t *I,
                           int64_t nI, const int32_t *colmap, int32_t *counts, void *stream);
int skb_csr_condense_fill(const int32_t *indptr, const int32_t *indices, const double *data,
                          const int32_t *I, int64_t nI, const int32_t *colmap,
                          const int32_t *new_indptr, int32_t *new_indices, double *new_data,
                          const double *x, const double *b, double *bout, void *stream);
/* y = A x with the row sums in scipy's csr_matvec order (hand-off to solvers) */
int skb_csr_spmv(const int32_t *indptr, const int32_t *indices, const double *data,
                 const double *x, double *y, int64_t nrows, void *stream);

/* Plan-time pass over the records' sliced-ELL index sections (the order in which
 * a CSR slot's triplets are added is unspecified in the reference, coo_data.py:
 * 34-36 -> scipy tocsr; it only has to be fixed): for each of the `ngroups`
 * 32-lane groups, whose len = grp_len[g] columns of 32 uint16 staging indices
 * start at rec16 + grp_pos[g], permute every lane's terms over the columns so that
 * each half-warp column touches every shared-memory bank pair at most
 * ceil(degree/len) times (bipartite edge colouring), and point unused cells at
 * the staged zero zero_base + b of the least loaded bank pair b.  In place.    */
int skb_p1_plan_spread(uint16_t *rec16, const int64_t *grp_pos, const int32_t *grp_len,
                       int64_t ngroups, int32_t zero_base, void *stream);

/* Plan-time pass over the records' tl / verts sections: renumbers every tile's local vertex
 * ids (bank pair + 16 * rank, greedy colouring of the (half-warp, slot) access sets) so that
 * the fused kernel's coordinate gathers are free of shared-memory bank conflicts.  The
 * vertex section of a record must hold 16 * (ceil(nv/16) + 1) entries (header word 0), all
 * valid vertex ids.  In place; any consistent numbering is valid.                      */
int skb_p1_plan_renumber(void *rec, const uint64_t *rec_start, int32_t ntiles,
                         int32_t tile_elems, void *stream);

/* L2 residency window on `stream` (cudaStreamAttributeAccessPolicyWindow, persisting hits)
 * for [ptr, ptr + bytes): used for the fused path's scratch array, written by the fused
 * kernel and read by skb_p1_combine right after.  bytes == 0 resets the stream's policy. */
int skb_l2_window(const void *ptr, int64_t bytes, void *stream);

/* Tuning knob (process-wide, default 0): the persistent fused kernel sizes its
 * grid for (SM count - sms) SMs, leaving room for kernels of other streams - the
 * NCCL all-to-all of the multi-GPU path, which otherwise cannot start before the
 * fused kernel of the next step has drained.                                   */
void skb_sm_reserve(int sms);

/* number of kernels of this library launched so far by this process (the
 * bench's `gpu_launches`); reset != 0 zeroes the counter after reading.     */
int64_t skb_launch_count(int reset);

/* library / build identification */
const char *skb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SKFEM_B200_H */
