"""Multi-GPU assembly: one process per GPU, element-partitioned, row-partitioned
result, interface rows exchanged over NCCL (NVLink 5 / NVSwitch).

The reference's only distributed route is host MPI + PETSc: METIS partition ->
per-rank sub-mesh + local->global DOF map (``Dofs.decompose``,
skfem/assembly/dofs.py:358-511) -> every rank runs the serial assembler ->
PETSc ``MATIS -> mpiaij`` adds the interface rows across ranks
(skfem/assembly/form/coo_data.py:118-172).  The same decomposition here:

* every rank assembles its own elements with the single-GPU engine into a
  *local* CSR (local DOF numbering, value-dependent pattern);
* ``l2g`` maps local DOFs to global ones (the reference's ``_globnums``);
  global rows are owned in contiguous ranges ``[ranges[k], ranges[k+1])``;
* plan (once): the (row, col) keys of slots whose row another rank owns are
  sent to that owner; the owner merges them with its own keys (sorted unique)
  into its final row block and remembers where every incoming value lands.
  A slot exists iff some rank has a non-zero contribution - the union of the
  ranks' value-dependent patterns, i.e. the reference's pattern;
* numeric (every assembly): pack the off-rank values, one all-to-all-v,
  then add the received segments in source-rank order.  Inside one segment all
  target slots are distinct, so the result is deterministic.

Only neighbours exchange data for slab-like partitions; the payload is one
vertex layer (a few MB), so the exchange is latency- not bandwidth-bound and is
left to NCCL point-to-point (no data-path collective beyond it).

The exchange logic is device agnostic (torch CPU tensors + gloo work too):
tests/test_distributed_cpu.py drives it with world_size 2 on CPU.
"""
from __future__ import annotations

import numpy as np


def _lib():
    from . import _lib as L
    return L.lib()


def _torch():
    import torch
    return torch


def balanced_ranges(n, world):
    """Contiguous ownership ranges of n rows over `world` ranks."""
    base, extra = divmod(int(n), int(world))
    ends = np.cumsum([base + (1 if k < extra else 0) for k in range(world)])
    return np.concatenate([[0], ends]).astype(np.int64)


def all_to_all_v(send, send_counts, recv_counts, group=None):
    """Variable all-to-all on the current backend via grouped point-to-point
    (NCCL: one ncclGroup of send/recv; gloo: the same calls on CPU tensors).
    `send` is laid out destination-major."""
    torch = _torch()
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    recv = torch.empty(int(sum(recv_counts)), dtype=send.dtype, device=send.device)
    if dist.get_backend(group) == "nccl":        # one grouped ncclSend/ncclRecv
        dist.all_to_all_single(recv, send.contiguous(), [int(c) for c in recv_counts],
                               [int(c) for c in send_counts], group=group)
        return recv
    so = np.concatenate([[0], np.cumsum(send_counts)]).astype(np.int64)
    ro = np.concatenate([[0], np.cumsum(recv_counts)]).astype(np.int64)
    ops = []
    for peer in range(world):
        if peer == rank:
            continue
        if recv_counts[peer]:
            ops.append(dist.P2POp(dist.irecv, recv[ro[peer]:ro[peer + 1]], peer, group))
        if send_counts[peer]:
            ops.append(dist.P2POp(dist.isend, send[so[peer]:so[peer + 1]].contiguous(), peer,
                                  group))
    if send_counts[rank]:
        recv[ro[rank]:ro[rank + 1]] = send[so[rank]:so[rank + 1]]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return recv


def _pack(vals, slots):
    """vals[slots] - on CUDA through the library's pack kernel."""
    if not vals.is_cuda:
        return vals[slots]
    import ctypes as C
    torch = _torch()
    from . import _lib as L
    out = torch.empty(slots.shape[0], dtype=vals.dtype, device=vals.device)
    L.check(_lib().skb_pack_interface(vals.data_ptr(), slots.data_ptr(), slots.shape[0],
                                      out.data_ptr(),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)),
            "skb_pack_interface")
    return out


def _add_segment(data, pos, recv):
    """data[pos] += recv for one source rank's segment (distinct targets)."""
    if not data.is_cuda:
        data.index_add_(0, pos, recv)
        return
    import ctypes as C
    torch = _torch()
    from . import _lib as L
    L.check(_lib().skb_unpack_add_interface(data.data_ptr(), pos.data_ptr(), recv.data_ptr(),
                                            pos.shape[0],
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)),
            "skb_unpack_add_interface")


class InterfaceExchange:
    """Plan + numeric phase of the interface-row reduction for one local CSR
    pattern.  ``grow``/``gcol``: global row / column of every local CSR slot
    (int64 tensors on the compute device)."""

    def __init__(self, grow, gcol, ranges, ncols, group=None):
        torch = _torch()
        import torch.distributed as dist
        self.group = group
        self.rank = rank = dist.get_rank(group)
        self.world = world = dist.get_world_size(group)
        self.ranges = np.asarray(ranges, dtype=np.int64)
        self.ncols = int(ncols)
        dev = grow.device
        i64 = torch.int64
        r0, r1 = int(self.ranges[rank]), int(self.ranges[rank + 1])
        nloc = int(grow.shape[0])
        key = grow * self.ncols + gcol
        mine = (grow >= r0) & (grow < r1)
        own_slots = torch.nonzero(mine).flatten()
        own_keys = key[own_slots]
        if own_keys.numel() > 1 and not bool((own_keys[1:] > own_keys[:-1]).all()):
            # l2g not monotone: sort once (the local CSR order is then not the global one)
            own_keys, o = torch.sort(own_keys)
            own_slots = own_slots[o]
        # only slots whose row another rank owns travel: destination-major, key-sorted
        rem_slots = torch.nonzero(~mine).flatten()
        ends = torch.as_tensor(self.ranges[1:], device=dev)
        rem_owner = torch.searchsorted(ends, grow[rem_slots], right=True)
        rem_keys = key[rem_slots]
        o = torch.argsort(rem_owner * (int(self.ranges[-1]) * self.ncols + 1) + rem_keys)
        rem_slots, rem_keys, rem_owner = rem_slots[o], rem_keys[o], rem_owner[o]
        cnt_t = torch.bincount(rem_owner, minlength=world)[:world].to(i64)
        recv_cnt = all_to_all_v(cnt_t, [1] * world, [1] * world, group)
        both = torch.stack([cnt_t, recv_cnt]).cpu()            # one synchronisation
        self.send_counts_remote = [int(c) for c in both[0]]
        self.recv_counts_remote = [int(c) for c in both[1]]
        recv_keys = all_to_all_v(rem_keys, self.send_counts_remote, self.recv_counts_remote,
                                 group)
        # my final row block = own keys (sorted, unique) merged with the received keys that
        # are new to me; no sort of the big array: new keys are few (one interface layer)
        pos = torch.searchsorted(own_keys, recv_keys).clamp(max=max(int(own_keys.numel()) - 1, 0))
        have = (own_keys[pos] == recv_keys) if own_keys.numel() else torch.zeros_like(
            recv_keys, dtype=torch.bool)
        new_keys = torch.unique(recv_keys[~have], sorted=True)
        shift = torch.searchsorted(new_keys, own_keys)          # new keys below each own key
        own_final = torch.arange(int(own_keys.numel()), device=dev, dtype=i64) + shift
        new_final = torch.searchsorted(own_keys, new_keys) + torch.arange(
            int(new_keys.numel()), device=dev, dtype=i64)
        self.nnz = int(own_keys.numel() + new_keys.numel())
        final = torch.empty(self.nnz, dtype=i64, device=dev)
        final[own_final] = own_keys
        final[new_final] = new_keys
        self.pos_recv_remote = torch.where(
            have, own_final[pos] if own_keys.numel() else pos,
            new_final[torch.searchsorted(new_keys, recv_keys).clamp(
                max=max(int(new_keys.numel()) - 1, 0))] if new_keys.numel() else pos)
        rows = final // self.ncols - r0
        self.nrows = r1 - r0
        self.row0 = r0
        self.indices = (final - (rows + r0) * self.ncols)
        rowcount = torch.bincount(rows, minlength=self.nrows)[:self.nrows]
        self.indptr = torch.cat([torch.zeros(1, dtype=i64, device=dev),
                                 torch.cumsum(rowcount, 0)])
        # the block is handed out like scipy's own CSR: int32 pattern whenever it fits
        if self.nnz < 2 ** 31 and self.ncols < 2 ** 31:
            self.indices = self.indices.to(torch.int32)
            self.indptr = self.indptr.to(torch.int32)
        self.ro_remote = np.concatenate([[0], np.cumsum(self.recv_counts_remote)]).astype(np.int64)
        self.bytes_per_exchange = 8 * sum(self.send_counts_remote)
        # ---- direct-write layout for the fused kernel --------------------------------
        # The kernel can write every local CSR slot straight to its destination:
        # out = [ my final row block (nnz) | send buffer, destination-major ].
        # slot_map[local slot] = index in `out`.
        self.own_slots, self.own_final, self.rem_slots = own_slots, own_final, rem_slots
        slot_map = torch.empty(nloc, dtype=i64, device=dev)
        slot_map[own_slots] = own_final
        self.nsend = int(rem_slots.shape[0])
        slot_map[rem_slots] = self.nnz + torch.arange(self.nsend, device=dev)
        self.slot_map = slot_map
        # final slots nobody on this rank writes (received-only) must start at zero
        self.unwritten = new_final

    def finish(self, out, zero=True):
        """Numeric phase when the kernel wrote `out` through ``slot_map``:
        exchange the send-buffer tail and add the received segments (source-rank
        order) into the row block at the head of `out`.  Returns the block.
        ``zero=False``: the caller already cleared the receive-only slots."""
        data = out[:self.nnz]
        if zero and self.unwritten.numel():
            data.index_fill_(0, self.unwritten, 0.0)
        if self.world > 1:
            recv = all_to_all_v(out[self.nnz:], self.send_counts_remote, self.recv_counts_remote,
                                self.group)
            for src in range(self.world):
                a, b = int(self.ro_remote[src]), int(self.ro_remote[src + 1])
                if b > a:
                    _add_segment(data, self.pos_recv_remote[a:b], recv[a:b])
        return data

    def reduce(self, local_vals):
        """Values of my row block from every rank's local values: own slots first, then the
        received segments in source-rank order (the same order as :meth:`finish`)."""
        torch = _torch()
        data = torch.zeros(self.nnz, dtype=local_vals.dtype, device=local_vals.device)
        data[self.own_final] = local_vals[self.own_slots]
        if self.world > 1:
            recv = all_to_all_v(_pack(local_vals, self.rem_slots), self.send_counts_remote,
                                self.recv_counts_remote, self.group)
            for src in range(self.world):                        # fixed order: deterministic
                a, b = int(self.ro_remote[src]), int(self.ro_remote[src + 1])
                if b > a:                                        # distinct targets per source
                    _add_segment(data, self.pos_recv_remote[a:b], recv[a:b])
        return data


class DistributedCSR:
    """Row block of the global matrix owned by this rank: rows
    ``[row0, row0+nrows)``, global column indices."""

    def __init__(self, indptr, indices, data, row0, shape, ready=None):
        self.indptr, self.indices, self.data = indptr, indices, data
        self.row0, self.shape = int(row0), tuple(shape)
        self.ready = ready     # CUDA event after which `data` is complete (pipelined mode)

    def wait(self):
        """Make the current stream wait until the values are complete."""
        if self.ready is not None:
            _torch().cuda.current_stream().wait_event(self.ready)
        return self

    @property
    def nnz(self):
        return int(self.data.shape[0])

    def to_scipy_block(self):
        from scipy.sparse import csr_matrix
        nrows = int(self.indptr.shape[0]) - 1
        self.wait()
        if self.data.is_cuda:          # pinned staging, three overlapping async copies
            torch = _torch()

            def d2h(t):
                h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                h.copy_(t, non_blocking=True)
                return h
            hd, hi, hp = d2h(self.data), d2h(self.indices), d2h(self.indptr)
            torch.cuda.current_stream().synchronize()
            A = csr_matrix((hd.numpy(), hi.numpy(), hp.numpy()), shape=(nrows, self.shape[1]),
                           copy=False)
            A.has_sorted_indices = True      # sorted unique keys by construction
            A.has_canonical_format = True
            return A
        return csr_matrix((self.data.cpu().numpy(), self.indices.cpu().numpy(),
                           self.indptr.cpu().numpy()), shape=(nrows, self.shape[1]))


class DistributedAssembler:
    """``form`` assembled over a rank-local basis into a row-partitioned CSR.

    basis   rank-local CellBasis (sub-mesh with local DOF numbering)
    l2g     int64 array, local DOF -> global DOF (the reference's _globnums)
    N       global number of DOFs
    ranges  ownership ranges of global rows (len world+1); default balanced
    """

    def __init__(self, form, basis, l2g, N, ranges=None, group=None, reuse_buffers=False,
                 graph_exchange=False, pipeline=False, sm_reserve=0, depth=2):
        # graph_exchange=True also captures the NCCL all-to-all in the CUDA graph; it hung
        # on the B200 box with torch 2.11 / NCCL 2.28 (round 1), so it is opt-in.
        import torch.distributed as dist
        self.form, self.basis, self.N, self.group = form, basis, int(N), group
        self.reuse_buffers = bool(reuse_buffers)
        self.graph_exchange = bool(graph_exchange)
        # pipeline=True (with reuse_buffers): `depth` output buffer sets; the interface
        # exchange + ordered add of step i run on a side stream while the local
        # kernels of step i+1 run on the caller's stream.  The returned block
        # carries the event that completes it (DistributedCSR.wait) and is
        # overwritten `depth` calls later (a third set was measured at 2 GPUs: same
        # 0.194 ms per step, so the wait for a set's own exchange is not what the step
        # loses against one GPU - profiles/r2_trace_2gpu.md).
        self.pipeline = bool(pipeline) and self.reuse_buffers
        self.depth = max(2, int(depth))
        # sm_reserve (pipelined mode): SMs the persistent fused kernel leaves free so
        # that the NCCL kernel of the previous step's exchange can run beside it
        self.sm_reserve = int(sm_reserve)
        self._sets = None
        self._graph = self._out = self._data = None
        self._graph_has_exchange = False
        self.world = dist.get_world_size(group)
        self.ranges = balanced_ranges(N, self.world) if ranges is None else np.asarray(ranges)
        self.l2g_host = np.asarray(l2g, dtype=np.int64)
        self.exchange = None

    def assemble(self):
        torch = _torch()
        if self.exchange is None:                             # cold: local plan + exchange plan
            A = self.form.assemble_device(self.basis)         # local CSR, local numbering
            dev = A.data.device
            l2g = torch.as_tensor(self.l2g_host, device=dev)
            counts = (A.indptr[1:] - A.indptr[:-1]).long()
            lrow = torch.repeat_interleave(torch.arange(A.shape[0], device=dev), counts)
            self.exchange = ex = InterfaceExchange(l2g[lrow], l2g[A.indices.long()],
                                                   self.ranges, self.N, self.group)
            data = ex.reduce(A.data)
            return DistributedCSR(ex.indptr, ex.indices, data, ex.row0, (self.N, self.N))
        # warm: the kernels write every value straight to its place in
        # [row block | send buffer]; then one exchange + ordered add
        ex = self.exchange
        if not self.reuse_buffers:
            out = torch.empty(ex.nnz + ex.nsend, dtype=torch.float64, device=ex.slot_map.device)
            self.form.assemble_device(self.basis, out=out, slot_map=ex.slot_map)
            data = ex.finish(out)
            return DistributedCSR(ex.indptr, ex.indices, data, ex.row0, (self.N, self.N))
        if self.pipeline:
            return self._assemble_pipelined()
        # re-assembly loop mode: persistent output buffer, the local kernels are
        # replayed from a CUDA graph, only the NCCL exchange is issued eagerly.
        # The returned block aliases the internal buffer (overwritten next call).
        if self._graph is None:
            self._out = torch.empty(ex.nnz + ex.nsend, dtype=torch.float64,
                                    device=ex.slot_map.device)
            self.form.assemble_device(self.basis, out=self._out, slot_map=ex.slot_map)  # plan
            ex.finish(self._out)                                  # warms NCCL channels
            torch.cuda.synchronize()
            self._graph_has_exchange = False
            if self.graph_exchange:
                # whole step (kernels + NCCL all-to-all + ordered adds) in one graph
                try:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self.form.assemble_device(self.basis, out=self._out,
                                                  slot_map=ex.slot_map)
                        self._data = ex.finish(self._out)
                    self._graph, self._graph_has_exchange = g, True
                except Exception:                                  # NCCL capture unsupported
                    torch.cuda.synchronize()
                    self._graph = None
            if self._graph is None:
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    if ex.unwritten.numel():
                        self._out[:ex.nnz].index_fill_(0, ex.unwritten, 0.0)
                    self.form.assemble_device(self.basis, out=self._out, slot_map=ex.slot_map)
        self._graph.replay()
        data = self._data if self._graph_has_exchange else ex.finish(self._out, zero=False)
        return DistributedCSR(ex.indptr, ex.indices, data, ex.row0, (self.N, self.N))

    def _assemble_pipelined(self):
        torch = _torch()
        ex = self.exchange
        cur = torch.cuda.current_stream()
        if self._sets is None:
            dev = ex.slot_map.device
            self._comm = torch.cuda.Stream(device=dev)
            self._sets, self._flip = [], 0
            for _ in range(self.depth):
                out = torch.empty(ex.nnz + ex.nsend, dtype=torch.float64, device=dev)
                self.form.assemble_device(self.basis, out=out, slot_map=ex.slot_map)  # plan, warm
                ex.finish(out)                                                         # NCCL warm
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                _lib().skb_sm_reserve(self.sm_reserve)      # grid size is baked into the graph
                try:
                    with torch.cuda.graph(g):
                        if ex.unwritten.numel():
                            out[:ex.nnz].index_fill_(0, ex.unwritten, 0.0)
                        self.form.assemble_device(self.basis, out=out, slot_map=ex.slot_map)
                finally:
                    _lib().skb_sm_reserve(0)
                self._sets.append({"out": out, "graph": g,
                                   "computed": torch.cuda.Event(), "done": torch.cuda.Event()})
        st = self._sets[self._flip]
        self._flip = (self._flip + 1) % self.depth
        cur.wait_event(st["done"])            # the exchange that last read this set is over
        st["graph"].replay()
        st["computed"].record(cur)
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(st["computed"])
            data = ex.finish(st["out"], zero=False)
            st["done"].record(self._comm)
        return DistributedCSR(ex.indptr, ex.indices, data, ex.row0, (self.N, self.N),
                              ready=st["done"])

    def wait(self):
        """Current stream waits for every outstanding exchange (pipelined mode)."""
        if self._sets:
            cur = _torch().cuda.current_stream()
            for st in self._sets:
                cur.wait_event(st["done"])


def slab_mesh_tet(cells_xy, cells_z, rank, world, mesh_cls=None):
    """Rank-local z-slab of a (cells_xy x cells_xy x cells_z*world) init_tensor
    tetrahedral mesh of the unit-spaced box, its local->global vertex map and
    the global vertex count.  Vertex numbering of ``MeshTet.init_tensor`` is
    z-major (index = iy + npy*ix + npy*npx*iz), so a z-slab owns a contiguous
    range of global vertices and shares exactly one vertex layer with each
    neighbour."""
    from .mesh import MeshTet
    mesh_cls = MeshTet if mesh_cls is None else mesh_cls
    x = np.linspace(0, 1, cells_xy + 1)
    z = np.linspace(rank * 1.0, rank + 1.0, cells_z + 1)
    m = mesh_cls.init_tensor(x, x, z)
    layer = (cells_xy + 1) ** 2
    l2g = np.arange(m.p.shape[1], dtype=np.int64) + rank * layer * cells_z
    N = layer * (cells_z * world + 1)
    # rows of a slab's top layer belong to the next rank (it owns z >= its base)
    ranges = np.array([k * layer * cells_z for k in range(world)] + [N], dtype=np.int64)
    return m, l2g, N, ranges


def partition(mesh, world, rank):
    """Element partition of one global mesh by owning row range: contiguous,
    element-balanced row ranges ``[ranges[k], ranges[k+1])``; element e belongs to the rank
    that owns the row of its smallest vertex DOF, so a rank's contributions go to its own
    rows or to higher ranks only.  Returns ``(submesh, l2g, N, ranges)`` like
    :func:`slab_mesh_tet`: the rank-local mesh (vertices renumbered in ascending global
    order, so local CSR rows stay sorted by global row), the local -> global vertex map (the
    reference's ``_globnums``, skfem/assembly/dofs.py:478-481; its METIS dual-graph partition,
    dofs.py:462-467, is replaced by row ranges because ``init_tensor`` numbers vertices
    z-major, mesh_tet_1.py:343-352 - and because raw element blocks interleave the six
    Kuhn tets of every cube, mesh_tet_1.py:385-391).  One DOF per vertex (P1-type spaces)."""
    t = np.asarray(mesh.t)
    N = int(mesh.p.shape[1])
    nel = int(t.shape[1])
    key = t.min(axis=0)
    ks = np.sort(key, kind="stable")
    cuts = [0]
    for k in range(1, world):
        cuts.append(max(int(ks[min((k * nel) // world, nel - 1)]), cuts[-1]))
    ranges = np.asarray(cuts + [N], dtype=np.int64)
    owner = np.searchsorted(ranges[1:], key, side="right")
    mine = np.nonzero(owner == rank)[0]
    tl = t[:, mine]
    l2g = np.unique(tl).astype(np.int64)
    t_local = np.searchsorted(l2g, tl).astype(np.int32)
    sub = type(mesh)(np.ascontiguousarray(mesh.p[:, l2g]), np.ascontiguousarray(t_local))
    return sub, l2g, N, ranges
