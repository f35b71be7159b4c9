"""Multi-GPU side of the explicit step: one process per GPU (torchrun), one mesh
partition per rank, shared nodes duplicated on every rank that touches them.

What the reference does (and this module mirrors):
  * partition maps -- local node numbering and the per-neighbour shared-node lists,
    src/io/PartitionMesh.cpp:485-536 (renumbering: local id = rank among the sorted unique
    global ids) and :566-1128 (sendProcessID ascending, per neighbour the shared nodes in
    ascending GLOBAL id, identical on both sides);
  * per step, the pairwise sum of partial internal forces at shared nodes,
    src/fem/SolidMechanics/GetForce_3D.cpp:54-102, and the MIN of the stable time step,
    src/timestep/StableTimeStep.cpp:33; once, the same sum for the lumped mass,
    src/fem/Mass/Mass3D.cpp:77-125.

Transport: `torch.distributed` point-to-point (NCCL over NVLink on GPUs, gloo on CPU for
the host-logic tests).  The payload layout is the reference's sendNodeDisplacement /
recvNodeDisplacement (neighbour-major, xyz interleaved), so slice i of the send window
goes to neighbour i's slice of the receive window and no index exchange is needed.
The element partition itself (ParMETIS in the reference) is an INPUT here: any
`part[]` array works; `brick_partition` provides the structured split used for the
synthetic scaling cubes.
"""
import numpy as np


# ------------------------------------------------------------------------------- partition maps
def maps_from_elements(global_conn, elem_ids_by_rank):
    """Restates PartitionMesh.cpp:485-536 and :1071-1108 for a given element distribution.

    global_conn: [E,8] global node ids (0-based); elem_ids_by_rank[r]: global element ids owned by rank r
    in the rank's local element order.  Returns per rank a dict with
      connectivity (local ids, [E_r*8]), globalNodeID (1-based, sorted), node_gids (0-based),
      sendProcessID, sendNeighbourCount, sendNeighbourCountCum, sendNodeIndex (local ids).
    O(N log N); the reference's O(N^2) scan (:492-513) gives the same result.
    """
    P = len(elem_ids_by_rank)
    gconn = np.asarray(global_conn).reshape(-1, 8)
    node_sets, out = [], []
    for r in range(P):
        c = gconn[np.asarray(elem_ids_by_rank[r], dtype=np.int64)]
        gids = np.unique(c)  # sorted unique global ids -> local id = rank in this list
        node_sets.append(gids)
        local = np.searchsorted(gids, c).astype(np.int32)
        out.append({"connectivity": local.reshape(-1), "node_gids": gids.astype(np.int64),
                    "globalNodeID": (gids + 1).astype(np.int32)})
    for r in range(P):
        pids, counts, idx = [], [], []
        for q in range(P):  # ascending neighbour rank
            if q == r:
                continue
            shared = np.intersect1d(node_sets[r], node_sets[q], assume_unique=True)  # ascending global id
            if shared.size:
                pids.append(q)
                counts.append(shared.size)
                idx.append(np.searchsorted(node_sets[r], shared).astype(np.int32))
        out[r]["sendProcessID"] = np.array(pids, dtype=np.int32)
        out[r]["sendNeighbourCount"] = np.array(counts, dtype=np.int32)
        out[r]["sendNeighbourCountCum"] = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        out[r]["sendNodeIndex"] = np.concatenate(idx).astype(np.int32) if idx else np.zeros(0, np.int32)
    return out


def proc_grid(P):
    """Near-cubic px*py*pz = P, px >= py >= pz (1,2,4,8 -> 1x1x1, 2x1x1, 2x2x1, 2x2x2)."""
    best = (P, 1, 1)
    for a in range(1, P + 1):
        if P % a:
            continue
        for b in range(1, P // a + 1):
            if (P // a) % b:
                continue
            c = P // a // b
            t = tuple(sorted((a, b, c), reverse=True))
            if max(t) - min(t) < max(best) - min(best):
                best = t
    return best


def brick_partition(n_local, pgrid, rank, L_local=0.005):
    """Rank `rank`'s brick of a structured hex8 box made of pgrid = (px,py,pz) bricks of n_local^3 elements each
    (n_local may be a triple (nx, ny, nz) for bricks that are not cubes: strong scaling of a fixed global cube).  Weak
    scaling: the global mesh grows with the rank count.  Node and element numbering, C3D8 ordering and spacing follow
    femtech_b200.mesh.cube_mesh / box_mesh; the maps follow the reference's rules (see maps_from_elements) but are built
    analytically, without the global mesh.  The spacing is L_local / nx.

    Returns dict(coordinates[N,3], connectivity[E,8] local ids, pid[E], node_gids[N], comm{...},
                 box=(Lx,Ly,Lz), dims=(Nx,Ny,Nz)).
    """
    px, py, pz = pgrid
    nx, ny, nz = (n_local, n_local, n_local) if np.isscalar(n_local) else tuple(int(v) for v in n_local)
    rx, ry, rz = rank % px, (rank // px) % py, rank // (px * py)
    Nx, Ny, Nz = nx * px, ny * py, nz * pz
    h = L_local / nx
    ox, oy, oz = rx * nx, ry * ny, rz * nz
    kk, jj, ii = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    gi, gj, gk = (ii + ox).reshape(-1), (jj + oy).reshape(-1), (kk + oz).reshape(-1)
    X = np.stack([gi * h, gj * h, gk * h], axis=-1).astype(np.float64)
    gids = gi.astype(np.int64) + (Nx + 1) * gj.astype(np.int64) + (Nx + 1) * (Ny + 1) * gk.astype(np.int64)
    ek, ej, ei = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    ei, ej, ek = ei.reshape(-1), ej.reshape(-1), ek.reshape(-1)
    nx1, nxy = nx + 1, (nx + 1) * (ny + 1)

    def nid(i, j, k):
        return (i + nx1 * j + nxy * k).astype(np.int32)

    conn = np.stack([nid(ei, ej, ek), nid(ei + 1, ej, ek), nid(ei + 1, ej + 1, ek), nid(ei, ej + 1, ek),
                     nid(ei, ej, ek + 1), nid(ei + 1, ej, ek + 1), nid(ei + 1, ej + 1, ek + 1),
                     nid(ei, ej + 1, ek + 1)], axis=1).astype(np.int32)
    # neighbours: every brick that shares at least one node (faces, edges, corners), ascending rank;
    # shared nodes of a pair = a face/edge/corner of the local node box, ascending global id == ascending local id
    pids, counts, idx = [], [], []
    local_ids = np.arange((nx + 1) * (ny + 1) * (nz + 1), dtype=np.int32).reshape(nz + 1, ny + 1, nx + 1)  # [k, j, i]
    for q in range(px * py * pz):
        if q == rank:
            continue
        qx, qy, qz = q % px, (q // px) % py, q // (px * py)
        dx, dy, dz = qx - rx, qy - ry, qz - rz
        if max(abs(dx), abs(dy), abs(dz)) > 1:
            continue
        sel = []
        for d, nn in ((dz, nz), (dy, ny), (dx, nx)):
            sel.append(slice(None) if d == 0 else (slice(nn, nn + 1) if d > 0 else slice(0, 1)))
        shared = local_ids[sel[0], sel[1], sel[2]].reshape(-1)
        pids.append(q)
        counts.append(shared.size)
        idx.append(np.sort(shared))
    comm = {
        "sendProcessID": np.array(pids, dtype=np.int32),
        "sendNeighbourCount": np.array(counts, dtype=np.int32),
        "sendNeighbourCountCum": np.concatenate([[0], np.cumsum(counts)]).astype(np.int32),
        "sendNodeIndex": np.concatenate(idx).astype(np.int32) if idx else np.zeros(0, np.int32),
    }
    return {"coordinates": X, "connectivity": conn, "pid": np.zeros(conn.shape[0], np.int32), "node_gids": gids,
            "comm": comm, "box": (Nx * h, Ny * h, Nz * h), "dims": (Nx, Ny, Nz)}


# ------------------------------------------------------------------------------------ transport
class HaloExchange:
    """Pairwise exchange of the shared-node windows (GetForce_3D.cpp:62-90) over torch.distributed."""

    def __init__(self, comm, device, dist):
        import torch
        self.torch, self.dist = torch, dist
        self.pids = [int(p) for p in comm["sendProcessID"]]
        self.cum = [int(c) for c in comm["sendNeighbourCountCum"]]
        total = self.cum[-1] if self.cum else 0
        self.count = total
        self.send = torch.zeros(3 * max(total, 1), dtype=torch.float64, device=device)
        self.recv = torch.zeros(3 * max(total, 1), dtype=torch.float64, device=device)

    def exchange(self):
        """send window slice i -> neighbour i; neighbour i's slice -> recv window slice i."""
        if not self.pids:
            return
        ops = []
        for i, p in enumerate(self.pids):
            lo, hi = 3 * self.cum[i], 3 * self.cum[i + 1]
            ops.append(self.dist.P2POp(self.dist.isend, self.send[lo:hi], p))
            ops.append(self.dist.P2POp(self.dist.irecv, self.recv[lo:hi], p))
        for w in self.dist.batch_isend_irecv(ops):
            w.wait()


def halo_add_host(field_aos, comm, recv):
    """CPU restatement of the add loop (GetForce_3D.cpp:92-97): slots in neighbour-major order."""
    idx = np.asarray(comm["sendNodeIndex"])
    f = field_aos.reshape(-1, 3)
    r = np.asarray(recv).reshape(-1, 3)
    for i in range(idx.size):  # sequential on purpose: a node shared with several neighbours is hit repeatedly
        f[idx[i]] += r[i]


def p2p_metadata(comms, rank):
    """For rank `rank`: per neighbour i -> (first slot of our slice in its receive window, our index in its
    neighbour list, its total slot count).  comms: every rank's comm pattern (the send lists are symmetric)."""
    c = comms[rank]
    off, idx, tot = [], [], []
    for q in c["sendProcessID"]:
        cq = comms[int(q)]
        j = int(np.where(np.asarray(cq["sendProcessID"]) == rank)[0][0])
        off.append(int(cq["sendNeighbourCountCum"][j]))
        idx.append(j)
        tot.append(int(cq["sendNeighbourCountCum"][-1]))
    return (np.array(off, dtype=np.int32), np.array(idx, dtype=np.int32), np.array(tot, dtype=np.int32))


class _DevDouble:
    """Zero-copy view of one device double for torch (CUDA array interface)."""

    def __init__(self, ptr):
        self.__cuda_array_interface__ = {"shape": (1,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class _DevInt32:
    """__cuda_array_interface__ view of n int32 in device memory owned by the C-ABI context."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<i4", "data": (int(ptr), False), "version": 2}


class DistFemTech:
    """One rank of a multi-GPU run: a solver.FemTech plus the exchange, driving the split step API."""

    def __init__(self, part, materialID, properties, rank, world, device, dist, **kw):
        import ctypes as C
        import torch
        from . import solver
        self.C, self.torch, self.dist = C, torch, dist
        self.rank, self.world = rank, world
        self.device = torch.device("cuda", device)
        self.m = solver.FemTech(part["coordinates"], part["connectivity"], part["pid"], materialID, properties,
                                comm=part["comm"], world_rank=rank, world_size=world, device=device, **kw)
        self.halo = HaloExchange(part["comm"], torch.device("cuda", device), dist)
        self._send = C.c_void_p(self.halo.send.data_ptr())
        self._recv = C.c_void_p(self.halo.recv.data_ptr())
        # kernels and NCCL ops must be ordered on ONE stream: torch's current stream inside these methods
        self.stream = torch.cuda.Stream(device=device, priority=-1)  # boundary elements + exchange ahead of the interior
        self.m.set_stream(self.stream.cuda_stream)

    def setup(self):
        with self.torch.cuda.stream(self.stream):
            self._setup()

    def explicit_begin(self, energy_every=1):
        with self.torch.cuda.stream(self.stream):
            self._explicit_begin(energy_every)

    def run(self, tMax, steps):
        """Enqueue `steps` time steps (no host synchronisation)."""
        with self.torch.cuda.stream(self.stream):
            self._run(tMax, steps)

    def _setup(self):
        m = self.m
        m.ShapeFunctions()
        m.AssembleLumpedMass()
        if self.halo.count:  # updateMassMatrixNeighbour, Mass3D.cpp:77-125
            m._check(m.L.ftb200_halo_pack(m._h, 1, self._send))
            self.halo.exchange()
            m._check(m.L.ftb200_halo_add(m._h, 1, self._recv))
            m.refresh_mass()
        else:
            self.halo.exchange()

    def _allreduce_dtmin(self, ptr):
        if self.world > 1:
            t = self.torch.as_tensor(_DevDouble(ptr.value), device=self.halo.send.device)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)

    def _explicit_begin(self, energy_every=1):
        m, C = self.m, self.C
        m.sync_in()
        ptr = C.c_void_p()
        m._check(m.L.ftb200_explicit_begin_dt(m._h, float(m.Time), float(m.ExplicitTimeStepReduction),
                                              float(m.FailureTimeStep), int(energy_every), C.byref(ptr)))
        self._allreduce_dtmin(ptr)
        m._check(m.L.ftb200_explicit_begin_force(m._h, self._send))
        self.halo.exchange()
        m._check(m.L.ftb200_explicit_begin_finish(m._h, self._recv))
        m._poll()

    def _run(self, tMax, steps):
        m, C = self.m, self.C
        m._check(m.L.ftb200_run_begin(m._h, float(tMax), int(steps)))
        ptr = C.c_void_p()
        for _ in range(steps):
            m._check(m.L.ftb200_step_begin(m._h, self._send, C.byref(ptr)))
            self.halo.exchange()
            m._check(m.L.ftb200_step_join(m._h))
            self._allreduce_dtmin(ptr)
            m._check(m.L.ftb200_step_end(m._h, self._recv))
            if self.injury:
                self._injury_select()

    # --- injury criteria across partitions: the 95th percentile is a global order statistic (math.cpp:160-199) -------
    injury = False

    def InitInjuryCriterion(self, exclude_pids=(), thresholds=None):
        import torch
        m, C = self.m, self.C
        m.InitInjuryCriterion(exclude_pids, thresholds)
        n = C.c_longlong()
        m._check(m.L.ftb200_injury_local_count(m._h, C.byref(n)))
        with torch.cuda.stream(self.stream):
            t = torch.tensor([n.value], dtype=torch.int64, device=self.device)
            self.dist.all_reduce(t)
            # read back on the SAME stream: torch's streams do not synchronise with the default stream, and a .item()
            # issued there could return this rank's own count before the all-reduce has run (seen on 8 GPUs: the global
            # 95th percentile then used a local rank index and came out too low)
            total = int(t.item())
        m._check(m.L.ftb200_injury_global_count(m._h, total))
        self.injury = True

    def _injury_select(self):
        import torch
        m, C = self.m, self.C
        hp, hn = C.c_void_p(), C.c_int()
        for p in range(m.L.ftb200_injury_passes()):
            m._check(m.L.ftb200_injury_select_hist(m._h, p, C.byref(hp), C.byref(hn)))
            with torch.cuda.stream(self.stream):
                h = torch.as_tensor(_DevInt32(hp.value, hn.value), device=self.device)
                self.dist.all_reduce(h)  # counts < 2^31: int32 sums of the per-rank histograms
            m._check(m.L.ftb200_injury_select_pick(m._h, p))

    def enable_p2p(self, part_comm):
        """Switch the loop to the peer-memory transport: exchange IPC handles and comm patterns once, then
        `run_p2p` drives the whole multi-GPU step from CUDA graphs with no NCCL call per step."""
        m, C, torch, dist = self.m, self.C, self.torch, self.dist
        handle = (C.c_ubyte * 64)()
        m._check(m.L.ftb200_p2p_export(m._h, handle, None))
        mine = {"handle": bytes(handle), "comm": {k: np.asarray(v).tolist() for k, v in part_comm.items()}}
        allr = [None] * self.world
        dist.all_gather_object(allr, mine)
        handles = b"".join(a["handle"] for a in allr)
        comms = [{k: np.asarray(v, dtype=np.int32) for k, v in a["comm"].items()} for a in allr]
        off, idx, tot = p2p_metadata(comms, self.rank)
        buf = (C.c_ubyte * len(handles)).from_buffer_copy(handles)
        ip = C.POINTER(C.c_int)
        m._check(m.L.ftb200_p2p_import(m._h, buf, 0, off.ctypes.data_as(ip), idx.ctypes.data_as(ip), tot.ctypes.data_as(ip)))
        self._p2p_keep = (off, idx, tot, buf)
        dist.barrier()

    def run_p2p(self, tMax, steps):
        with self.torch.cuda.stream(self.stream):
            self.m.run_async(tMax, steps)

    def energy(self):
        """Wint, Wext, WKE summed over ranks (CheckEnergy.cpp:54-64)."""
        e = self.torch.tensor(self.m.energy()[:3], dtype=self.torch.float64, device=self.halo.send.device)
        if self.world > 1:
            self.dist.all_reduce(e)
        e = e.cpu().numpy()
        return np.array([e[0], e[1], e[2], abs(e[2] + e[0] - e[1])])


# ------------------------------------------------------------- in-process group (tests, 1 GPU)
class LocalGroup:
    """P partitions driven in lockstep inside ONE process on one GPU: the same C-ABI call sequence as
    DistFemTech, with the exchange done by device-to-device copies between the ranks' windows.  Used by
    the GPU parity tests to cover the multi-rank kernels (boundary/interior split, pack, neighbour sum,
    cross-rank dt) on a single-GPU box."""

    def __init__(self, parts, materialID, properties, device=0, **kw):
        import ctypes as C
        import torch
        from . import solver
        self.C, self.torch = C, torch
        self.P = len(parts)
        self.models = [solver.FemTech(p["coordinates"], p["connectivity"], p["pid"], materialID, properties,
                                      comm=p["comm"], world_rank=r, world_size=self.P, device=device, **kw)
                       for r, p in enumerate(parts)]
        dev = torch.device("cuda", device)
        self.comms = [p["comm"] for p in parts]
        self.send = [torch.zeros(3 * max(int(c["sendNeighbourCountCum"][-1]), 1), dtype=torch.float64, device=dev)
                     for c in self.comms]
        self.recv = [torch.zeros_like(t) for t in self.send]

    def _ptr(self, t):
        return self.C.c_void_p(t.data_ptr())

    def _sync(self):
        self.torch.cuda.synchronize()

    def _exchange(self):
        self._sync()
        for r, c in enumerate(self.comms):
            cum = c["sendNeighbourCountCum"]
            for i, q in enumerate(c["sendProcessID"]):
                cq = self.comms[q]
                j = int(np.where(cq["sendProcessID"] == r)[0][0])
                lo, hi = 3 * int(cum[i]), 3 * int(cum[i + 1])
                qlo, qhi = 3 * int(cq["sendNeighbourCountCum"][j]), 3 * int(cq["sendNeighbourCountCum"][j + 1])
                assert hi - lo == qhi - qlo
                self.recv[r][lo:hi].copy_(self.send[q][qlo:qhi])
        self._sync()

    def _min_dt(self, ptrs):
        self._sync()
        ts = [self.torch.as_tensor(_DevDouble(p.value), device=self.send[0].device) for p in ptrs]
        mn = self.torch.stack([t[0] for t in ts]).min()
        for t in ts:
            t[0] = mn
        self._sync()

    def setup(self):
        for m in self.models:
            m.ShapeFunctions()
            m.AssembleLumpedMass()
        for r, m in enumerate(self.models):
            m._check(m.L.ftb200_halo_pack(m._h, 1, self._ptr(self.send[r])))
        self._exchange()
        for r, m in enumerate(self.models):
            m._check(m.L.ftb200_halo_add(m._h, 1, self._ptr(self.recv[r])))
            m.refresh_mass()
        self._sync()

    def explicit_begin(self, energy_every=1):
        C = self.C
        ptrs = []
        for m in self.models:
            m.sync_in()
            p = C.c_void_p()
            m._check(m.L.ftb200_explicit_begin_dt(m._h, float(m.Time), float(m.ExplicitTimeStepReduction),
                                                  float(m.FailureTimeStep), int(energy_every), C.byref(p)))
            ptrs.append(p)
        self._min_dt(ptrs)
        for r, m in enumerate(self.models):
            m._check(m.L.ftb200_explicit_begin_force(m._h, self._ptr(self.send[r])))
        self._exchange()
        for r, m in enumerate(self.models):
            m._check(m.L.ftb200_explicit_begin_finish(m._h, self._ptr(self.recv[r])))
            m._poll()

    def run(self, tMax, steps):
        C = self.C
        for m in self.models:
            m._check(m.L.ftb200_run_begin(m._h, float(tMax), int(steps)))
        for _ in range(steps):
            ptrs = []
            for r, m in enumerate(self.models):
                p = C.c_void_p()
                m._check(m.L.ftb200_step_begin(m._h, self._ptr(self.send[r]), C.byref(p)))
                ptrs.append(p)
            self._exchange()
            for m in self.models:
                m._check(m.L.ftb200_step_join(m._h))
            self._min_dt(ptrs)
            for r, m in enumerate(self.models):
                m._check(m.L.ftb200_step_end(m._h, self._ptr(self.recv[r])))
            if self.injury:
                self._injury_select()
        self._sync()
        for m in self.models:
            m._poll()

    injury = False

    def InitInjuryCriterion(self, exclude_pids=(), thresholds=None):
        C = self.C
        total = 0
        for m in self.models:
            m.InitInjuryCriterion(exclude_pids, thresholds)
            n = C.c_longlong()
            m._check(m.L.ftb200_injury_local_count(m._h, C.byref(n)))
            total += n.value
        for m in self.models:
            m._check(m.L.ftb200_injury_global_count(m._h, total))
        self.injury = True

    def _injury_select(self):
        C, torch = self.C, self.torch
        dev = self.send[0].device
        for p in range(self.models[0].L.ftb200_injury_passes()):
            hs = []
            for m in self.models:
                hp, hn = C.c_void_p(), C.c_int()
                m._check(m.L.ftb200_injury_select_hist(m._h, p, C.byref(hp), C.byref(hn)))
                hs.append(torch.as_tensor(_DevInt32(hp.value, hn.value), device=dev))
            self._sync()
            tot = torch.stack(hs).sum(dim=0).to(torch.int32)  # what the all-reduce does
            for h in hs:
                h.copy_(tot)
            self._sync()
            for m in self.models:
                m._check(m.L.ftb200_injury_select_pick(m._h, p))

    def enable_p2p(self):
        """Peer-memory transport between the in-process ranks (device pointers instead of IPC handles)."""
        C = self.C
        wins = []
        for m in self.models:
            w = C.c_void_p()
            m._check(m.L.ftb200_p2p_export(m._h, None, C.byref(w)))
            wins.append(w.value)
        arr = (C.c_void_p * self.P)(*wins)
        ip = C.POINTER(C.c_int)
        self._keep = []
        for r, m in enumerate(self.models):
            off, idx, tot = p2p_metadata(self.comms, r)
            self._keep.append((off, idx, tot))
            m._check(m.L.ftb200_p2p_import(m._h, arr, 1, off.ctypes.data_as(ip), idx.ctypes.data_as(ip), tot.ctypes.data_as(ip)))

    def run_p2p(self, tMax, steps):
        # direct launches (no graph capture/instantiate while a peer's wait kernel may be spinning: all the
        # ranks share one CUDA context here; separate processes use the captured graph)
        # Enqueued in chunks of a few steps, rank by rank: a rank's wait kernel spins until its neighbours' steps arrive, and
        # the host cannot enqueue those while it is blocked on a full launch queue behind the first rank's whole run
        # (mixed-material partitions launch a dozen kernels per step).
        for m in self.models:
            m.profile(True)
        left = int(steps)
        while left > 0:
            c = min(left, 8)
            for m in self.models:
                m.run_async(tMax, c)
            left -= c
        self._sync()
        for m in self.models:
            m._poll()
            m.profile(False)

    def energy(self):
        e = np.sum([m.energy()[:3] for m in self.models], axis=0)
        return np.array([e[0], e[1], e[2], abs(e[2] + e[0] - e[1])])

    def close(self):
        for m in self.models:
            m.close()
