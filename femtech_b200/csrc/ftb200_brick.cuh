// ftb200_brick.cuh -- the brick-fused explicit step (sm_100a).
//
// The two-kernel step (k_elem -> k_node) sends every element's 24 nodal forces through HBM: 192 B written and 192 B
// read per element and step, more than half of the step's traffic.  Here the mesh is cut into BRICKS of at most
// BRICK_NT elements (host: plan_bricks in ftb200_capi.cu; geometric boxes of the element centroids) and the element
// forces never leave the SM:
//
//   k_brick   persistent thread blocks (two per SM), each walking bricks b, b + G, b + 2G, ...; per brick
//             START     first kick + drift + boundary condition of the step for every node the brick touches
//                       (Benchmarking-Parallel.cpp:115-135,184-244), new displacements staged in shared memory;
//             SETUP     per thread = element: nodal gather, displacement modes, dU/dxi columns, (8 J0)^-1, dt factor;
//             LOOP      eight Gauss points: F, material, P cof(J0) -> 24 nodal forces, element dt
//                       (GetForce_3D.cpp:15-46, CalculateTimeStep.cpp:7-21);
//             FINISH    per local node: fixed-order sum of the brick's contributions through shared memory
//                       (GetForce_3D.cpp:39-44).  INTERIOR nodes (every element of the node lies in this brick) are
//                       finished here: a = (fe - fi)/m, second kick, energy partial (CalculateAcclerations.cpp:4-13,
//                       Benchmarking-Parallel.cpp:146-151, CheckEnergy.cpp:19-52), state written once.  SURFACE nodes
//                       get one partial sum per brick (24 B) in a slot of the partial planes.
//   k_surf    per surface node: the partials of its bricks in ascending brick order, then the same start + finish.
//
// What makes it fast is that no warp ever waits for HBM.  A thread-per-element fp64 kernel is bound by the latency of
// its own dependency chains: it needs all 16 warps an SM's registers allow inside the Gauss loop all the time, and a
// block that is loading or storing nodes is not.  So the memory traffic of brick b + G runs UNDER the Gauss loop of
// brick b (a software pipeline inside the persistent block):
//   * the nodal state (u, v, a, flags, X) of the next brick and the metadata of the one after it are copied
//     global -> shared with cp.async (no registers) right before the loop starts; mass and previous internal force of
//     the current brick's interior nodes at its START; the local node -> (element, slot) map with one bulk copy
//     (cp.async.bulk + mbarrier, SASS UBLKCP);
//   * START, SETUP and FINISH then read shared memory only;
//   * shared memory has room for those buffers because the element scratch does not live there: the 36 dU/dxi column
//     entries of every thread sit in TENSOR MEMORY (tcgen05.st / tcgen05.ld, shape 32x32b: a TMEM column is one 32-bit
//     word per thread of the warp's lane quarter, 72 columns per thread, dynamically indexed by the Gauss point; SASS
//     STTM / LDTM).  No MMA is involved -- TMEM is used as 256 KB of per-thread scratch next to the register file.
//     Only (8 J0)^-1 (9 doubles per thread, re-read twice per Gauss point) stays in shared memory: measured TMEM read
//     bandwidth is 94 B/clk/SM against 128 B/clk/SM for shared memory (experiments/tmem_probe.cu), so the 27 scratch
//     loads per Gauss point are split 9 / 18 between the two.
//
// Deterministic: every sum has a fixed order (interior nodes: ascending reference element id, exactly the order of
// the two-kernel step; surface nodes: per brick, then ascending brick id), no floating-point atomics.
//
// Internal node order in brick mode: [interior nodes of brick 0 | of brick 1 | ... | surface nodes by owning brick |
// padding], so the interior nodes of a brick are one contiguous index range (coalesced loads and stores, no index list).
#pragma once
#include "ftb200_kernels.cuh"

namespace ftb {

constexpr int BRICK_NT = 256;      // threads per block = maximum number of elements of a brick
constexpr int BRICK_NIMAX = 160;   // maximum number of interior nodes of a brick (10 x 5 x 5 elements: 147)
constexpr int BRICK_NSMAX = 256;   // maximum number of surface local nodes of a brick (10 x 5 x 5 elements: 249)
constexpr int BRICK_NLMAX = BRICK_NIMAX + BRICK_NSMAX;
constexpr int BRICK_NODE_TRIPS = (BRICK_NLMAX + BRICK_NT - 1) / BRICK_NT;
constexpr int BRICK_TMEM_COLS = 256;  // per block; two blocks per SM use all 512 columns

struct BrickHdr {
  int e0, nEl;      // elements [e0, e0 + nEl) of the internal element order
  int ibase, nInt;  // interior nodes: internal node ids [ibase, ibase + nInt) = local nodes [0, nInt)
  int nLoc;         // local nodes; [nInt, nLoc) are surface nodes, their internal ids in halo[]
  int slot0;        // first partial slot of this brick; local surface node l writes slot0 + (l - nInt)
  int pad0, pad1;
};

struct BrickArgs {
  const BrickHdr* hdr;
  const uint16_t* conn16;  // [nB][BRICK_NT][8] local node index of C3D8 node k of local element t (one 16-byte copy per thread)
  const int* halo;         // [nB][BRICK_NSMAX] internal node ids of the surface local nodes
  const uint16_t* map16;   // [nB][8][BRICK_NLMAX] local node -> row * BRICK_NT + local element of the force a local
                           // element puts on it (row = its C3D8 slot), 0xFFFF = none; ascending reference element id
  const int* pe;           // [nE] part id | (element skipped by StableTimeStep) << 30, internal element order
  const unsigned* flags32; // node flags widened to 32 bits (cp.async moves at least 4 bytes)
  const double* X[3];
  double* u[3];
  double* v[3];
  double* a[3];
  double* fi[3];
  const double* fe[3];     // nullptr planes when the external force is identically zero
  const double* m;
  const double* mp;
  double* part[3];         // partial sums of the surface nodes, one slot per (brick, surface local node)
  double* epart;           // [3][nEpart] energy partials: warps of k_brick (brick * 8 + warp) first, then those of k_surf
  int nEpart;
  int nB;
  DevScalars* sc;
  int store_fi;
};

// one pipeline stage of brick metadata (copied two bricks ahead)
struct BrickMeta {
  uint16_t conn16[BRICK_NT][8];
  int halo[BRICK_NSMAX];
  int pe[BRICK_NT];
};
struct BrickSmem {
  double ji[9][BRICK_NT];           // (8 J0)^-1 of every thread; rows 0..7 double as the force exchange buffer of FINISH
  double ust[3][BRICK_NLMAX];       // new displacements of the local nodes
  double xs[3][BRICK_NLMAX];        // reference coordinates of the local nodes
  double rawI[2][9][BRICK_NIMAX];   // u, v, a of the interior nodes (this brick's are read again by FINISH: two stages)
  unsigned flI[2][BRICK_NIMAX];
  double lateI[4][BRICK_NIMAX];     // m, fi_prev[3] of the current brick's interior nodes
  double rawS[9][BRICK_NSMAX];      // u, v, a of the surface local nodes
  unsigned flS[BRICK_NSMAX];
  uint16_t map16[8][BRICK_NLMAX];
  BrickMeta meta[2];
  BrickHdr hdr[3];
  double times[4];                  // dt1, dt2, dtn, T of the step (StepTimes), re-read behind the Gauss loop
  unsigned long long bar;
  unsigned tmem_base, pad;
};
constexpr size_t BRICK_SMEM_BYTES = sizeof(BrickSmem);
static_assert(sizeof(BrickSmem) <= 115712, "two blocks per SM");

// ---- bulk asynchronous copy global -> shared, completion on an mbarrier (SASS: UBLKCP / SYNCS) ----
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(b), "r"(parity)
        : "memory");
  }
}

// ---- the node update of one step, shared by k_brick (interior nodes) and k_surf (surface nodes) --------------------
struct StepTimes {
  double dt1, dt2, dtn, T;  // first kick t_half - t_n, second kick t_np1 - t_half, drift dt, end time of the step
};
__device__ __forceinline__ StepTimes step_times(const DevScalars* sc) {
  StepTimes t;
  const double tn = sc->nt_n, th = sc->nt_half, t1 = sc->nt_np1;
  t.dt1 = th - tn; t.dt2 = t1 - th; t.dtn = sc->ndt; t.T = t1;
  return t;
}
// START of the step for one node (Benchmarking-Parallel.cpp:115-135 and ApplyBoundaryConditions :184-244): the new
// displacement, and the velocity / acceleration the FINISH of the same step starts from.  Explicit fused
// multiply-adds: every block that touches the node (and k_surf) must produce the same bits.
__device__ __forceinline__ void node_start(const unsigned fl, const StepTimes& t, const double* bc_rate, const double u[3],
                                           const double v[3], const double a[3], double un[3], double vs[3], double as[3]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const bool b = (fl >> c) & 1u;
    const unsigned kind = (fl >> (4 + 2 * c)) & 3u;
    un[c] = b ? u[c] : fma(t.dtn, fma(t.dt1, a[c], v[c]), u[c]);
    vs[c] = v[c];
    as[c] = a[c];
    if (kind) {
      const double r = bc_rate[kind];
      un[c] = t.T * r;
      vs[c] = r;
      as[c] = 0.0;
    }
  }
}
// displacement only (what the elements of a brick need from a node another block finishes)
__device__ __forceinline__ void node_start_u(const unsigned fl, const StepTimes& t, const double* bc_rate, const double u[3],
                                             const double v[3], const double a[3], double un[3]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const bool b = (fl >> c) & 1u;
    const unsigned kind = (fl >> (4 + 2 * c)) & 3u;
    un[c] = b ? u[c] : fma(t.dtn, fma(t.dt1, a[c], v[c]), u[c]);
    if (kind) un[c] = t.T * bc_rate[kind];
  }
}
// FINISH: accelerations, second kick, energy terms of the node (CalculateAcclerations.cpp:7-11,
// Benchmarking-Parallel.cpp:146-151, CheckEnergy.cpp:19-52).  f = assembled internal force.
template <bool ENERGY>
__device__ __forceinline__ void node_finish(const unsigned fl, const StepTimes& t, const double m, const double f[3],
                                            const double fext[3], const double fprev[3], const double u_old[3],
                                            const double un[3], const double vs[3], const double as[3], double vn[3],
                                            double an[3], double& wke, double& wint, double& wext) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const bool b = (fl >> c) & 1u;
    const double fnet = fext[c] - f[c];  // GetForce_3D.cpp:11,49-51
    an[c] = as[c];
    vn[c] = vs[c];
    if (!b) {
      an[c] = fnet / m;
      vn[c] = fma(t.dt2, an[c], fma(t.dt1, as[c], vs[c]));
    }
    if (ENERGY && !(fl & FTB_FLAG_NOTOWNED)) {
      const double dd = un[c] - u_old[c];
      wke += m * vn[c] * vn[c];
      if (b) wext += dd * (fprev[c] + f[c] + m * (an[c] + as[c]));
      wint += dd * (fprev[c] + f[c]);
      wext += dd * (fext[c] + fext[c]);  // fe_prev == fe: the reference never updates fe
    }
  }
}

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// nodal input of hex8_brick_setup: node-indexed shared-memory tables, local node ids packed two per word
struct BrickIn {
  const double* xs;    // [3][BRICK_NLMAX]
  const double* ust;   // [3][BRICK_NLMAX]
  unsigned ln[4];
  __device__ __forceinline__ int id(const int k) const { return (int)((ln[k >> 1] >> (16 * (k & 1))) & 0xFFFFu); }
  __device__ __forceinline__ void getX(const int c, double x[4]) const {
    x[0] = xs[c * BRICK_NLMAX + id(0)]; x[1] = xs[c * BRICK_NLMAX + id(1)];
    x[2] = xs[c * BRICK_NLMAX + id(3)]; x[3] = xs[c * BRICK_NLMAX + id(4)];
  }
  __device__ __forceinline__ void getU(const int c, double nu[8]) const {
#pragma unroll
    for (int k = 0; k < 8; ++k) nu[k] = ust[c * BRICK_NLMAX + id(k)];
  }
};
// Element scratch of k_brick: the dU/dxi columns in tensor memory (this thread's lane, columns t0 + 2 i, t0 + 2 i + 1),
// (8 J0)^-1 in shared memory.  All tcgen05 instructions are .sync.aligned: every lane of the warp executes them together.
struct TmemScratch {
  uint32_t t0;   // TMEM address (lane quarter of the warp << 16 | first column of the warp's slice)
  double* ji;    // &smem.ji[0][threadIdx.x]
  __device__ __forceinline__ void st_col(const int i, const double x) const {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(t0 + 2u * (unsigned)i), "r"((unsigned)b), "r"((unsigned)(b >> 32)) : "memory");
  }
  __device__ __forceinline__ void cols_written() const { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
  __device__ __forceinline__ void ld_cols9(const int idx[9], double out[9]) const {
    unsigned lo[9], hi[9];
#pragma unroll
    for (int k = 0; k < 9; ++k)
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo[k]), "=r"(hi[k]) : "r"(t0 + 2u * (unsigned)idx[k]) : "memory");
    // the loads are asynchronous: the destination registers may be read only behind the wait, so they pass through it
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(lo[0]), "+r"(hi[0]), "+r"(lo[1]), "+r"(hi[1]), "+r"(lo[2]), "+r"(hi[2]), "+r"(lo[3]), "+r"(hi[3]), "+r"(lo[4]),
                   "+r"(hi[4]), "+r"(lo[5]), "+r"(hi[5]), "+r"(lo[6]), "+r"(hi[6]), "+r"(lo[7]), "+r"(hi[7]), "+r"(lo[8]), "+r"(hi[8])
                 :
                 : "memory");
#pragma unroll
    for (int k = 0; k < 9; ++k) out[k] = __hiloint2double((int)hi[k], (int)lo[k]);
  }
  __device__ __forceinline__ void st_ji(const int i, const double x) const { ji[i * BRICK_NT] = x; }
  __device__ __forceinline__ double ld_ji(const int i) const {  // re-read in every iteration: keeps 18 registers free
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(ji + i * BRICK_NT)) : "memory");
    return v;
  }
};

// copies of one brick's metadata / nodal state into a pipeline stage (cp.async: in flight under the Gauss loop)
__device__ __forceinline__ void brick_load_meta(const BrickArgs& A, BrickSmem& S, const int b, const int stage, const int hslot, const int tid) {
  cp_async16(&S.meta[stage].conn16[tid][0], A.conn16 + ((size_t)b * BRICK_NT + tid) * 8);
  cp_async4(&S.meta[stage].halo[tid], A.halo + (size_t)b * BRICK_NSMAX + tid);
  if (tid < 2) cp_async16(reinterpret_cast<char*>(&S.hdr[hslot]) + 16 * tid, reinterpret_cast<const char*>(A.hdr + b) + 16 * tid);
}
__device__ __forceinline__ void brick_load_pe(const BrickArgs& A, BrickSmem& S, const BrickHdr& H, const int stage, const int tid) {
  const int t = tid < H.nEl ? tid : H.nEl - 1;
  cp_async4(&S.meta[stage].pe[tid], A.pe + H.e0 + t);
}
__device__ __forceinline__ void brick_load_nodes(const BrickArgs& A, BrickSmem& S, const BrickHdr& H, const int mstage, const int istage, const int tid) {
#pragma unroll
  for (int r = 0; r < BRICK_NODE_TRIPS; ++r) {
    const int l = tid + r * BRICK_NT;
    if (l < H.nLoc) {
      if (l < H.nInt) {
        const int g = H.ibase + l;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          cp_async8(&S.rawI[istage][c][l], A.u[c] + g);
          cp_async8(&S.rawI[istage][3 + c][l], A.v[c] + g);
          cp_async8(&S.rawI[istage][6 + c][l], A.a[c] + g);
          cp_async8(&S.xs[c][l], A.X[c] + g);
        }
        cp_async4(&S.flI[istage][l], A.flags32 + g);
      } else {
        const int sidx = l - H.nInt;
        const int g = S.meta[mstage].halo[sidx];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          cp_async8(&S.rawS[c][sidx], A.u[c] + g);
          cp_async8(&S.rawS[3 + c][sidx], A.v[c] + g);
          cp_async8(&S.rawS[6 + c][sidx], A.a[c] + g);
          cp_async8(&S.xs[c][l], A.X[c] + g);
        }
        cp_async4(&S.flS[sidx], A.flags32 + g);
      }
    }
  }
}

// per-run refresh of the two derived arrays k_brick copies with cp.async (4-byte granularity): node flags widened to 32 bits,
// part id and the StableTimeStep skip flag of an element in one word
__global__ void k_brick_prep(const uint16_t* __restrict__ flags, unsigned* flags32, int nN, const int* __restrict__ pid,
                             const uint8_t* __restrict__ eflag, int* pe, int nE) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nN) flags32[i] = flags[i];
  if (i < nE) pe[i] = pid[i] | ((int)(eflag[i] != 0) << 30);
}

template <int MATSEL, bool ENERGY>
__global__ void __launch_bounds__(BRICK_NT, 2) k_brick(const BrickArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BrickSmem& S = *reinterpret_cast<BrickSmem*>(smem_raw);
  const DevScalars* sc = A.sc;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int G = gridDim.x;
  int b = blockIdx.x;
  if (sc->last | sc->done) return;
  if (b >= A.nB) return;
  const double* bc_rate = sc->bc_rate;  // read only by nodes that carry a boundary-condition kind
  // ---- one-time set-up of the block: tensor memory, the barrier of the bulk copies, the first two bricks' metadata ----
  if (tid == 32) {
    const StepTimes T0 = step_times(sc);
    S.times[0] = T0.dt1; S.times[1] = T0.dt2; S.times[2] = T0.dtn; S.times[3] = T0.T;
  }
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&S.tmem_base)), "n"(BRICK_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&S.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  brick_load_meta(A, S, b, 0, 0, tid);
  cp_async_commit();
  cp_async_wait<0>();
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  // warps w and w + 4 share the lanes of sub-partition w % 4: each takes half of the block's columns
  const TmemScratch TS{S.tmem_base + ((uint32_t)((w & 3) * 32) << 16) + (uint32_t)((w >> 2) * (BRICK_TMEM_COLS / 2)), &S.ji[0][tid]};
  brick_load_pe(A, S, S.hdr[0], 0, tid);
  brick_load_nodes(A, S, S.hdr[0], 0, 0, tid);
  if (b + G < A.nB) brick_load_meta(A, S, b + G, 1, 1, tid);
  cp_async_commit();

  for (int it = 0; b < A.nB; b += G, ++it) {
    const int cur = it & 1, nxt = cur ^ 1;
    cp_async_wait<0>();
    __syncthreads();  // S1: this brick's nodes and metadata, the next brick's metadata have arrived; the previous FINISH is done
    const BrickHdr H = S.hdr[it % 3];
    // m and fi_prev of the interior nodes (needed by FINISH only), the node -> force map
    for (int l = tid; l < H.nInt; l += BRICK_NT) {
      cp_async8(&S.lateI[0][l], A.m + H.ibase + l);
      if (ENERGY) {
#pragma unroll
        for (int c = 0; c < 3; ++c) cp_async8(&S.lateI[1 + c][l], A.fi[c] + H.ibase + l);
      }
    }
    cp_async_commit();
    if (tid == 0) {
      mbar_expect_tx(&S.bar, 8 * BRICK_NLMAX * 2);
      bulk_g2s(&S.map16[0][0], A.map16 + (size_t)b * 8 * BRICK_NLMAX, 8 * BRICK_NLMAX * 2, &S.bar);
    }
    // ---- START of the step for the brick's local nodes (shared memory in, shared memory out) ------------------------
    {
    const StepTimes T{S.times[0], S.times[1], S.times[2], S.times[3]};
#pragma unroll
    for (int r = 0; r < BRICK_NODE_TRIPS; ++r) {
      const int l = tid + r * BRICK_NT;
      if (l < H.nLoc) {
        double uu[3], vv[3], aa[3], un[3];
        unsigned fl;
        if (l < H.nInt) {
          fl = S.flI[cur][l];
#pragma unroll
          for (int c = 0; c < 3; ++c) { uu[c] = S.rawI[cur][c][l]; vv[c] = S.rawI[cur][3 + c][l]; aa[c] = S.rawI[cur][6 + c][l]; }
        } else {
          const int sidx = l - H.nInt;
          fl = S.flS[sidx];
#pragma unroll
          for (int c = 0; c < 3; ++c) { uu[c] = S.rawS[c][sidx]; vv[c] = S.rawS[3 + c][sidx]; aa[c] = S.rawS[6 + c][sidx]; }
        }
        node_start_u(fl, T, bc_rate, uu, vv, aa, un);
#pragma unroll
        for (int c = 0; c < 3; ++c) S.ust[c][l] = un[c];
      }
    }
    }
    __syncthreads();  // S2: the new displacements are staged
    // ---- SETUP of the element (threads beyond the brick's last element repeat it: the tcgen05 instructions are warp-wide)
    const bool has_el = tid < H.nEl;
    const int pe = S.meta[cur].pe[tid];
    const double* mp = A.mp + (size_t)(pe & 0x3FFFFFFF) * FTB_MP_STRIDE;
    double det, dtk;
    int status;
    {
      const int te = has_el ? tid : H.nEl - 1;
      const uint4 cw = *reinterpret_cast<const uint4*>(&S.meta[cur].conn16[te][0]);
      const BrickIn in{&S.xs[0][0], &S.ust[0][0], {cw.x, cw.y, cw.z, cw.w}};
      status = hex8_brick_setup(in, mp, TS, &det, &dtk);
    }
    __syncthreads();  // S3: xs, rawS, this stage's metadata are free
    // ---- the next brick's nodes and the metadata of the one after it: in flight under the Gauss loop -------------------
    if (b + G < A.nB) {
      const BrickHdr Hn = S.hdr[(it + 1) % 3];
      brick_load_pe(A, S, Hn, nxt, tid);
      brick_load_nodes(A, S, Hn, nxt, nxt, tid);
      if (b + 2 * G < A.nB) brick_load_meta(A, S, b + 2 * G, cur, (it + 2) % 3, tid);
    }
    cp_async_commit();
    // ---- LOOP --------------------------------------------------------------------------------------------------------
    double fe[8][3];
    {
      double d;
      status |= hex8_brick_loop<MATSEL>(MATSEL, mp, true, NoHistory(), NoOutput(), TS, det, dtk, fe, &d);
      double dte = (has_el && !(pe >> 30)) ? d : 1e300;
      unsigned long long bits = dt_to_bits(dte);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t2 = __shfl_xor_sync(0xffffffffu, bits, o);
        bits = t2 < bits ? t2 : bits;
      }
      if (lane == 0) atomicMin(&A.sc->dtmin_bits, bits);
      if (has_el && status) atomicOr(&A.sc->status, status);
    }
    // ---- FINISH: the brick's contributions per local node, one force component at a time through the exchange rows ----
    // node roles of this thread: surface local node nInt + tid, and one run of interior nodes per warp.  (The header is
    // read again here rather than kept in registers across the loop.)
    const volatile BrickHdr& H2 = S.hdr[it % 3];
    const int nInt = H2.nInt, nLoc = H2.nLoc;
    const int per = (nInt + BRICK_NT / 32 - 1) / (BRICK_NT / 32);
    const int li = w * per + lane;                       // interior node of this thread (if lane < per and li < nInt)
    const bool has_i = lane < per && li < nInt;          // (per <= 32: BRICK_NIMAX <= 8 * 32)
    const int ls = nInt + tid;                           // surface node of this thread
    const bool has_s = ls < nLoc;
    double fI[3] = {0.0, 0.0, 0.0}, fS[3] = {0.0, 0.0, 0.0};
    mbar_wait(&S.bar, (unsigned)(it & 1));
    cp_async_wait<1>();  // m and fi_prev of this brick have arrived (the next brick's copies may still be in flight)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int k = 0; k < 8; ++k) S.ji[k][tid] = fe[k][c];
      __syncthreads();
      if (has_i) {
        double f = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {  // ascending reference element id
          const unsigned en = S.map16[q][li];
          if (en != 0xFFFFu) f += (&S.ji[0][0])[en];
        }
        fI[c] = f;
      }
      if (has_s) {
        double f = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const unsigned en = S.map16[q][ls];
          if (en != 0xFFFFu) f += (&S.ji[0][0])[en];
        }
        fS[c] = f;
      }
      __syncthreads();
    }
    double wke = 0.0, wint = 0.0, wext = 0.0;
    if (has_s) {
      const size_t s = (size_t)H2.slot0 + tid;
      A.part[0][s] = fS[0]; A.part[1][s] = fS[1]; A.part[2][s] = fS[2];
    }
    if (has_i) {
      const StepTimes T{S.times[0], S.times[1], S.times[2], S.times[3]};
      const int g = H2.ibase + li;
      const unsigned fl = S.flI[cur][li];
      const double mm = S.lateI[0][li];
      double uo[3], vo[3], ao[3], fprev[3] = {0.0, 0.0, 0.0}, fext[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        uo[c] = S.rawI[cur][c][li]; vo[c] = S.rawI[cur][3 + c][li]; ao[c] = S.rawI[cur][6 + c][li];
        if (ENERGY) fprev[c] = S.lateI[1 + c][li];
        if (A.fe[c]) fext[c] = A.fe[c][g];
      }
      double un[3], vs[3], as[3], vn[3], an[3];
      node_start(fl, T, bc_rate, uo, vo, ao, un, vs, as);
      node_finish<ENERGY>(fl, T, mm, fI, fext, fprev, uo, un, vs, as, vn, an, wke, wint, wext);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        A.u[c][g] = un[c]; A.v[c][g] = vn[c]; A.a[c][g] = an[c];
        if (A.store_fi) A.fi[c][g] = fI[c];
      }
    }
    if (ENERGY) {  // fixed-shape tree inside the warp, one partial per (brick, warp)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        wke += __shfl_down_sync(0xffffffffu, wke, o);
        wint += __shfl_down_sync(0xffffffffu, wint, o);
        wext += __shfl_down_sync(0xffffffffu, wext, o);
      }
      if (lane == 0) {
        const int i = b * (BRICK_NT / 32) + w;
        A.epart[i] = wke;
        A.epart[A.nEpart + i] = wint;
        A.epart[2 * A.nEpart + i] = wext;
      }
    }
  }
  cp_async_wait<0>();
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(S.tmem_base), "n"(BRICK_TMEM_COLS));
}

// ---------------------------------------------------------------------------------------------
// Surface nodes: internal node ids [node0, node0 + nS).  ell[q][s] = partial slot of the q-th brick (ascending brick id)
// that touches surface node s, -1 = none; nodes touched by more than 8 bricks continue in the CSR arrays.
struct SurfArgs {
  const int* ell;      // [8][nS]
  const int* ov_off;   // [nS + 1] entries beyond the eighth (CSR), or nullptr
  const int* ov_ent;
  double* u[3];
  double* v[3];
  double* a[3];
  double* fi[3];
  const double* fe[3];
  const double* m;
  const uint16_t* flags;
  const double* part[3];
  double* epart;
  int nEpart, eoff;    // this kernel's warps write epart[eoff + warp index]
  int node0, nS;
  DevScalars* sc;
  int store_fi;
};
constexpr int SURF_BLOCK = 128;

template <bool ENERGY>
__global__ void __launch_bounds__(SURF_BLOCK, 6) k_surf(const SurfArgs A) {
  const DevScalars* sc = A.sc;
  const int s = blockIdx.x * SURF_BLOCK + threadIdx.x;
  const int g = A.node0 + s;
  unsigned fl = 0;
  double uo[3] = {0.0, 0.0, 0.0}, vo[3] = {0.0, 0.0, 0.0}, ao[3] = {0.0, 0.0, 0.0};
  int ent[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
  if (s < A.nS) {
#pragma unroll
    for (int q = 0; q < 8; ++q) ent[q] = __ldg(A.ell + (size_t)q * A.nS + s);
    fl = A.flags[g];
#pragma unroll
    for (int c = 0; c < 3; ++c) { uo[c] = A.u[c][g]; vo[c] = A.v[c][g]; ao[c] = A.a[c][g]; }
  }
  if (sc->last | sc->done) return;
  double wke = 0.0, wint = 0.0, wext = 0.0;
  if (s < A.nS) {
    double fv[8][3];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int en = ent[q] < 0 ? 0 : ent[q];
#pragma unroll
      for (int c = 0; c < 3; ++c) fv[q][c] = (ent[q] >= 0) ? __ldcg(A.part[c] + en) : 0.0;
    }
    const double mm = A.m[g];
    double fprev[3] = {0.0, 0.0, 0.0}, fext[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (ENERGY) fprev[c] = A.fi[c][g];
      if (A.fe[c]) fext[c] = A.fe[c][g];
    }
    double f[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int c = 0; c < 3; ++c) f[c] += fv[q][c];
    if (A.ov_off)
      for (int j = A.ov_off[s], j1 = A.ov_off[s + 1]; j < j1; ++j) {
        const int en = __ldg(A.ov_ent + j);
#pragma unroll
        for (int c = 0; c < 3; ++c) f[c] += __ldcg(A.part[c] + en);
      }
    const StepTimes T = step_times(sc);
    const double* bc_rate = sc->bc_rate;
    double un[3], vs[3], as[3], vn[3], an[3];
    node_start(fl, T, bc_rate, uo, vo, ao, un, vs, as);
    node_finish<ENERGY>(fl, T, mm, f, fext, fprev, uo, un, vs, as, vn, an, wke, wint, wext);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      A.u[c][g] = un[c]; A.v[c][g] = vn[c]; A.a[c][g] = an[c];
      if (A.store_fi) A.fi[c][g] = f[c];
    }
  }
  if (ENERGY) {  // one partial per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      wke += __shfl_down_sync(0xffffffffu, wke, o);
      wint += __shfl_down_sync(0xffffffffu, wint, o);
      wext += __shfl_down_sync(0xffffffffu, wext, o);
    }
    if ((threadIdx.x & 31) == 0) {
      const int i = A.eoff + blockIdx.x * (SURF_BLOCK / 32) + (threadIdx.x >> 5);
      A.epart[i] = wke;
      A.epart[A.nEpart + i] = wint;
      A.epart[2 * A.nEpart + i] = wext;
    }
  }
}

}  // namespace ftb
