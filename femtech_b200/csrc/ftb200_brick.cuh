// ftb200_brick.cuh -- the brick-fused explicit step (sm_100a).
//
// The two-kernel step (k_elem -> k_node) sends every element's 24 nodal forces through HBM: 192 B written and 192 B
// read per element and step, more than half of the step's traffic.  Here the mesh is cut into BRICKS of at most
// BRICK_NT elements (host: build_bricks in ftb200_capi.cu; geometric boxes of the element centroids), one thread block
// per brick, and the element forces never leave the SM:
//
//   k_brick   prologue  first kick + drift + boundary condition of the step for every node the brick touches
//                       (Benchmarking-Parallel.cpp:115-135,184-244), new displacements staged in shared memory;
//             elements  hex8_element_brick_in per thread: F, material, P cof(J0) -> 24 nodal forces, element dt
//                       (GetForce_3D.cpp:15-46, CalculateTimeStep.cpp:7-21); the forces overwrite the thread's scratch;
//             epilogue  per local node: fixed-order sum of the brick's contributions from shared memory
//                       (GetForce_3D.cpp:39-44).  INTERIOR nodes (every element of the node lies in this brick) are
//                       finished here: a = (fe - fi)/m, second kick, energy partial (CalculateAcclerations.cpp:4-13,
//                       Benchmarking-Parallel.cpp:146-151, CheckEnergy.cpp:19-52), state written once.  SURFACE nodes
//                       get one partial sum per brick (24 B) in a slot of the partial planes.
//   k_surf    per surface node: the partials of its bricks in ascending brick order, then the same start + finish.
//
// Per element and step at 100^3 (10 x 5 x 5 bricks): ~50 B of brick metadata, ~120 B of nodal state read, ~45 B
// written, ~25 B of partials each way, ~100 B in k_surf -- about half of the 717 B of the two-kernel step.
// Deterministic: every sum has a fixed order (interior nodes: ascending reference element id, exactly the order of
// the two-kernel step; surface nodes: per brick, then ascending brick id), no floating-point atomics.
// Brick metadata arrive by bulk asynchronous copies (cp.async.bulk + mbarrier): one descriptor-free TMA transfer per
// array, issued by one thread, no registers.
//
// Internal node order in brick mode: [interior nodes of brick 0 | of brick 1 | ... | surface nodes by owning brick |
// padding], so the interior nodes of a brick are one contiguous index range (coalesced loads and stores, no index list).
#pragma once
#include "ftb200_kernels.cuh"

namespace ftb {

constexpr int BRICK_NT = 256;      // threads per brick = maximum number of elements of a brick
constexpr int BRICK_NLMAX = 416;   // maximum number of local nodes of a brick (10 x 5 x 5 elements: 396)
constexpr int BRICK_NODE_TRIPS = (BRICK_NLMAX + BRICK_NT - 1) / BRICK_NT;

struct BrickHdr {
  int e0, nEl;      // elements [e0, e0 + nEl) of the internal element order
  int ibase, nInt;  // interior nodes: internal node ids [ibase, ibase + nInt) = local nodes [0, nInt)
  int nLoc;         // local nodes; [nInt, nLoc) are surface nodes, their internal ids in halo[]
  int slot0;        // first partial slot of this brick; local surface node l writes slot0 + (l - nInt)
  int pad0, pad1;
};

struct BrickArgs {
  const BrickHdr* hdr;
  const uint16_t* conn16;  // [nB][BRICK_NT][8] local node index of C3D8 node k of local element t (one 16-byte load per thread)
  const int* xid;          // [nE][4] internal node ids of the reference nodes 0, 1, 3, 4 of every element (internal element order)
  const int* halo;         // [nB][BRICK_NLMAX] internal node ids of the surface local nodes
  const uint16_t* map16;   // [nB][8][BRICK_NLMAX] local node -> scratch word 3 * slot * BRICK_NT + local element of the
                           // force a local element puts on it, 0xFFFF = none; ascending reference element id
  const double* X[3];
  double* u[3];
  double* v[3];
  double* a[3];
  double* fi[3];
  const double* fe[3];     // nullptr planes when the external force is identically zero
  const double* m;
  const uint16_t* flags;
  const int* pid;
  const uint8_t* eflag;
  const double* mp;
  double* part[3];         // partial sums of the surface nodes, one slot per (brick, surface local node)
  double* epart;           // [3][nEpart] energy partials: blocks of k_brick first, then those of k_surf
  int nEpart;
  DevScalars* sc;
  int store_fi;
};

constexpr size_t BRICK_SMEM_BYTES = (size_t)FTB_BRICK_SLOTS * BRICK_NT * 8 + 3 * BRICK_NLMAX * 8 + 8 * BRICK_NLMAX * 2 + 16;

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

struct BrickIn {
  const double* scr;   // &scratch[0][threadIdx.x]: reference nodes 0, 1, 3, 4 staged in FTB_BSTAGE_X slots
  const double* ust;   // [3][BRICK_NLMAX] new displacements of the brick's local nodes
  unsigned ln[4];      // local node ids, two per word
  __device__ __forceinline__ void getX(const int c, double x[4]) const {
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] = scr[FTB_BSTAGE_X(k, c) * BRICK_NT];
  }
  __device__ __forceinline__ void getU(const int c, double nu[8]) const {
#pragma unroll
    for (int k = 0; k < 8; ++k) nu[k] = ust[c * BRICK_NLMAX + ((ln[k >> 1] >> (16 * (k & 1))) & 0xFFFFu)];
  }
};
struct SmemScratchBrick {
  double* base;  // &scratch[0][threadIdx.x]
  __device__ __forceinline__ void st(int i, double x) { base[i * BRICK_NT] = x; }
  __device__ __forceinline__ double ld(int i) const { return base[i * BRICK_NT]; }
  __device__ __forceinline__ double ld_inloop(int i) const {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(base + i * BRICK_NT)) : "memory");
    return v;
  }
};

template <int MATSEL, bool ENERGY>
__global__ void __launch_bounds__(BRICK_NT, 2) k_brick(const BrickArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* scr = reinterpret_cast<double*>(smem_raw);                          // [45][NT]
  double* ust = scr + FTB_BRICK_SLOTS * BRICK_NT;                            // [3][NLMAX]
  uint16_t* s_map = reinterpret_cast<uint16_t*>(ust + 3 * BRICK_NLMAX);      // [8][NLMAX]
  unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_map + 8 * BRICK_NLMAX);  // [1]

  const DevScalars* sc = A.sc;
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const BrickHdr H = A.hdr[b];
  const int lastdone = sc->last | sc->done;
  const StepTimes T = step_times(sc);
  const double* bc_rate = sc->bc_rate;  // read only by nodes that carry a boundary-condition kind
  if (lastdone) return;
  // the local node -> (element, slot) map is needed by the epilogue only: one bulk copy, waited for after the elements
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&s_bar[0], 8 * BRICK_NLMAX * 2);
    bulk_g2s(s_map, A.map16 + (size_t)b * 8 * BRICK_NLMAX, 8 * BRICK_NLMAX * 2, &s_bar[0]);
  }
  // ---- prologue: everything that depends on the header only is requested at once --------------------------------
  const bool has_el = tid < H.nEl;
  const int e = H.e0 + tid;
  int p = 0;
  unsigned skip = 0;
  uint4 cw = make_uint4(0, 0, 0, 0);   // local node ids of the element, two per word
  int4 xg = make_int4(0, 0, 0, 0);     // internal ids of its reference nodes 0, 1, 3, 4
  if (has_el) {
    cw = __ldg(reinterpret_cast<const uint4*>(A.conn16) + (size_t)b * BRICK_NT + tid);
    xg = __ldg(reinterpret_cast<const int4*>(A.xid) + e);
    p = __ldg(A.pid + e);
    skip = __ldg(A.eflag + e);
  }
  {
    int g[BRICK_NODE_TRIPS];
#pragma unroll
    for (int r = 0; r < BRICK_NODE_TRIPS; ++r) {
      const int l = tid + r * BRICK_NT;
      g[r] = -1;
      if (l < H.nLoc) g[r] = l < H.nInt ? H.ibase + l : __ldg(A.halo + (size_t)b * BRICK_NLMAX + (l - H.nInt));
    }
    if (has_el) {  // the element's reference nodes 0, 1, 3, 4: global -> scratch, no registers
      const int gx[4] = {xg.x, xg.y, xg.z, xg.w};
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int c = 0; c < 3; ++c) cp_async8(scr + FTB_BSTAGE_X(kk, c) * BRICK_NT + tid, A.X[c] + gx[kk]);
    }
    double uu[BRICK_NODE_TRIPS][3], vv[BRICK_NODE_TRIPS][3], aa[BRICK_NODE_TRIPS][3];
    unsigned fl[BRICK_NODE_TRIPS];
#pragma unroll
    for (int r = 0; r < BRICK_NODE_TRIPS; ++r)
      if (g[r] >= 0) {
        fl[r] = A.flags[g[r]];
#pragma unroll
        for (int c = 0; c < 3; ++c) { uu[r][c] = A.u[c][g[r]]; vv[r][c] = A.v[c][g[r]]; aa[r][c] = A.a[c][g[r]]; }
      }
    // START of the step for the brick's local nodes
#pragma unroll
    for (int r = 0; r < BRICK_NODE_TRIPS; ++r) {
      const int l = tid + r * BRICK_NT;
      if (g[r] >= 0) {
        double un[3];
        node_start_u(fl[r], T, bc_rate, uu[r], vv[r], aa[r], un);
#pragma unroll
        for (int c = 0; c < 3; ++c) ust[c * BRICK_NLMAX + l] = un[c];
      }
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  // ---- elements ----------------------------------------------------------------------------------------------------
  {
    double dte = 1e300;
    int status = 0;
    if (has_el) {
      const double* mp = A.mp + (size_t)p * FTB_MP_STRIDE;
      double fe[8][3];
      double d;
      SmemScratchBrick S{scr + tid};
      BrickIn in{scr + tid, ust, {cw.x, cw.y, cw.z, cw.w}};
      status = hex8_element_brick_in<MATSEL>(in, MATSEL, mp, true, NoHistory(), NoOutput(), S, fe, &d);
      dte = skip ? 1e300 : d;
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) scr[(3 * k + c) * BRICK_NT + tid] = fe[k][c];
    }
    unsigned long long bits = dt_to_bits(dte);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long t2 = __shfl_xor_sync(0xffffffffu, bits, o);
      bits = t2 < bits ? t2 : bits;
    }
    if ((tid & 31) == 0) atomicMin(&A.sc->dtmin_bits, bits);
    if (status) atomicOr(&A.sc->status, status);
  }
  mbar_wait(&s_bar[0], 0);
  __syncthreads();
  // ---- epilogue: assemble the brick's contributions; finish the interior nodes, park the surface partials -------------
  // surface nodes: one thread each, a coalesced store of the partial sum
  for (int l = H.nInt + tid; l < H.nLoc; l += BRICK_NT) {
    double f[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int q = 0; q < 8; ++q) {  // ascending reference element id; + 0.0 for a missing entry is exact
      const unsigned en = s_map[q * BRICK_NLMAX + l];
      if (en != 0xFFFFu) { f[0] += scr[en]; f[1] += scr[en + BRICK_NT]; f[2] += scr[en + 2 * BRICK_NT]; }
    }
    const size_t s = (size_t)H.slot0 + (l - H.nInt);
    A.part[0][s] = f[0]; A.part[1][s] = f[1]; A.part[2][s] = f[2];
  }
  // interior nodes: spread evenly over the warps (a contiguous run of nodes per warp), finished here
  double wke = 0.0, wint = 0.0, wext = 0.0;
  {
    const int w = tid >> 5, lane = tid & 31;
    const int per = (H.nInt + BRICK_NT / 32 - 1) / (BRICK_NT / 32);
    const int l1 = min((w + 1) * per, H.nInt);
    for (int l = w * per + lane; l < l1; l += 32) {
      const int g = H.ibase + l;
      // the node's own state again (this block read it a few microseconds ago)
      const unsigned fl = A.flags[g];
      const double mm = A.m[g];
      double uo[3], vo[3], ao[3], fprev[3] = {0.0, 0.0, 0.0}, fext[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        uo[c] = A.u[c][g]; vo[c] = A.v[c][g]; ao[c] = A.a[c][g];
        if (ENERGY) fprev[c] = A.fi[c][g];
        if (A.fe[c]) fext[c] = A.fe[c][g];
      }
      double f[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const unsigned en = s_map[q * BRICK_NLMAX + l];
        if (en != 0xFFFFu) { f[0] += scr[en]; f[1] += scr[en + BRICK_NT]; f[2] += scr[en + 2 * BRICK_NT]; }
      }
      double un[3], vs[3], as[3], vn[3], an[3];
      node_start(fl, T, bc_rate, uo, vo, ao, un, vs, as);
      node_finish<ENERGY>(fl, T, mm, f, fext, fprev, uo, un, vs, as, vn, an, wke, wint, wext);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        A.u[c][g] = un[c]; A.v[c][g] = vn[c]; A.a[c][g] = an[c];
        if (A.store_fi) A.fi[c][g] = f[c];
      }
    }
  }
  if (ENERGY) {  // fixed-shape tree inside the warp, one partial per warp: no block barrier at the end of the kernel
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      wke += __shfl_down_sync(0xffffffffu, wke, o);
      wint += __shfl_down_sync(0xffffffffu, wint, o);
      wext += __shfl_down_sync(0xffffffffu, wext, o);
    }
    if ((tid & 31) == 0) {
      const int i = b * (BRICK_NT / 32) + (tid >> 5);
      A.epart[i] = wke;
      A.epart[A.nEpart + i] = wint;
      A.epart[2 * A.nEpart + i] = wext;
    }
  }
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
