// ftb200_kernels.cuh -- CUDA kernels of the FemTech explicit step for sm_100a.
//
// Data layout in HBM (all fp64 / int32, SoA, 256-byte aligned planes):
//   nodal   X,u,v,a,fi,du : 3 planes of nNodes doubles each;  m : nNodes;  flags : uint16 per node
//   element conn          : 8 planes of nElements int32 (plane k = k-th C3D8 node of every element)
//           pid           : nElements int32;  eflag : uint8 (1 = every node fully constrained)
//           felem         : 24 planes of nElements doubles (plane 3k+c = force on node k, component c)
//           hist          : 3 arrays x 6 components x 8 Gauss points planes of nElements doubles
//   CSR     node_off[nNodes+1], node_ent[8 nElements] = element*8+slot, ascending REFERENCE element id
// One thread per element (K_elem) / per node (K_node): every global access of a warp is a run of
// consecutive 8-byte words, gathers excepted.
//
// Kernels (SURVEY.md section 2.1 naming):
//   k_elem   K1  fused gather -> F -> material -> B^T sigma element forces (+ K6 per-element dt, block min)
//   k_node   K2+K5  deterministic CSR gather of f_int + a = f/m + both velocity kicks + drift + BC (+K8 energy partials)
//   k_adv        1-thread scalar update: Time, dt = reduction*min, Prony factors, loop control
//   k_energy K8  fixed-order reduction of the energy partials
//   k_mass_* K7  lumped mass
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "hex8_element.cuh"

namespace ftb {

// flags word per node: bit c (0..2) = boundary[3n+c]; bits 4+2c..5+2c = BC kind of dof c (0 = none);
// bit 12 = node is shared with another rank; bit 13 = energy of this node is owned by a lower rank
#define FTB_FLAG_SHARED 0x1000
#define FTB_FLAG_NOTOWNED 0x2000
#define FTB_FLAG_OVERFLOW 0x4000  // node has more than 8 elements: entries 9.. are read from the CSR arrays
#define FTB_FLAG_RIGID 0x0400     // bit 10: node follows the prescribed rigid-body motion (k_rigid_step, DevRigid)

struct DevScalars {
  double Time;                      // end time of the last finished step
  double t_n, t_np1, t_half, dt;    // step being finished
  double nt_n, nt_np1, nt_half, ndt;  // next step
  double tMax, reduction, failure_dt;
  unsigned long long dtmin_bits;    // atomicMin target: bit pattern of a non-negative double
  long long step;                   // finished steps since explicit_begin
  long long steps_left;             // budget of the current run call
  int last;                         // the step being finished is the last of this run
  int done;                         // nothing left to do in this run
  int active;                       // this loop iteration is live (set by k_adv)
  int status;                       // OR of element status bits | 16 = dt below FailureTimeStep
  double Wint, Wext, WKE, Etot;     // CheckEnergy running sums (CheckEnergy.cpp:4-5,60-64)
  double bc_rate[4];
  long long hist_cap;               // capacity of dt/energy history (steps)
  int energy_every;
  double* ring;                     // step ring in mapped pinned host memory (ftb200_step_ring), or nullptr
  long long ring_cap;               // records (8 doubles each)
};

// One record per finished step, written by the device straight into pinned host memory (64 bytes over PCIe, no host
// synchronisation): Time, next dt, finished steps, status bits, Wint, Wext, WKE, |total|.  The step counter is stored
// last, behind a system-scope fence: a host that sees record[2] == k may read the rest of the record of step k.
__device__ __forceinline__ void step_ring_write(DevScalars* sc) {
  double* r = sc->ring;
  if (!r || sc->ring_cap <= 0 || sc->step <= 0) return;
  volatile double* q = r + 8 * ((sc->step - 1) % sc->ring_cap);
  q[2] = -1.0;
  __threadfence_system();
  q[0] = sc->Time; q[1] = sc->ndt; q[3] = (double)sc->status;
  q[4] = sc->Wint; q[5] = sc->Wext; q[6] = sc->WKE; q[7] = sc->Etot;
  __threadfence_system();
  q[2] = (double)sc->step;
}

__device__ __forceinline__ unsigned long long dt_to_bits(double v) {
  // NaN is never selected by the reference (`dtElem < dtMin` is false); a negative dt would be
  // selected and trip FailureTimeStep -> map it to 0.
  if (!(v == v)) return 0x7FF0000000000000ULL;
  if (v < 0.0) return 0ULL;
  return (unsigned long long)__double_as_longlong(v);
}

// Prony history layout: tiles of 32 consecutive elements x 144 values (3 arrays x 6 components x 8 Gauss points): one
// contiguous 36 KB block per tile.  A warp of the thread-per-element kernels still reads 32 consecutive doubles per
// value, but all 144 values of an element now sit in one tile (plane-major [144][E] put them 8 MB apart: one TLB entry
// and one DRAM page per value; measured +2.4 % on the material-5 kernel).
#define FTB_HIDX(j, gp, e) ((((size_t)(e) >> 5) * 144 + (size_t)(j) * 8 + (size_t)(gp)) * 32 + ((size_t)(e) & 31))
// Element force planes, same tiling: 32 consecutive elements x 24 values (8 nodes x 3 components) per 6 KB tile.  K_elem
// writes one contiguous tile per warp, K_node finds the three components of an (element, node) entry 256 B apart
// instead of in three planes 8 MB apart.
#define FTB_FIDX(s3, e) ((((size_t)(e) >> 5) * 24 + (size_t)(s3)) * 32 + ((size_t)(e) & 31))
struct DevHist {
  double* base;  // [3][6][8][E]
  size_t E;
  size_t e;
  __device__ __forceinline__ void load(int gp, GpHistory& g) const {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      g.h1[i] = base[FTB_HIDX(0 * 6 + i, gp, e)];
      g.h2[i] = base[FTB_HIDX(1 * 6 + i, gp, e)];
      g.s0[i] = base[FTB_HIDX(2 * 6 + i, gp, e)];
    }
  }
  __device__ __forceinline__ void store(int gp, const GpHistory& g) const {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      base[FTB_HIDX(0 * 6 + i, gp, e)] = g.h1[i];
      base[FTB_HIDX(1 * 6 + i, gp, e)] = g.h2[i];
      base[FTB_HIDX(2 * 6 + i, gp, e)] = g.s0[i];
    }
  }
};

#ifndef FTB_ELEM_BLOCK
#define FTB_ELEM_BLOCK 64
#endif
// Cache policy of the streams (FTB_STREAM_HINTS): everything the node kernel touches once per step (its own state, the
// node map) and the connectivity of the element kernel is loaded / stored with the evict-first policy (ld.global.cs /
// st.global.cs), so that the lines that ARE reused across the kernel boundary -- the force tiles the element kernel
// wrote last and the node kernel (walking backwards) reads first -- survive longer in the 126 MB L2.
#ifndef FTB_STREAM_HINTS
#define FTB_STREAM_HINTS 1
#endif
template <class T>
__device__ __forceinline__ T ld_stream(const T* p) {
#if FTB_STREAM_HINTS
  return __ldcs(p);
#else
  return *p;
#endif
}
template <class T>
__device__ __forceinline__ void st_stream(T* p, const T v) {
#if FTB_STREAM_HINTS
  __stcs(p, v);
#else
  *p = v;
#endif
}
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
// Prony history of the viscoelastic material with the loads of Gauss point gp + 1 in flight while gp is integrated:
// 18 doubles per point are copied global -> shared (cp.async, no registers) into a two-stage buffer; load(gp) waits
// for its stage and immediately issues the next one.  With 36 history doubles per Gauss point and only 8 warps per
// SM the plain loads left the kernel latency-bound at ~60 % of the HBM roof.
struct DevHistStaged {
  double* base;   // [3][6][8][E]
  size_t E;
  size_t e;
  double* stage;  // &stage_smem[0][threadIdx.x], layout [2][18][FTB_ELEM_BLOCK]
  __device__ __forceinline__ void prefetch(int gp) const {
    double* st = stage + (size_t)(gp & 1) * 18 * FTB_ELEM_BLOCK;
#pragma unroll
    for (int j = 0; j < 18; ++j) cp_async8(st + j * FTB_ELEM_BLOCK, base + FTB_HIDX(j, gp, e));
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  __device__ __forceinline__ void load(int gp, GpHistory& g) const {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    const double* st = stage + (size_t)(gp & 1) * 18 * FTB_ELEM_BLOCK;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      g.h1[i] = st[(0 * 6 + i) * FTB_ELEM_BLOCK];
      g.h2[i] = st[(1 * 6 + i) * FTB_ELEM_BLOCK];
      g.s0[i] = st[(2 * 6 + i) * FTB_ELEM_BLOCK];
    }
    if (gp < 7) prefetch(gp + 1);
  }
  __device__ __forceinline__ void store(int gp, const GpHistory& g) const {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      base[FTB_HIDX(0 * 6 + i, gp, e)] = g.h1[i];
      base[FTB_HIDX(1 * 6 + i, gp, e)] = g.h2[i];
      base[FTB_HIDX(2 * 6 + i, gp, e)] = g.s0[i];
    }
  }
};

template <bool STAGED>
struct HistSel {
  static __device__ __forceinline__ DevHist make(double* b, size_t E, size_t e, double*) { return DevHist{b, E, e}; }
};
template <>
struct HistSel<true> {
  static __device__ __forceinline__ DevHistStaged make(double* b, size_t E, size_t e, double* st) { return DevHistStaged{b, E, e, st}; }
};
constexpr int HIST_STAGE_BYTES = 2 * 18 * FTB_ELEM_BLOCK * (int)sizeof(double);

struct PackArgs;  // peer-memory exchange fused into the element kernels (below, next to P2PArgs)
struct ElemArgs {
  const PackArgs* pk;  // nullptr outside the partitioned loop
  int pk_nEb;          // elements [0, pk_nEb) of the internal order touch shared nodes
  const double* X[3];
  const double* u[3];
  const int* conn;  // 8 planes
  const int* pid;
  const uint8_t* eflag;
  const double* mp;  // per-part parameter blocks
  double* felem;     // 24 planes
  double* hist;      // or nullptr
  DevScalars* sc;
  int nE;            // plane stride
  int e0, e1;        // element range of this launch
  int ignore_loop_flags;
  const uint8_t* etype;    // 1 = C3D4 (nodes in conn planes 0..3), nullptr = all C3D8; internal order
  // injury criteria (k_elem<..., WITH_INJ>), internal element order; see InjState below
  double* inj_ps;          // PS_Old: max principal strain of the previous step in, of this step out (ex5.cpp:1367)
  double* inj_psxsr;       // PSxSRArray (:1368)
  double* inj_smin;        // this step's minimum principal strain (clipped at <= 0)
  double* inj_shear;       // this step's maximum shear strain
  uint8_t* inj_flags;      // FTB_INJ_* bits
  const uint8_t* inj_incl; // 1 = element takes part (its part is not excluded, ex5.cpp:1251-1281)
  double inj_thr[4];       // MPS > thr0, MPS > thr1, PSR > thr2, PSxSR > thr3 (0.15, 0.30, 120, 28 in ex5.cpp:1335-1365)
};
#define FTB_INJ_MPS_LO 1u
#define FTB_INJ_MPS_HI 2u
#define FTB_INJ_PSR 4u
#define FTB_INJ_PSXSR 8u
#define FTB_INJ_LIST95 16u
#define FTB_INJ_LISTX95 32u

#ifndef FTB_ELEM_BLOCK
#define FTB_ELEM_BLOCK 64
#endif
// epilogue of the hexahedron kernels inside the partitioned loop (A.pk != nullptr); every thread of the block calls it
__device__ void elem_p2p_epilogue(const PackArgs* pk, const int* conn, const int nE, const double* felem, const int e, const bool valid);
#ifndef FTB_ELEM_MINBLOCKS
#define FTB_ELEM_MINBLOCKS 6
#endif
constexpr int ELEM_BLOCK = FTB_ELEM_BLOCK;
constexpr int ELEM_MINBLOCKS = FTB_ELEM_MINBLOCKS;
#ifndef FTB_INJ_MINBLOCKS
#define FTB_INJ_MINBLOCKS 5
#endif

// shared-memory scratch of hex8_element: [72][ELEM_BLOCK] doubles, thread t owns column t
struct SmemScratch {
  double* base;  // &sm[0][threadIdx.x]
  __device__ __forceinline__ void st(int i, double x) { base[i * ELEM_BLOCK] = x; }
  __device__ __forceinline__ double ld(int i) const { return base[i * ELEM_BLOCK]; }
};

// Staged gather of k_elem: component 0 of X and u is loaded into registers, components 1 and 2 are copied
// global -> shared with cp.async (no registers) into the column slots they will later overwrite, so all 48 gathers
// of an element are in flight at once although only 16 values occupy registers.  (Left to itself ptxas sinks the
// last third of the loads behind the first butterflies to save registers: a second exposed memory latency.)
struct StagedIn {
  const double* x0;  // [8] component 0, registers
  const double* u0;
  double* base;      // &sm[0][threadIdx.x]
  __device__ __forceinline__ void get(const int c, double nx[8], double nu[8]) const {
    if (c == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) { nx[k] = x0[k]; nu[k] = u0[k]; }
    } else {
      if (c == 1) asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        nx[k] = base[FTB_STAGE_SLOT(0, k, c) * ELEM_BLOCK];
        nu[k] = base[FTB_STAGE_SLOT(1, k, c) * ELEM_BLOCK];
      }
    }
  }
};

// K1 (+K6).  MATSEL >= 0: every element of the launch has that material (no switch).
// material 5 (36 history doubles per Gauss point in flight) and the generic per-element switch need more
// registers than 168: they run with 4 resident blocks per SM instead of 6
template <int MATSEL, bool WITH_FORCE, bool WITH_DT, bool WITH_INJ = false>
__global__ void __launch_bounds__(ELEM_BLOCK, (MATSEL == 5 || MATSEL < 0) ? 4 : (WITH_INJ ? FTB_INJ_MINBLOCKS : ELEM_MINBLOCKS)) k_elem(const ElemArgs A) {
  const int e = A.e0 + blockIdx.x * ELEM_BLOCK + threadIdx.x;
  const size_t E = (size_t)A.nE;
  // the connectivity is requested before the loop-control flags are tested: one exposed latency, not two
  int nd[8];
  int p = 0;
  unsigned skip = 0;
  if (e < A.e1) {
#pragma unroll
    for (int k = 0; k < 8; ++k) nd[k] = __ldg(A.conn + (size_t)k * E + e);
    p = __ldg(A.pid + e);                       // with the connectivity: the parameter block is a dependent load too
    if (WITH_DT) skip = __ldg(A.eflag + e);     // element skipped by StableTimeStep (:13-19); not a late, exposed load
  }
  if (!A.ignore_loop_flags && (A.sc->last | A.sc->done)) return;
  __shared__ double sm_cols[WITH_FORCE ? 72 : 1][ELEM_BLOCK];
  extern __shared__ double sm_hstage[];  // material 5 only: [2][18][ELEM_BLOCK] (dynamic: beyond the 48 KB static limit)
  constexpr bool STAGED_HIST = WITH_FORCE && MATSEL == 5;
  if (STAGED_HIST && e < A.e1) {  // the history address does not depend on the connectivity: first thing in flight
    const DevHistStaged hs{A.hist, E, (size_t)e, sm_hstage + threadIdx.x};
    hs.prefetch(0);
  }
  double dte = 1e300;
  int status = 0;
  if (e < A.e1) {
    double X0[8], U0[8];
    double X[8][3], U[8][3];  // dt-only variant
    if (WITH_FORCE) {
      double* colbase = &sm_cols[0][threadIdx.x];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        X0[k] = __ldg(A.X[0] + nd[k]);
        U0[k] = __ldg(A.u[0] + nd[k]);
      }
#pragma unroll
      for (int c = 1; c < 3; ++c)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          cp_async8(colbase + FTB_STAGE_SLOT(0, k, c) * ELEM_BLOCK, A.X[c] + nd[k]);
          cp_async8(colbase + FTB_STAGE_SLOT(1, k, c) * ELEM_BLOCK, A.u[c] + nd[k]);
        }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          X[k][c] = __ldg(A.X[c] + nd[k]);
          U[k][c] = __ldg(A.u[c] + nd[k]);
        }
    }
    const double* mp = A.mp + (size_t)p * FTB_MP_STRIDE;
    const int mat = (MATSEL >= 0) ? MATSEL : (int)mp[MP_MATID];
    if (WITH_FORCE) {
      double fe[8][3];
      const auto h = HistSel<STAGED_HIST>::make(A.hist, E, (size_t)e, sm_hstage + threadIdx.x);
      double d;
      SmemScratch S{&sm_cols[0][threadIdx.x]};
      if (WITH_INJ) {
        // strain/injury outputs fused into the force kernel: F never leaves the registers
        double cs[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        status = hex8_element_in<MATSEL, WITH_DT>(StagedIn{X0, U0, S.base}, mat, mp, true, h, StrainSink{cs}, S, fe, &d);
        if (A.inj_incl[e]) {  // ex5.cpp:1313-1369, one element of the loop
          double smax, smin, shear;
          principal_strains(cs, &smax, &smin, &shear);
          const double PSR = (smax - A.inj_ps[e]) / A.sc->ndt;  // first-order backward difference, :1347
          const double PSxSR = smax * PSR;
          unsigned f = A.inj_flags[e];
          if (smax > A.inj_thr[0]) f |= FTB_INJ_MPS_LO;
          if (smax > A.inj_thr[1]) f |= FTB_INJ_MPS_HI;
          if (PSR > A.inj_thr[2]) f |= FTB_INJ_PSR;
          if (PSxSR > A.inj_thr[3]) f |= FTB_INJ_PSXSR;
          A.inj_flags[e] = (uint8_t)f;
          A.inj_ps[e] = smax; A.inj_psxsr[e] = PSxSR; A.inj_smin[e] = smin; A.inj_shear[e] = shear;
        }
      } else {
        status = hex8_element_in<MATSEL, WITH_DT>(StagedIn{X0, U0, S.base}, mat, mp, true, h, NoOutput(), S, fe, &d);
      }
      if (WITH_DT) dte = d;
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) A.felem[FTB_FIDX(3 * k + c, e)] = fe[k][c];
    } else {
      // dt only (legacy StableTimeStep, and the pre-pass of explicit_begin)
      double xm[7][3];
      double n[8], g[7];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int k = 0; k < 8; ++k) n[k] = X[k][c] + U[k][c];
        hex_modes(n, g);
#pragma unroll
        for (int m = 0; m < 7; ++m) xm[m][c] = g[m];
      }
      dte = hex_char_length(xm) / mp[MP_CE];
    }
    if (WITH_DT && skip) dte = 1e300;
  }
  if (WITH_DT) {
    unsigned long long b = dt_to_bits(dte);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      unsigned long long t = __shfl_xor_sync(0xffffffffu, b, o);
      b = t < b ? t : b;
    }
    // one atomic per warp, no block barrier; min is order independent: deterministic
    if ((threadIdx.x & 31) == 0) atomicMin(&A.sc->dtmin_bits, b);
  }
  if (status) atomicOr(&A.sc->status, status);
  if (A.pk && e < A.pk_nEb) elem_p2p_epilogue(A.pk, A.conn, A.nE, A.felem, e, e < A.e1);
}

// K_elem for runs of hexahedra whose reference geometry is affine (hex8_element_affine_in: parallelepipeds, e.g. every
// element of a structured or voxel mesh; the host sorts them into their own runs with hex8_is_affine).  Same step semantics as
// k_elem<MATSEL, true, true, WITH_INJ>; what changes is the cost: cof(J0), det J0, J0^-1 once per element, 54 instead of
// 72 scratch doubles per thread, 36 instead of 48 gathers -- which also lets more blocks share an SM.
#ifndef FTB_AFF_MINBLOCKS
#define FTB_AFF_MINBLOCKS 8
#endif
// Scratch of k_elem_affine: cof(J0) and J0^-1 are loop invariants; left to itself the compiler hoists their 18 loads out
// of the Gauss loop and pays 36 registers for it.  ld_inloop() is a volatile shared-memory load, so they are re-read in
// every iteration (18 LDS against ~170 fp64 instructions) and the kernel fits FTB_AFF_MINBLOCKS blocks per SM.
#ifndef FTB_AFF_REREAD
#define FTB_AFF_REREAD 1
#endif
struct SmemScratchAffine {
  double* base;  // &sm[0][threadIdx.x]
  __device__ __forceinline__ void st(int i, double x) { base[i * ELEM_BLOCK] = x; }
  __device__ __forceinline__ double ld(int i) const { return base[i * ELEM_BLOCK]; }
  __device__ __forceinline__ double ld_inloop(int i) const {
#if FTB_AFF_REREAD
    double v;
    // "memory": the slot was written with ordinary stores before the loop; the compiler must not move them past this
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(base + i * ELEM_BLOCK)) : "memory");
    return v;
#else
    return base[i * ELEM_BLOCK];
#endif
  }
};
struct StagedInAffine {
  const double* x0;  // [4] component 0 of nodes 0, 1, 3, 4 (registers)
  const double* u0;  // [8] component 0
  double* base;      // &sm[0][threadIdx.x]
  __device__ __forceinline__ void getX(const int c, double x[4]) const {
    if (c == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) x[k] = x0[k];
    } else {
      if (c == 1) asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
      for (int k = 0; k < 4; ++k) x[k] = base[FTB_ASTAGE_X(k, c) * ELEM_BLOCK];
    }
  }
  __device__ __forceinline__ void getU(const int c, double nu[8]) const {
    if (c == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) nu[k] = u0[k];
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) nu[k] = base[FTB_ASTAGE_U(k, c) * ELEM_BLOCK];
    }
  }
};
#ifndef FTB_AFF_INJ_MINBLOCKS
#define FTB_AFF_INJ_MINBLOCKS 8  // measured at 100^3 with the criteria on: 8 blocks (128 registers, some spills) 218 us, 5-6 blocks (162-168) 238 us
#endif
template <int MATSEL, bool WITH_INJ>
__global__ void __launch_bounds__(ELEM_BLOCK, MATSEL == 5 ? 4 : (WITH_INJ ? FTB_AFF_INJ_MINBLOCKS : FTB_AFF_MINBLOCKS)) k_elem_affine(const ElemArgs A) {
  const int e = A.e0 + blockIdx.x * ELEM_BLOCK + threadIdx.x;
  const size_t E = (size_t)A.nE;
  int nd[8];
  int p = 0;
  unsigned skip = 0;
  if (e < A.e1) {
#pragma unroll
    for (int k = 0; k < 8; ++k) nd[k] = __ldg(A.conn + (size_t)k * E + e);
    p = __ldg(A.pid + e);
    skip = __ldg(A.eflag + e);
  }
  if (!A.ignore_loop_flags && (A.sc->last | A.sc->done)) return;
  __shared__ double sm_cols[FTB_AFFINE_SLOTS][ELEM_BLOCK];
  extern __shared__ double sm_hstage[];
  constexpr bool STAGED_HIST = MATSEL == 5;
  if (STAGED_HIST && e < A.e1) {
    const DevHistStaged hs{A.hist, E, (size_t)e, sm_hstage + threadIdx.x};
    hs.prefetch(0);
  }
  double dte = 1e300;
  int status = 0;
  if (e < A.e1) {
    double X0[4], U0[8];
    double* colbase = &sm_cols[0][threadIdx.x];
    const int nx[4] = {nd[0], nd[1], nd[3], nd[4]};
#pragma unroll
    for (int k = 0; k < 8; ++k) U0[k] = __ldg(A.u[0] + nd[k]);
#pragma unroll
    for (int k = 0; k < 4; ++k) X0[k] = __ldg(A.X[0] + nx[k]);
#pragma unroll
    for (int c = 1; c < 3; ++c) {
#pragma unroll
      for (int k = 0; k < 8; ++k) cp_async8(colbase + FTB_ASTAGE_U(k, c) * ELEM_BLOCK, A.u[c] + nd[k]);
#pragma unroll
      for (int k = 0; k < 4; ++k) cp_async8(colbase + FTB_ASTAGE_X(k, c) * ELEM_BLOCK, A.X[c] + nx[k]);
    }
    const double* mp = A.mp + (size_t)p * FTB_MP_STRIDE;
    double fe[8][3];
    const auto h = HistSel<STAGED_HIST>::make(A.hist, E, (size_t)e, sm_hstage + threadIdx.x);
    double d;
    SmemScratchAffine S{colbase};
    if (WITH_INJ) {
      double cs[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      status = hex8_element_affine_in<MATSEL, true>(StagedInAffine{X0, U0, colbase}, MATSEL, mp, true, h, StrainSink{cs}, S, fe, &d);
      if (A.inj_incl[e]) {  // ex5.cpp:1313-1369, one element of the loop (same as k_elem)
        double smax, smin, shear;
        principal_strains(cs, &smax, &smin, &shear);
        const double PSR = (smax - A.inj_ps[e]) / A.sc->ndt;
        const double PSxSR = smax * PSR;
        unsigned f = A.inj_flags[e];
        if (smax > A.inj_thr[0]) f |= FTB_INJ_MPS_LO;
        if (smax > A.inj_thr[1]) f |= FTB_INJ_MPS_HI;
        if (PSR > A.inj_thr[2]) f |= FTB_INJ_PSR;
        if (PSxSR > A.inj_thr[3]) f |= FTB_INJ_PSXSR;
        A.inj_flags[e] = (uint8_t)f;
        A.inj_ps[e] = smax; A.inj_psxsr[e] = PSxSR; A.inj_smin[e] = smin; A.inj_shear[e] = shear;
      }
    } else {
      status = hex8_element_affine_in<MATSEL, true>(StagedInAffine{X0, U0, colbase}, MATSEL, mp, true, h, NoOutput(), S, fe, &d);
    }
    dte = skip ? 1e300 : d;
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) A.felem[FTB_FIDX(3 * k + c, e)] = fe[k][c];
  }
  unsigned long long b = dt_to_bits(dte);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, b, o);
    b = t < b ? t : b;
  }
  if ((threadIdx.x & 31) == 0) atomicMin(&A.sc->dtmin_bits, b);
  if (status) atomicOr(&A.sc->status, status);
  if (A.pk && e < A.pk_nEb) elem_p2p_epilogue(A.pk, A.conn, A.nE, A.felem, e, e < A.e1);
}

// K_elem for runs of neo-Hookean (MAT 1, the headline configuration) or HGO (MAT 4) parallelepipeds without the strain
// outputs: the element in current-Jacobian form (hex8_element_affine_cj; MAT 1: the linear part of the stress summed over
// the Gauss points in closed form, 9 shared loads and ~95 fp64 instructions per point).  Same gather plan, same outputs
// and step semantics as k_elem_affine<MAT, false>; 44 scratch slots per thread.  FTB200_NH=0 sends such runs through
// k_elem_affine again.
#ifndef FTB_NH_MINBLOCKS
#define FTB_NH_MINBLOCKS 8
#endif
#ifndef FTB_ELEM_PREFETCH
#define FTB_ELEM_PREFETCH 0  // L2 prefetch of the connectivity this many elements ahead (one resident wave = 148 * 8 * 64):
                             // measured 110.8 us on against 110.7 us off (profiles/r02_k_elem_affine_cj_prefetch_register_variants.txt) -- off
#endif
template <int MAT, bool P2P = false>  // P2P: the boundary elements of the partitioned loop (fused exchange, elem_p2p_epilogue)
#ifdef FTB_NH_MAXNREG
__global__ void __maxnreg__(FTB_NH_MAXNREG) k_elem_affine_cj(const ElemArgs A) {
#else
__global__ void __launch_bounds__(ELEM_BLOCK, FTB_NH_MINBLOCKS) k_elem_affine_cj(const ElemArgs A) {
#endif
  const int e = A.e0 + blockIdx.x * ELEM_BLOCK + threadIdx.x;
  const size_t E = (size_t)A.nE;
  int nd[8];
  int p = 0;
  unsigned skip = 0;
  if (e < A.e1) {
#pragma unroll
    for (int k = 0; k < 8; ++k) nd[k] = ld_stream(A.conn + (size_t)k * E + e);
    p = ld_stream(A.pid + e);
    // the flag is only needed at the very end; as an ordinary load the compiler sinks it there and the warp then waits a
    // full DRAM latency per element (6 % of the stall samples, profiles/r02_k_elem_affine_cj_line_profile.txt)
    asm volatile("ld.global.cs.u8 %0, [%1];" : "=r"(skip) : "l"(A.eflag + e));
#if FTB_ELEM_PREFETCH
    // connectivity of the elements one resident wave ahead, into L2: lanes 0-7 / 8-15 fetch the line of plane k that holds
    // the first / last element of this warp's successor, lanes 16-18 its pid and eflag lines
    {
      const int lane = threadIdx.x & 31;
      const size_t ef = (size_t)(e - lane) + FTB_ELEM_PREFETCH + ((lane & 8) ? 31 : 0);
      if (ef < (size_t)A.e1) {
        if (lane < 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.conn + (size_t)(lane & 7) * E + ef));
        else if (lane < 18) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.pid + min(ef + (size_t)((lane & 1) * 31), (size_t)A.e1 - 1)));
        else if (lane == 18) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.eflag + ef));
      }
    }
#endif
  }
  if (!A.ignore_loop_flags && (A.sc->last | A.sc->done)) return;
  __shared__ double sm_cols[FTB_NH_SLOTS][ELEM_BLOCK];
  double dte = 1e300;
  int status = 0;
  if (e < A.e1) {
    double X0[4], U0[8];
    double* colbase = &sm_cols[0][threadIdx.x];
    const int nx[4] = {nd[0], nd[1], nd[3], nd[4]};
#pragma unroll
    for (int k = 0; k < 8; ++k) U0[k] = __ldg(A.u[0] + nd[k]);
#pragma unroll
    for (int k = 0; k < 4; ++k) X0[k] = __ldg(A.X[0] + nx[k]);
#pragma unroll
    for (int c = 1; c < 3; ++c) {
#pragma unroll
      for (int k = 0; k < 8; ++k) cp_async8(colbase + FTB_ASTAGE_U(k, c) * ELEM_BLOCK, A.u[c] + nd[k]);
#pragma unroll
      for (int k = 0; k < 4; ++k) cp_async8(colbase + FTB_ASTAGE_X(k, c) * ELEM_BLOCK, A.X[c] + nx[k]);
    }
    const double* mp = A.mp + (size_t)p * FTB_MP_STRIDE;
    double fe[8][3];
    double d;
    SmemScratchAffine S{colbase};
    status = hex8_element_affine_cj<MAT, true>(StagedInAffine{X0, U0, colbase}, mp, S, fe, &d);
    dte = skip ? 1e300 : d;
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) A.felem[FTB_FIDX(3 * k + c, e)] = fe[k][c];
  }
  unsigned long long b = dt_to_bits(dte);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, b, o);
    b = t < b ? t : b;
  }
  if ((threadIdx.x & 31) == 0) atomicMin(&A.sc->dtmin_bits, b);
  if (status) atomicOr(&A.sc->status, status);
  if (P2P && e < A.pk_nEb) elem_p2p_epilogue(A.pk, A.conn, A.nE, A.felem, e, e < A.e1);
}

// The C3D4 elements of a mixed mesh: they sit in their own index ranges of the internal element
// order, so this kernel sees only tetrahedra and k_elem only hexahedra.  One thread per element, generic material.
constexpr int TET_BLOCK = 128;
template <bool WITH_FORCE, bool WITH_DT, bool WITH_INJ>
__global__ void __launch_bounds__(TET_BLOCK) k_elem_tet(const ElemArgs A) {
  const int e = A.e0 + blockIdx.x * TET_BLOCK + threadIdx.x;
  if (!A.ignore_loop_flags && (A.sc->last | A.sc->done)) return;
  const size_t E = (size_t)A.nE;
  double dte = 1e300;
  int status = 0;
  if (e < A.e1) {
    double X[4][3], U[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int nd = __ldg(A.conn + (size_t)k * E + e);
#pragma unroll
      for (int c = 0; c < 3; ++c) { X[k][c] = __ldg(A.X[c] + nd); U[k][c] = __ldg(A.u[c] + nd); }
    }
    const double* mp = A.mp + (size_t)__ldg(A.pid + e) * FTB_MP_STRIDE;
    const int mat = (int)mp[MP_MATID];
    double fe[4][3], d = 1e300;
    DevHist h{A.hist, E, (size_t)e};
    if (WITH_INJ) {
      double cs[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      status = tet4_element<-1, WITH_DT>(X, U, mat, mp, WITH_FORCE, h, StrainSink{cs}, fe, &d);
      if (A.inj_incl[e]) {  // ex5.cpp:1313-1369 with GaussPoints[e] = 1
        double smax, smin, shear;
        principal_strains(cs, &smax, &smin, &shear, 1);
        const double PSR = (smax - A.inj_ps[e]) / A.sc->ndt;
        const double PSxSR = smax * PSR;
        unsigned f = A.inj_flags[e];
        if (smax > A.inj_thr[0]) f |= FTB_INJ_MPS_LO;
        if (smax > A.inj_thr[1]) f |= FTB_INJ_MPS_HI;
        if (PSR > A.inj_thr[2]) f |= FTB_INJ_PSR;
        if (PSxSR > A.inj_thr[3]) f |= FTB_INJ_PSXSR;
        A.inj_flags[e] = (uint8_t)f;
        A.inj_ps[e] = smax; A.inj_psxsr[e] = PSxSR; A.inj_smin[e] = smin; A.inj_shear[e] = shear;
      }
    } else {
      status = tet4_element<-1, WITH_DT>(X, U, WITH_FORCE ? mat : 0, mp, WITH_FORCE, h, NoOutput(), fe, &d);
    }
    if (WITH_DT) dte = __ldg(A.eflag + e) ? 1e300 : d;
    if (WITH_FORCE) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) A.felem[FTB_FIDX(3 * k + c, e)] = fe[k][c];
    }
  }
  if (WITH_DT) {
    unsigned long long b = dt_to_bits(dte);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      unsigned long long t = __shfl_xor_sync(0xffffffffu, b, o);
      b = t < b ? t : b;
    }
    if ((threadIdx.x & 31) == 0) atomicMin(&A.sc->dtmin_bits, b);
  }
  if (WITH_FORCE && status) atomicOr(&A.sc->status, status);
}

// ---------------------------------------------------------------------------------------------
// Rigid-body prescribed motion of the brain drivers (examples/ex5/ex5.cpp:339-371, :976-1020).
// 12 states y = [omega, r (generator of the rotation quaternion), v, d] driven by six acceleration traces, advanced
// once per time step by one thread with the Dormand-Prince 5(4) step that boost::odeint's runge_kutta_dopri5 takes in
// do_step(sys, y, ydot, Time - dt, dt) (FSAL: ydot in = derivative at t, out = at t + dt); the kinematics of the
// step then sit in this struct and k_node applies them to the flagged nodes.
struct DevRigid {
  double y[12], ydot[12];
  double R[4], Rinv[4];
  double omega[3], alpha[3], vel[3], acc[3], disp[3];
  int size[6], off[6];   // trace k occupies tab_t/tab_v[off[k] .. off[k] + size[k])
  const double* tab_t;
  const double* tab_v;
};
__device__ inline double rb_interp(const DevRigid* rb, int k, double value) {  // math.cpp:99-119
  const int n = rb->size[k];
  const double* x = rb->tab_t + rb->off[k];
  const double* y = rb->tab_v + rb->off[k];
  if (value < x[0]) return 0.0;
  if (value > x[n - 1]) return y[n - 1];
  if (value == x[0]) return y[0];
  int index = 0;
  for (int i = 1; i < n; ++i)
    if (value <= x[i]) { index = i - 1; break; }
  return y[index] + (y[index + 1] - y[index]) * (value - x[index]) / (x[index + 1] - x[index]);
}
__device__ inline void rb_cross(const double* a, const double* b, double* r) {  // math.cpp:50-54
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = -a[0] * b[2] + a[2] * b[0];
  r[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ inline void rb_qmul(const double* q1, const double* q2, double* qr) {  // math.cpp:134-139
  qr[0] = q1[0] * q2[0] - q1[1] * q2[1] - q1[2] * q2[2] - q1[3] * q2[3];
  qr[1] = q1[0] * q2[1] + q1[1] * q2[0] + q1[2] * q2[3] - q1[3] * q2[2];
  qr[2] = q1[0] * q2[2] - q1[1] * q2[3] + q1[2] * q2[0] + q1[3] * q2[1];
  qr[3] = q1[0] * q2[3] + q1[1] * q2[2] - q1[2] * q2[1] + q1[3] * q2[0];
}
__device__ inline void rb_derivatives(const DevRigid* rb, const double* y, double* ydot, double t) {  // ex5.cpp:976-1020
  for (int k = 0; k < 3; ++k) {
    ydot[k] = rb_interp(rb, k, t);
    ydot[6 + k] = rb_interp(rb, 3 + k, t);
    ydot[9 + k] = y[6 + k];
  }
  double r[3] = {y[3], y[4], y[5]};
  double* rdot = &ydot[3];
  const double rMagnitude = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  if (rMagnitude < 1e-10) {
    rdot[0] = 0.5 * y[0]; rdot[1] = 0.5 * y[1]; rdot[2] = 0.5 * y[2];
  } else {
    const double rCotR = rMagnitude / tan(rMagnitude);
    const double omega[3] = {y[0], y[1], y[2]};
    rb_cross(omega, r, rdot);
    for (int i = 0; i < 3; ++i) r[i] = r[i] / rMagnitude;
    const double rDotOmega = r[0] * omega[0] + r[1] * omega[1] + r[2] * omega[2];
    for (int i = 0; i < 3; ++i) rdot[i] = 0.5 * (rdot[i] + rCotR * y[i] + (1.0 - rCotR) * rDotOmega * r[i]);
  }
}
// begin = 1: before the START of a run's first step (skipped when the run is already done);
// begin = 0: after k_adv, before k_node's START of the next step (skipped on a dead or last iteration)
__global__ void k_rigid_step(const DevScalars* sc, DevRigid* rb, const int begin) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (begin ? (sc->done != 0) : (!sc->active || sc->last)) return;
  const double t = sc->nt_n, dt = sc->ndt;  // ex5.cpp:344: from Time - dt over dt
  const double a2 = 1.0 / 5.0, a3 = 3.0 / 10.0, a4 = 4.0 / 5.0, a5 = 8.0 / 9.0;
  const double b21 = 1.0 / 5.0, b31 = 3.0 / 40.0, b32 = 9.0 / 40.0, b41 = 44.0 / 45.0, b42 = -56.0 / 15.0, b43 = 32.0 / 9.0;
  const double b51 = 19372.0 / 6561.0, b52 = -25360.0 / 2187.0, b53 = 64448.0 / 6561.0, b54 = -212.0 / 729.0;
  const double b61 = 9017.0 / 3168.0, b62 = -355.0 / 33.0, b63 = 46732.0 / 5247.0, b64 = 49.0 / 176.0, b65 = -5103.0 / 18656.0;
  const double c1 = 35.0 / 384.0, c3 = 500.0 / 1113.0, c4 = 125.0 / 192.0, c5 = -2187.0 / 6784.0, c6 = 11.0 / 84.0;
  double x[12], k1[12], xt[12], k2[12], k3[12], k4[12], k5[12], k6[12];
  for (int i = 0; i < 12; ++i) { x[i] = rb->y[i]; k1[i] = rb->ydot[i]; }
  for (int i = 0; i < 12; ++i) xt[i] = x[i] + dt * b21 * k1[i];
  rb_derivatives(rb, xt, k2, t + dt * a2);
  for (int i = 0; i < 12; ++i) xt[i] = x[i] + dt * b31 * k1[i] + dt * b32 * k2[i];
  rb_derivatives(rb, xt, k3, t + dt * a3);
  for (int i = 0; i < 12; ++i) xt[i] = x[i] + dt * b41 * k1[i] + dt * b42 * k2[i] + dt * b43 * k3[i];
  rb_derivatives(rb, xt, k4, t + dt * a4);
  for (int i = 0; i < 12; ++i) xt[i] = x[i] + dt * b51 * k1[i] + dt * b52 * k2[i] + dt * b53 * k3[i] + dt * b54 * k4[i];
  rb_derivatives(rb, xt, k5, t + dt * a5);
  for (int i = 0; i < 12; ++i)
    xt[i] = x[i] + dt * b61 * k1[i] + dt * b62 * k2[i] + dt * b63 * k3[i] + dt * b64 * k4[i] + dt * b65 * k5[i];
  rb_derivatives(rb, xt, k6, t + dt);
  for (int i = 0; i < 12; ++i) x[i] = x[i] + dt * c1 * k1[i] + dt * c3 * k3[i] + dt * c4 * k4[i] + dt * c5 * k5[i] + dt * c6 * k6[i];
  rb_derivatives(rb, x, k1, t + dt);
  for (int i = 0; i < 12; ++i) { rb->y[i] = x[i]; rb->ydot[i] = k1[i]; }
  // ex5.cpp:345-350: R = exp(r) (math.cpp:122-132), its inverse (:141-145), the vectors of the kinematics
  const double q1[4] = {0.0, x[3], x[4], x[5]};
  double R[4];
  double vMag = q1[1] * q1[1] + q1[2] * q1[2] + q1[3] * q1[3];
  if (vMag == 0) {
    R[0] = 1.0; R[1] = 0.0; R[2] = 0.0; R[3] = 0.0;
  } else {
    vMag = sqrt(vMag);
    const double d1 = exp(q1[0]);
    const double d2 = d1 * sin(vMag) / vMag;
    R[0] = d1 * cos(vMag); R[1] = d2 * q1[1]; R[2] = d2 * q1[2]; R[3] = d2 * q1[3];
  }
  const double norm = R[0] * R[0] + R[1] * R[1] + R[2] * R[2] + R[3] * R[3];
  rb->R[0] = R[0]; rb->R[1] = R[1]; rb->R[2] = R[2]; rb->R[3] = R[3];
  rb->Rinv[0] = R[0] / norm; rb->Rinv[1] = -R[1] / norm; rb->Rinv[2] = -R[2] / norm; rb->Rinv[3] = -R[3] / norm;
  for (int j = 0; j < 3; ++j) {
    rb->omega[j] = x[j]; rb->alpha[j] = k1[j]; rb->vel[j] = x[6 + j]; rb->acc[j] = k1[6 + j]; rb->disp[j] = x[9 + j];
  }
}

// ---------------------------------------------------------------------------------------------
struct NodeArgs {
  const int* ell;       // fixed-width node -> (element, slot) map [8][nN], or nullptr (CSR loop)
  const double* X[3];   // reference coordinates (rigid-body nodes only)
  const DevRigid* rigid;  // or nullptr
  double* aprev[3];     // previous acceleration of the rigid-body nodes (energy check), or nullptr planes
  double* u[3];
  double* v[3];
  double* a[3];
  double* fi[3];        // written when store_fi
  double* du[3];        // energy only
  const double* fe[3];  // nullptr planes when the external force is identically zero
  const double* m;
  uint16_t* flags;
  const double* felem;
  const int* node_off;
  const int* node_ent;
  const double* halo_recv;   // device recv window or nullptr
  const double* halo_recv_alt;  // peer-memory transport: second buffer; the step parity selects (nullptr = static)
  const unsigned long long* p2p_seq;
  const int* halo_off;       // per shared node CSR into halo_slot (ascending neighbour order)
  const int* halo_slot;
  const int* halo_node_idx;  // node -> index into halo_off (or -1), nullptr when no halo
  double* epart;             // [3][gridDim.x] energy partials
  DevScalars* sc;
  int nN, nE;
  int store_fi;
};

// k_node launch shape: 128-thread blocks, 8 per SM (64 registers, 32 warps/SM).  Measured at 100^3 (k_node with the energy
// check): 256 threads x 3 blocks (80 registers) 99.8 us, 256 x 4 (64) 98.2 us, 128 x 8 (64) 96.5 us.
#ifndef FTB_NODE_BLOCK
#define FTB_NODE_BLOCK 128
#endif
#ifndef FTB_NODE_MINBLOCKS
#define FTB_NODE_MINBLOCKS 8
#endif
constexpr int NODE_BLOCK = FTB_NODE_BLOCK;

// K2 + K5 (+ K8 partials).  FINISH: gather fi, a = (fe-fi)/m, second kick.  START: first kick of the
// next step, drift, boundary condition.  KICK2 = false for step 0 (accelerations only).
struct DevScalars;
__device__ __forceinline__ double adv_step(DevScalars* sc, double* dt_hist);
__device__ __forceinline__ void prony_update(double* mp, int nPID, double dt, int tid, int nthreads);
template <bool FINISH, bool START, bool KICK2, bool ENERGY>
__global__ void __launch_bounds__(NODE_BLOCK, FTB_NODE_MINBLOCKS) k_node(const NodeArgs A) {
  DevScalars* sc = A.sc;
  // Blocks walk the nodes from the END of the internal order (FTB_NODE_REVERSE): the element kernel before this launch
  // wrote the force tiles in ascending element order, so the tiles of the last elements are the ones still in the
  // 126 MB L2 when this kernel starts -- and the displacements this kernel writes last (the first nodes) are the ones
  // the next element kernel gathers first.  Results are bit-identical: lb is only the block's place in the node order.
#ifndef FTB_NODE_REVERSE
#define FTB_NODE_REVERSE 1
#endif
  const int lb = FTB_NODE_REVERSE ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
  const int n = lb * NODE_BLOCK + threadIdx.x;
  // Preamble: the node's own state (last written by the previous node kernel) and its entries of the static node ->
  // element map.  Nothing here is written by the element kernel or k_adv of the current step, so under a programmatic
  // dependent launch these loads are in flight while those two finish.
  unsigned fl = 0;
  double uu[3] = {0.0, 0.0, 0.0}, vv[3] = {0.0, 0.0, 0.0}, aa[3] = {0.0, 0.0, 0.0};
  int ent[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
  if (n < A.nN) {
    fl = ld_stream(A.flags + n);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uu[c] = ld_stream(A.u[c] + n);
      vv[c] = ld_stream(A.v[c] + n);
      aa[c] = ld_stream(A.a[c] + n);
    }
    if (FINISH && A.ell) {
#pragma unroll
      for (int q = 0; q < 8; ++q) ent[q] = ld_stream(A.ell + (size_t)q * A.nN + n);
    }
  }
  if (FINISH && !sc->active) return;
  if (!FINISH && (sc->done | sc->last)) return;  // START alone: at the beginning of a run
  const bool last_new = sc->last != 0;
  const bool do_start = START && !(FINISH && last_new);
  double wke = 0.0, wint = 0.0, wext = 0.0;
  if (n < A.nN) {
    if (FINISH) {
      // deterministic assembly: ascending element id (GetForce_3D.cpp:15,39-44)
      double f[3] = {0.0, 0.0, 0.0};
      const size_t E = (size_t)A.nE;
      if (A.ell) {
        // fixed-width map [8][nN] (-1 = no entry): the 8 entries and then all 24 force loads are issued before the
        // first add, instead of a dependent load per trip of a variable-length loop; same ascending order
        // (the 8 entries were loaded in the preamble)
#ifndef FTB_NODE_GATHER_BATCH
#define FTB_NODE_GATHER_BATCH 4  // measured at 100^3: 8 in flight (64 registers, 64 B of spills) 92.3 us, 4 (16 B) 90.1 us, 2 89.9 us
#endif
        // FTB_NODE_GATHER_BATCH entries (x 3 components) in flight at a time; the additions keep the ascending order
#pragma unroll
        for (int q0 = 0; q0 < 8; q0 += FTB_NODE_GATHER_BATCH) {
          double fv[FTB_NODE_GATHER_BATCH][3];
#pragma unroll
          for (int q = 0; q < FTB_NODE_GATHER_BATCH; ++q) {
            const int en = ent[q0 + q] < 0 ? 0 : ent[q0 + q];
#pragma unroll
            for (int c = 0; c < 3; ++c) fv[q][c] = (ent[q0 + q] >= 0) ? __ldg(A.felem + FTB_FIDX(3 * (en & 7) + c, en >> 3)) : 0.0;
          }
#pragma unroll
          for (int q = 0; q < FTB_NODE_GATHER_BATCH; ++q)  // + 0.0 for a missing entry is exact
#pragma unroll
            for (int c = 0; c < 3; ++c) f[c] += fv[q][c];
        }
        if (fl & FTB_FLAG_OVERFLOW)
          for (int j = A.node_off[n] + 8, j1 = A.node_off[n + 1]; j < j1; ++j) {
            const int en = __ldg(A.node_ent + j);
#pragma unroll
            for (int c = 0; c < 3; ++c) f[c] += __ldg(A.felem + FTB_FIDX(3 * (en & 7) + c, en >> 3));
          }
      } else {
        const int j0 = A.node_off[n], j1 = A.node_off[n + 1];
        for (int j = j0; j < j1; ++j) {
          const int ent = __ldg(A.node_ent + j);
          const size_t e = (size_t)(ent >> 3);
          const int s = ent & 7;
#pragma unroll
          for (int c = 0; c < 3; ++c) f[c] += __ldg(A.felem + FTB_FIDX(3 * s + c, e));
        }
      }
      if (A.halo_node_idx && (fl & FTB_FLAG_SHARED)) {  // shared node: add the neighbours' partial sums, ascending neighbour (:92-97)
        const int h = A.halo_node_idx[n];  // (only the ~3 % shared nodes pay for this dependent load)
        if (h >= 0) {
          // peer-memory transport: the window that was filled during this step ((seq - 1) & 1, seq already advanced)
          const double* rv = A.halo_recv;
          if (A.halo_recv_alt && ((*A.p2p_seq - 1) & 1ULL)) rv = A.halo_recv_alt;
          for (int j = A.halo_off[h]; j < A.halo_off[h + 1]; ++j) {
            const int slot = A.halo_slot[j];
#pragma unroll
            for (int c = 0; c < 3; ++c) f[c] += __ldcg(rv + 3 * (size_t)slot + c);
          }
        }
      }
      const double m = ld_stream(A.m + n);
      const double dt1 = sc->t_half - sc->t_n, dt2 = sc->t_np1 - sc->t_half;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const bool b = (fl >> c) & 1u;
        const double fext = A.fe[c] ? A.fe[c][n] : 0.0;
        const double fnet = fext - f[c];  // GetForce_3D.cpp:11,49-51
        // accelerations_prev of the energy check: for a rigid-body node the START of this step replaced a, the value
        // before it was parked in aprev (ex5.cpp:209 memcpy before :222 ApplyAccBoundaryConditions)
        const double a_old = (A.rigid && (fl & FTB_FLAG_RIGID)) ? A.aprev[c][n] : aa[c];
        if (!b) aa[c] = fnet / m;  // CalculateAcclerations.cpp:7-11
        if (KICK2) {
          // displacement increment of the step for the energy check (displacements - displacements_prev, CheckEnergy.cpp:35):
          // rebuilt from what the START of this step did instead of being carried through HBM (48 B per node and step) --
          // a free dof moved by dt * v_half, a prescribed one by rate * (t_np1 - t_n) as the difference of the two stored
          // values, a held one not at all; only rigid-body nodes (arbitrary kinematics) keep their stored increment
          double dd = 0.0;
          if (!b) {
            const double vhalf = vv[c] + dt1 * a_old;  // Benchmarking-Parallel.cpp:115-122
            vv[c] = vhalf + dt2 * aa[c];               // :146-151
            if (ENERGY) dd = sc->dt * vhalf;
          } else if (ENERGY) {
            const unsigned kind = (fl >> (4 + 2 * c)) & 3u;
            if (A.rigid && (fl & FTB_FLAG_RIGID)) dd = A.du[c][n];
            else if (kind) dd = sc->t_np1 * sc->bc_rate[kind] - sc->t_n * sc->bc_rate[kind];
          }
          if (ENERGY && !(fl & FTB_FLAG_NOTOWNED)) {   // CheckEnergy.cpp:19-52
            const double fprev = ld_stream(A.fi[c] + n);
            wke += m * vv[c] * vv[c];
            if (b) wext += dd * (fprev + f[c] + m * (aa[c] + a_old));
            wint += dd * (fprev + f[c]);
            wext += dd * (fext + fext);  // fe_prev == fe: the reference never updates fe
          }
        }
        if (A.store_fi) st_stream(A.fi[c] + n, f[c]);
      }
    }
    if (do_start) {
      const double dt1 = sc->nt_half - sc->nt_n, dtn = sc->ndt, T = sc->nt_np1;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const bool b = (fl >> c) & 1u;
        const unsigned kind = (fl >> (4 + 2 * c)) & 3u;
        const double u_old = uu[c];
        if (!b) {
          const double vhalf = vv[c] + dt1 * aa[c];
          uu[c] = u_old + dtn * vhalf;  // :131-135
        }
        if (kind) {  // ApplyBoundaryConditions, :184-244
          const double r = sc->bc_rate[kind];
          uu[c] = T * r;
          vv[c] = r;
          aa[c] = 0.0;
        }
      }
      if (A.rigid && (fl & FTB_FLAG_RIGID)) {  // ApplyAccBoundaryConditions, ex5.cpp:352-371
        const DevRigid* rb = A.rigid;
        const double locV[3] = {A.X[0][n], A.X[1][n], A.X[2][n]};
        const double V[4] = {0.0, locV[0], locV[1], locV[2]};
        double Rv[4], Vp[4], omegaR[3], omegaOmegaR[3], omegaVel[3], alphaR[3];
        rb_qmul(rb->R, V, Rv);
        rb_qmul(Rv, rb->Rinv, Vp);  // Vp = R V R^-1
        rb_cross(rb->omega, &Vp[1], omegaR);
        rb_cross(rb->omega, omegaR, omegaOmegaR);
        rb_cross(rb->omega, rb->vel, omegaVel);
        rb_cross(rb->alpha, &Vp[1], alphaR);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const double u_old = uu[j];
          A.aprev[j][n] = aa[j];
          uu[j] = Vp[j + 1] - locV[j] + rb->disp[j];
          vv[j] = omegaR[j] + rb->vel[j];
          aa[j] = 2.0 * omegaVel[j] + omegaOmegaR[j] + rb->acc[j] + alphaR[j];
          if (ENERGY) A.du[j][n] = uu[j] - u_old;
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (do_start) A.u[c][n] = uu[c];  // default policy: the next element kernel gathers it
      st_stream(A.v[c] + n, vv[c]);
      st_stream(A.a[c] + n, aa[c]);
    }
  }
  if (FINISH && KICK2 && ENERGY) {
    // fixed-shape tree: warp shuffle, then shared memory, one partial per block
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      wke += __shfl_down_sync(0xffffffffu, wke, o);
      wint += __shfl_down_sync(0xffffffffu, wint, o);
      wext += __shfl_down_sync(0xffffffffu, wext, o);
    }
    __shared__ double sw[3][NODE_BLOCK / 32];
    if ((threadIdx.x & 31) == 0) {
      sw[0][threadIdx.x >> 5] = wke;
      sw[1][threadIdx.x >> 5] = wint;
      sw[2][threadIdx.x >> 5] = wext;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
      for (int w = 0; w < NODE_BLOCK / 32; ++w) {
        s0 += sw[0][w];
        s1 += sw[1][w];
        s2 += sw[2][w];
      }
      A.epart[lb] = s0;
      A.epart[gridDim.x + lb] = s1;
      A.epart[2 * gridDim.x + lb] = s2;
    }
  }
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// one loop iteration's scalar bookkeeping (thread 0 of one block); returns the next dt

__device__ __forceinline__ double adv_step(DevScalars* sc, double* dt_hist) {
  double dtmin = __longlong_as_double((long long)*(volatile unsigned long long*)&sc->dtmin_bits);
  if (dtmin > 1e20) dtmin = 1e20;  // `huge`, GlobalVariables.h:16
  sc->dtmin_bits = 0x7FF0000000000000ULL;
  sc->t_n = sc->nt_n; sc->t_np1 = sc->nt_np1; sc->t_half = sc->nt_half; sc->dt = sc->ndt;
  sc->Time = sc->t_np1;
  if (dt_hist && sc->step < sc->hist_cap) dt_hist[sc->step] = sc->dt;
  sc->step += 1;
  sc->steps_left -= 1;
  if (dtmin < sc->failure_dt) { sc->status |= 16; sc->last = 1; }  // TerminateFemTech(19)
  const double ndt = sc->reduction * dtmin;
  sc->ndt = ndt;
  sc->nt_n = sc->Time;
  sc->nt_np1 = sc->Time + ndt;                    // t_np1 = Time + dt
  sc->nt_half = 0.5 * (sc->nt_np1 + sc->nt_n);    // t_nphalf = 0.5*(t_np1 + t_n)
  if (!(sc->Time < sc->tMax) || sc->steps_left <= 0) sc->last = 1;
  if (!sc->energy_every) step_ring_write(sc);
  return ndt;
}
__device__ __forceinline__ void prony_update(double* mp, int nPID, double dt, int tid, int nthreads) {
  for (int p = tid; p < nPID; p += nthreads) {  // HGOIsotropicViscoelastic.cpp:126-131
    double* q = mp + (size_t)p * FTB_MP_STRIDE;
    if ((int)q[MP_MATID] == 5) {
      const double rt1 = dt / q[MP_T1], rt2 = dt / q[MP_T2];
      const double c11 = exp(-rt1), c12 = exp(-rt2);
      q[MP_C11] = c11;
      q[MP_C12] = c12;
      q[MP_C21] = q[MP_G1] * (1 - c11) / rt1;
      q[MP_C22] = q[MP_G2] * (1 - c12) / rt2;
    }
  }
}

// Scalar bookkeeping of the time loop (Benchmarking-Parallel.cpp:106-112,168 and
// StableTimeStep.cpp:33-38).  One thread block; thread p < nPID refreshes the Prony factors of part p
// for the next dt (HGOIsotropicViscoelastic.cpp:126-131).
template <bool INIT>
__global__ void k_adv(DevScalars* sc, double* mp, int nPID, double Time0, double* dt_hist) {
  __shared__ double s_ndt;
  __shared__ int s_live;
  if (threadIdx.x == 0) {
    int live = 1;
    if (!INIT) {
      if (sc->done) live = 0;
      else if (sc->last) { sc->done = 1; live = 0; }
    }
    sc->active = live;
    if (live) {
      double dtmin = __longlong_as_double((long long)sc->dtmin_bits);
      if (dtmin > 1e20) dtmin = 1e20;  // `huge`, GlobalVariables.h:16
      sc->dtmin_bits = 0x7FF0000000000000ULL;
      if (INIT) {
        sc->Time = Time0;
        sc->t_n = sc->t_np1 = sc->t_half = Time0;
        sc->dt = 0.0;
      } else {
        sc->t_n = sc->nt_n; sc->t_np1 = sc->nt_np1; sc->t_half = sc->nt_half; sc->dt = sc->ndt;
        sc->Time = sc->t_np1;
        if (dt_hist && sc->step < sc->hist_cap) dt_hist[sc->step] = sc->dt;
        sc->step += 1;
        sc->steps_left -= 1;
      }
      if (dtmin < sc->failure_dt) { sc->status |= 16; sc->last = 1; }  // TerminateFemTech(19)
      const double ndt = sc->reduction * dtmin;
      sc->ndt = ndt;
      sc->nt_n = sc->Time;
      sc->nt_np1 = sc->Time + ndt;                    // t_np1 = Time + dt
      sc->nt_half = 0.5 * (sc->nt_np1 + sc->nt_n);    // t_nphalf = 0.5*(t_np1 + t_n)
      if (!INIT && (!(sc->Time < sc->tMax) || sc->steps_left <= 0)) sc->last = 1;
      if (!INIT && !sc->energy_every) step_ring_write(sc);  // with the energy check on, k_energy writes the record
      s_ndt = ndt;
    }
    s_live = live;
  }
  __syncthreads();
  if (!s_live) return;
  for (int p = threadIdx.x; p < nPID; p += blockDim.x) {
    double* q = mp + (size_t)p * FTB_MP_STRIDE;
    if ((int)q[MP_MATID] == 5) {
      const double rt1 = s_ndt / q[MP_T1], rt2 = s_ndt / q[MP_T2];
      const double c11 = exp(-rt1), c12 = exp(-rt2);
      q[MP_C11] = c11;
      q[MP_C12] = c12;
      q[MP_C21] = q[MP_G1] * (1 - c11) / rt1;
      q[MP_C22] = q[MP_G2] * (1 - c12) / rt2;
    }
  }
}

// Prony factors for an explicit dt (legacy GetForce: the driver's global dt)
__global__ void k_prony(double* mp, int nPID, double dt) {
  for (int p = threadIdx.x; p < nPID; p += blockDim.x) {
    double* q = mp + (size_t)p * FTB_MP_STRIDE;
    if ((int)q[MP_MATID] == 5) {
      const double rt1 = dt / q[MP_T1], rt2 = dt / q[MP_T2];
      const double c11 = exp(-rt1), c12 = exp(-rt2);
      q[MP_C11] = c11;
      q[MP_C12] = c12;
      q[MP_C21] = q[MP_G1] * (1 - c11) / rt1;
      q[MP_C22] = q[MP_G2] * (1 - c12) / rt2;
    }
  }
}

// scalars of the last finished step as 8 doubles (ftb200_explicit_poll_async)
__global__ void k_scalars_out(const DevScalars* sc, double* out8) {
  out8[0] = sc->Time; out8[1] = sc->ndt; out8[2] = (double)sc->step; out8[3] = (double)sc->status;
  out8[4] = sc->Wint; out8[5] = sc->Wext; out8[6] = sc->WKE; out8[7] = sc->Etot;
}

__global__ void k_begin_run(DevScalars* sc, double tMax, long long steps) {
  sc->tMax = tMax;
  sc->steps_left = steps;
  sc->last = 0;
  sc->active = 0;
  sc->done = (!(sc->Time < tMax) || steps <= 0 || (sc->status & 16)) ? 1 : 0;
}

// K8: fixed-order reduction of the per-block partials, running sums as in CheckEnergy.cpp:60-64
__global__ void k_energy(DevScalars* sc, const double* epart, int nblocks, double* ehist) {
  if (!sc->active) return;
  __shared__ double sh[3][256];
  double s[3] = {0, 0, 0};
  int i = threadIdx.x;
  for (; i + 3 * 256 < nblocks; i += 4 * 256) {  // four independent loads per operand in flight; same summation order
    double q[3][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      q[0][j] = epart[i + j * 256];
      q[1][j] = epart[nblocks + i + j * 256];
      q[2][j] = epart[2 * nblocks + i + j * 256];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { s[0] += q[0][j]; s[1] += q[1][j]; s[2] += q[2][j]; }
  }
  for (; i < nblocks; i += 256) {
    s[0] += epart[i];
    s[1] += epart[nblocks + i];
    s[2] += epart[2 * nblocks + i];
  }
  sh[0][threadIdx.x] = s[0]; sh[1][threadIdx.x] = s[1]; sh[2][threadIdx.x] = s[2];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
      sh[2][threadIdx.x] += sh[2][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double WKE = 0.5 * sh[0][0];
    sc->Wint += 0.5 * sh[1][0];
    sc->Wext += 0.5 * sh[2][0];
    sc->WKE = WKE;
    sc->Etot = fabs(WKE + sc->Wint - sc->Wext);
    const long long k = sc->step - 1;
    if (ehist && k >= 0 && k < sc->hist_cap) {
      ehist[4 * k + 0] = sc->Wint; ehist[4 * k + 1] = sc->Wext; ehist[4 * k + 2] = WKE; ehist[4 * k + 3] = sc->Etot;
    }
    step_ring_write(sc);  // the step's record is complete once its energies are
  }
}

// ---------------------------------------------------------------------------------------------
// layout conversion at the API boundary (host arrays are AoS xyz in the caller's node numbering,
// GlobalVariables.h; device planes use the internal node order, nref[i] = caller's id or -1 for padding)
__global__ void k_aos_to_soa(const double* __restrict__ aos, double* x, double* y, double* z, const int* __restrict__ nref, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = nref[i];
  if (r >= 0) { x[i] = aos[3 * (size_t)r]; y[i] = aos[3 * (size_t)r + 1]; z[i] = aos[3 * (size_t)r + 2]; }
}
__global__ void k_soa_to_aos(const double* __restrict__ x, const double* __restrict__ y,
                             const double* __restrict__ z, double* aos, const int* __restrict__ nref, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = nref[i];
  if (r >= 0) { aos[3 * (size_t)r] = x[i]; aos[3 * (size_t)r + 1] = y[i]; aos[3 * (size_t)r + 2] = z[i]; }
}
__global__ void k_boundary_to_flags(const int* __restrict__ boundary, uint16_t* flags, const int* __restrict__ nref, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = nref[i];
  if (r < 0) return;
  unsigned f = flags[i] & ~7u;
  f |= (boundary[3 * (size_t)r] ? 1u : 0u) | (boundary[3 * (size_t)r + 1] ? 2u : 0u) | (boundary[3 * (size_t)r + 2] ? 4u : 0u);
  if (f & FTB_FLAG_RIGID) f |= 7u;  // rigid-body nodes stay fully constrained whatever the caller's array says
  flags[i] = (uint16_t)f;
}
__global__ void k_flags_to_boundary(const uint16_t* __restrict__ flags, int* boundary, const int* __restrict__ nref, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = nref[i];
  if (r < 0) return;
  const unsigned f = flags[i];
  boundary[3 * (size_t)r] = f & 1u; boundary[3 * (size_t)r + 1] = (f >> 1) & 1u; boundary[3 * (size_t)r + 2] = (f >> 2) & 1u;
}
__global__ void k_set_bc_kinds(const int* __restrict__ kind, uint16_t* flags, const int* __restrict__ nref, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = nref[i];
  if (r < 0) return;
  unsigned f = flags[i] & ~0x3F0u;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const unsigned k = (unsigned)kind[3 * (size_t)r + c] & 3u;
    f |= k << (4 + 2 * c);
  }
  flags[i] = (uint16_t)f;
}
// ApplyBoundaryConditions at a given Time (Benchmarking-Parallel.cpp:184-244)
__global__ void k_apply_bc(double* ux, double* uy, double* uz, double* vx, double* vy, double* vz, double* ax,
                           double* ay, double* az, uint16_t* flags, const DevScalars* sc, double Time, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned f = flags[i];
  double* U[3] = {ux, uy, uz};
  double* V[3] = {vx, vy, vz};
  double* Ac[3] = {ax, ay, az};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const unsigned k = (f >> (4 + 2 * c)) & 3u;
    if (k) {
      const double r = sc->bc_rate[k];
      f |= 1u << c;
      U[c][i] = Time * r;
      V[c][i] = r;
      Ac[c][i] = 0.0;
    }
  }
  flags[i] = (uint16_t)f;
}
// element is skipped by StableTimeStep when every node is constrained in x, y and z (:13-19)
__global__ void k_eflag(const int* __restrict__ conn, const uint16_t* __restrict__ flags, uint8_t* eflag, int nE) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nE) return;
  bool rigid = true;
#pragma unroll
  for (int k = 0; k < 8; ++k) rigid = rigid && ((flags[conn[(size_t)k * nE + e]] & 7u) == 7u);
  eflag[e] = rigid ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// legacy-path node kernels
// fi = CSR gather (+ optional halo), f_net = fe - fi  (GetForce_3D.cpp:39-51)
__global__ void k_gather_force(const NodeArgs A, double* fnx, double* fny, double* fnz) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= A.nN) return;
  double f[3] = {0.0, 0.0, 0.0};
  const size_t E = (size_t)A.nE;
  for (int j = A.node_off[n]; j < A.node_off[n + 1]; ++j) {
    const int ent = A.node_ent[j];
    const size_t e = (size_t)(ent >> 3);
    const int s = ent & 7;
#pragma unroll
    for (int c = 0; c < 3; ++c) f[c] += A.felem[FTB_FIDX(3 * s + c, e)];
  }
  if (A.halo_node_idx && A.halo_recv) {
    const int h = A.halo_node_idx[n];
    if (h >= 0)
      for (int j = A.halo_off[h]; j < A.halo_off[h + 1]; ++j) {
        const int slot = A.halo_slot[j];
#pragma unroll
        for (int c = 0; c < 3; ++c) f[c] += A.halo_recv[3 * (size_t)slot + c];
      }
  }
  double* fn[3] = {fnx, fny, fnz};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    A.fi[c][n] = f[c];
    fn[c][n] = (A.fe[c] ? A.fe[c][n] : 0.0) - f[c];
  }
}
__global__ void k_accel(const double* fnx, const double* fny, const double* fnz, const double* m,
                        const uint16_t* flags, double* ax, double* ay, double* az, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned f = flags[i];
  const double mm = m[i];
  if (!(f & 1u)) ax[i] = fnx[i] / mm;
  if (!(f & 2u)) ay[i] = fny[i] / mm;
  if (!(f & 4u)) az[i] = fnz[i] / mm;
}
// masked merge on the way out: host accelerations keep their value on boundary dofs
__global__ void k_merge_free(const double* __restrict__ src_aos, double* dst_aos, const int* __restrict__ boundary, int ndof) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ndof && !boundary[i]) dst_aos[i] = src_aos[i];
}

// K7: lumped mass.  Element part writes me[8 planes]; the gather sums in ascending element order
// (Mass3D.cpp:127-157).  detmin: smallest reference Jacobian determinant (as ordered bits).
__global__ void k_mass_elem(const ElemArgs A, double* me, unsigned long long* detmin_bits, int* nonpos) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= A.nE) return;
  const size_t E = (size_t)A.nE;
  double X[8][3];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int nd = A.conn[(size_t)k * E + e];
#pragma unroll
    for (int c = 0; c < 3; ++c) X[k][c] = A.X[c][nd];
  }
  const double rho = A.mp[(size_t)A.pid[e] * FTB_MP_STRIDE + MP_RHO];
  double m8[8];
  double dmin;
  if (A.etype && A.etype[e]) {
    double Xt[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) Xt[k][c] = X[k][c];
    dmin = tet4_lumped_mass(Xt, rho, m8);
#pragma unroll
    for (int k = 4; k < 8; ++k) m8[k] = 0.0;
  } else {
    dmin = hex8_lumped_mass(X, rho, m8);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) me[(size_t)k * E + e] = m8[k];
  if (!(dmin > 0.0)) atomicAdd(nonpos, 1);
  else atomicMin(detmin_bits, (unsigned long long)__double_as_longlong(dmin));
}
__global__ void k_mass_gather(const double* me, const int* node_off, const int* node_ent, double* m, int nN, int nE) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nN) return;
  double s = 0.0;
  for (int j = node_off[n]; j < node_off[n + 1]; ++j) {
    const int ent = node_ent[j];
    s += me[(size_t)(ent & 7) * nE + (ent >> 3)];
  }
  m[n] = s;
}

// halo pack / add (GetForce_3D.cpp:56-61,92-97; Mass3D.cpp:79-85,116-121)
__global__ void k_halo_pack(const double* fx, const double* fy, const double* fz, const int* sendNodeIndex,
                            double* send, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int n = sendNodeIndex[i];
  send[3 * (size_t)i] = fx[n];
  send[3 * (size_t)i + 1] = fy[n];
  send[3 * (size_t)i + 2] = fz[n];
}
// one thread per shared node; its slots are visited in ascending neighbour order
__global__ void k_halo_add(double* fx, double* fy, double* fz, const int* halo_nodes, const int* halo_off,
                           const int* halo_slot, const double* recv, int nshared) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= nshared) return;
  const int n = halo_nodes[h];
  double f0 = fx[n], f1 = fy[n], f2 = fz[n];
  for (int j = halo_off[h]; j < halo_off[h + 1]; ++j) {
    const int s = halo_slot[j];
    f0 += recv[3 * (size_t)s];
    f1 += recv[3 * (size_t)s + 1];
    f2 += recv[3 * (size_t)s + 2];
  }
  fx[n] = f0; fy[n] = f1; fz[n] = f2;
}
// partial internal force of the shared nodes only (boundary elements have been computed)
__global__ void k_gather_shared(const double* felem, const int* node_off, const int* node_ent, const int* halo_nodes,
                                double* fx, double* fy, double* fz, int nshared, int nE) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= nshared) return;
  const int n = halo_nodes[h];
  double f[3] = {0, 0, 0};
  for (int j = node_off[n]; j < node_off[n + 1]; ++j) {
    const int ent = node_ent[j];
    const size_t e = (size_t)(ent >> 3);
    const int s = ent & 7;
#pragma unroll
    for (int c = 0; c < 3; ++c) f[c] += felem[FTB_FIDX(3 * s + c, e)];
  }
  fx[n] = f[0]; fy[n] = f[1]; fz[n] = f[2];
}

// Gauss-point outputs in the reference's layouts (lazy; never on the hot path)
struct OutSink {
  static constexpr bool enabled = true;
  static constexpr bool want_S = true;
  double *F, *detF, *pk2;
  size_t e;
  __device__ __forceinline__ void put(int gp, const double Fm[3][3], double J, const double Sv[6]) const {
    if (F) {
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) F[72 * e + 9 * gp + 3 * j + i] = Fm[i][j];  // column-major, fptr[e]=72e
    }
    if (detF) detF[8 * e + gp] = J;
    if (pk2) {
#pragma unroll
      for (int i = 0; i < 6; ++i) pk2[48 * e + 6 * gp + i] = Sv[i];
    }
  }
};
// ref_of[e]: reference (caller) element id of internal element e
__global__ void k_gp_outputs(const ElemArgs A, const int* ref_of, const int* gpoff, double* F, double* detF, double* pk2, double* Eavg) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= A.nE) return;
  const size_t E = (size_t)A.nE;
  double X[8][3], U[8][3];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int nd = A.conn[(size_t)k * E + e];
#pragma unroll
    for (int c = 0; c < 3; ++c) { X[k][c] = A.X[c][nd]; U[k][c] = A.u[c][nd]; }
  }
  const double* mp = A.mp + (size_t)A.pid[e] * FTB_MP_STRIDE;
  const int mat = (int)mp[MP_MATID];
  const size_t re = (size_t)ref_of[e];
  if (A.etype && A.etype[e]) {  // C3D4: one Gauss point, packed layouts (fptr = 9 gpoff, pk2ptr = 6 gpoff)
    double Xt[4][3], Ut[4][3], Fl1[9], ft[4][3], dd;
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) { Xt[k][c] = X[k][c]; Ut[k][c] = U[k][c]; }
    const size_t g0 = (size_t)gpoff[re];
    OutSink st{Eavg ? Fl1 : (F ? F + 9 * g0 : nullptr), detF ? detF + g0 : nullptr, pk2 ? pk2 + 6 * g0 : nullptr, 0};
    if (Eavg && !F) { st.detF = nullptr; st.pk2 = nullptr; }
    DevHist ht{A.hist, E, (size_t)e};
    tet4_element<-1, false>(Xt, Ut, mat, mp, false, ht, st, ft, &dd);
    if (Eavg) {
      double Em[9];
      for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) {
          double sum = 0.0;
          for (int l = 0; l < 3; ++l) sum += Fl1[l + 3 * i] * Fl1[l + 3 * j];
          Em[i + 3 * j] = 0.5 * sum;
        }
      Em[0] -= 0.5; Em[4] -= 0.5; Em[8] -= 0.5;
      for (int i = 0; i < 9; ++i) Eavg[9 * re + i] = Em[i];
      if (F || detF || pk2) {
        OutSink s2{F ? F + 9 * g0 : nullptr, detF ? detF + g0 : nullptr, pk2 ? pk2 + 6 * g0 : nullptr, 0};
        tet4_element<-1, false>(Xt, Ut, mat, mp, false, ht, s2, ft, &dd);
      }
    }
    return;
  }
  const size_t ge = gpoff ? (size_t)gpoff[re] : 8 * re;  // first Gauss point of the element in the packed arrays
  double Fl[72];
  OutSink sink{Eavg ? Fl : (F ? F + 9 * ge : nullptr), detF ? detF + ge : nullptr, pk2 ? pk2 + 6 * ge : nullptr, 0};
  if (Eavg) { sink.detF = nullptr; sink.pk2 = nullptr; }
  DevHist h{A.hist, E, (size_t)e};
  double fe[8][3], d;
  LocalScratch S;
  hex8_element<-1, false>(X, U, mat, mp, false, h, sink, S, fe, &d);
  if (Eavg) {
    // CalculateStrain.cpp:77-97: E = sum_gp (0.5/8) F^T F - 0.5 I
    double Em[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) Em[i] = 0.0;
    for (int gp = 0; gp < 8; ++gp)
      for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) {
          double s = 0.0;
          for (int l = 0; l < 3; ++l) s += Fl[9 * gp + l + 3 * i] * Fl[9 * gp + l + 3 * j];
          Em[i + 3 * j] += (0.5 / 8.0) * s;
        }
    Em[0] -= 0.5; Em[4] -= 0.5; Em[8] -= 0.5;
    for (int i = 0; i < 9; ++i) Eavg[9 * re + i] = Em[i];
    // second pass for the other outputs, if requested
    if (F || detF || pk2) {
      OutSink s2{F ? F + 9 * ge : nullptr, detF ? detF + ge : nullptr, pk2 ? pk2 + 6 * ge : nullptr, 0};
      hex8_element<-1, false>(X, U, mat, mp, false, h, s2, S, fe, &d);
    }
  }
}

// K8 for the legacy CheckEnergy call: all operands supplied by the host (AoS, staged on the device)
__global__ void k_energy_legacy(const double* u, const double* up, const double* v, const double* a, const double* ap,
                                const double* fi, const double* fip, const double* fe, const double* fep,
                                const int* boundary, const double* m, const uint16_t* flags, const int* nint,
                                double* epart, int nN) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;  // caller's node id
  double wke = 0, wint = 0, wext = 0;
  if (n < nN && !(flags[nint[n]] & FTB_FLAG_NOTOWNED)) {
    const double mm = m[nint[n]];
    for (int c = 0; c < 3; ++c) {
      const size_t i = 3 * (size_t)n + c;
      const double dd = u[i] - up[i];
      wke += mm * v[i] * v[i];
      if (boundary[i]) wext += dd * (fip[i] + fi[i] + mm * (a[i] + ap[i]));
      wint += dd * (fip[i] + fi[i]);
      wext += dd * ((fep ? fep[i] : 0.0) + (fe ? fe[i] : 0.0));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    wke += __shfl_down_sync(0xffffffffu, wke, o);
    wint += __shfl_down_sync(0xffffffffu, wint, o);
    wext += __shfl_down_sync(0xffffffffu, wext, o);
  }
  __shared__ double sw[3][NODE_BLOCK / 32];
  if ((threadIdx.x & 31) == 0) { sw[0][threadIdx.x >> 5] = wke; sw[1][threadIdx.x >> 5] = wint; sw[2][threadIdx.x >> 5] = wext; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s0 = 0, s1 = 0, s2 = 0;
    for (int w = 0; w < NODE_BLOCK / 32; ++w) { s0 += sw[0][w]; s1 += sw[1][w]; s2 += sw[2][w]; }
    epart[blockIdx.x] = s0; epart[gridDim.x + blockIdx.x] = s1; epart[2 * gridDim.x + blockIdx.x] = s2;
  }
}
__global__ void k_sum3(const double* epart, int nblocks, double* out3) {
  __shared__ double sh[3][256];
  double s[3] = {0, 0, 0};
  for (int i = threadIdx.x; i < nblocks; i += 256) { s[0] += epart[i]; s[1] += epart[nblocks + i]; s[2] += epart[2 * nblocks + i]; }
  sh[0][threadIdx.x] = s[0]; sh[1][threadIdx.x] = s[1]; sh[2][threadIdx.x] = s[2];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; sh[2][threadIdx.x] += sh[2][threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out3[0] = 0.5 * sh[0][0]; out3[1] = 0.5 * sh[1][0]; out3[2] = 0.5 * sh[2][0]; }
}

// =============================================================================================
// Peer-memory transport of the shared-node exchange (NVLink / NVSwitch, no NCCL on the data path).
// Every rank owns one device "window" (cudaMalloc, exported with cudaIpcGetMemHandle):
//     hflag[P2P_MAXNB]   per neighbour index: sequence number of the last step whose partials have arrived
//     dflag[P2P_MAXP]    per source rank:    sequence number of the last step whose dt has arrived
//     dtslot[2][P2P_MAXP] local stable dt of every rank, double buffered by step parity
//     recv[2][3*H]       receive window in the reference's recvNodeDisplacement layout, double buffered
// k_p2p_pack sums the element forces of each shared node and STORES the partial straight into the
// neighbours' windows (the neighbour's slice offset is known from the symmetric send lists), then the last
// block publishes the sequence number with a system-scope fence.  k_adv_p2p writes this rank's dt into
// every rank's window, waits for all dt and all neighbours' partials of this step (bounded spin), takes the
// MIN and does the scalar update of k_adv.  Two buffers suffice: a rank cannot run two steps ahead of a
// neighbour because it needs that neighbour's dt to finish each step.  Everything is graph-capturable.
constexpr int P2P_MAXNB = 64;
constexpr int P2P_MAXP = 64;
constexpr unsigned long long P2P_DT_EMPTY = 0x7FF8000000000001ULL;  // a NaN: dt_to_bits never produces one
struct P2PHeader {
  unsigned long long hflag[P2P_MAXNB];
  unsigned long long dflag[P2P_MAXP];
  // dt of every rank for step seq in dtslot[seq & 3]: the value is its own arrival flag (a slot holds P2P_DT_EMPTY until
  // the rank's store lands; the reader re-arms slot (seq + 2) & 3 after consuming slot seq & 3 -- a peer can only write
  // that slot after it has seen this rank's dt of step seq + 1, which is published after the re-arm)
  double dtslot[4][P2P_MAXP];
  unsigned long long iflag[P2P_MAXP];  // per source rank: (step sequence * 8 + radix pass + 1) of its last injury histogram
};
struct P2PArgs {
  char* self;                    // this rank's window
  char* peer_nb[P2P_MAXNB];      // windows of the neighbours (by neighbour index)
  int peer_slot_off[P2P_MAXNB];  // first slot of this rank's slice in the neighbour's receive window
  int peer_my_index[P2P_MAXNB];  // this rank's neighbour index in the neighbour's list (its hflag entry)
  int peer_H[P2P_MAXNB];         // slots of the neighbour's receive window (offset of its second buffer)
  int nb_cum[P2P_MAXNB + 1];     // sendNeighbourCountCum
  char* peer_rank[P2P_MAXP];     // windows of all ranks (dt exchange)
  int n_nb, n_ranks, rank, H;
  unsigned long long* seq;       // monotone step sequence (device), never reset
  unsigned* blocks_done;
};
// window layout: header | injury histograms [2][n_ranks][P2P_IHIST counters] | receive buffers [2][3 H]; the first two
// parts have the same size on every rank, so a peer's histogram area is found without knowing its halo count
constexpr int P2P_IHIST = 2 * 2048;  // = 2 * INJ_BINS (static_assert below)
__host__ __device__ __forceinline__ size_t p2p_ihist_off() { return (sizeof(P2PHeader) + 255) & ~(size_t)255; }
__host__ __device__ __forceinline__ size_t p2p_recv_off(int n_ranks) {
  return p2p_ihist_off() + (size_t)2 * (size_t)n_ranks * P2P_IHIST * sizeof(unsigned);
}
__host__ __device__ __forceinline__ double* p2p_recv(char* win, int H, int buf, int n_ranks) {
  return reinterpret_cast<double*>(win + p2p_recv_off(n_ranks)) + (size_t)buf * 3 * (size_t)H;
}

// ---------------------------------------------------------------------------------------------
// The exchange of the partitioned step FUSED INTO THE ELEMENT KERNELS (default; FTB200_P2P_FUSED=0 restores the separate
// k_p2p_pack launch and the publish phase of k_adv_p2p).  Why: a kernel launched behind the interior elements cannot get
// a block onto an SM before the interior grid has handed out all of its blocks -- k_elem_affine_cj allocates every
// register of an SM, and neither stream priorities nor explicit launch priorities change the order in which the block
// scheduler serves the two grids (profiles/r02_p2p_trace_2gpu_*.txt: the pack kernel, launched 10 us into the step,
// started after 103 us).  So nothing is launched: the boundary elements themselves send.
//   * shared-node partial sums: every boundary element, after its force stores, counts itself in at each of its shared
//     nodes; the LAST element to arrive at a node sums the node's local contributions in the node map's order (the same
//     order, so the same bits, as k_p2p_pack / the reference's scatter) and stores the sum into the neighbours' receive
//     windows over NVLink; the thread that packs the last shared node raises this rank's flag in every neighbour's
//     header.  The partial sums leave ~10 us into the step, under the interior elements.
// Only the launch of the boundary elements carries the epilogue (ElemArgs::pk); the dt MIN stays in k_adv_p2p: having the
// last element block of the step publish it was measured -- the fence + counter of every block costs the element phase
// 10 us at 100^3 (a block that has just streamed out 12 KB of forces waits for them instead of retiring).
// No spinning inside the element kernels.
struct PackArgs {
  P2PArgs P;
  const int* halo_node_idx;  // node -> shared-node index, -1 for the others
  const int* halo_off;       // shared node -> its send positions (CSR), ascending neighbour
  const int* halo_slot;      // send position i = position in sendNodeIndex = position in the neighbour's slice of its window
  const int* node_off;       // node -> (element, slot) map
  const int* node_ent;
  unsigned* node_ctr;        // [n_shared] arrivals, back to zero when the node is packed
  unsigned* packed;          // shared nodes packed this step
  int n_shared, nEb;
};
__device__ __noinline__ void elem_p2p_epilogue(const PackArgs* pk, const int* conn, const int nE, const double* felem, const int e,
                                               const bool valid) {
  const PackArgs& K = *pk;
  if (!(valid && e < K.nEb)) return;  // only elements touching shared nodes (they come first in the internal order)
  const P2PArgs& P = K.P;
  __threadfence();  // this element's force stores, device-wide, before it counts itself in
  const size_t E = (size_t)nE;
  const unsigned long long seq = *P.seq;
  const int buf = (int)(seq & 1ULL);
  unsigned mine = 0;  // bit k: this thread is the last contributor of its node k
  int nd[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) nd[k] = conn[(size_t)k * E + e];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int h = K.halo_node_idx[nd[k]];
    if (h < 0) continue;
    const unsigned deg = (unsigned)(K.node_off[nd[k] + 1] - K.node_off[nd[k]]);
    if (atomicAdd(&K.node_ctr[h], 1u) == deg - 1) {
      K.node_ctr[h] = 0;  // every contributor has arrived: nobody touches the counter again in this step
      mine |= 1u << k;
    }
  }
  if (!mine) return;
  __threadfence();
  unsigned count = 0;
  for (int k = 0; k < 8; ++k) {
    if (!((mine >> k) & 1u)) continue;
    const int n = nd[k];
    const int h = K.halo_node_idx[n];
    double f[3] = {0.0, 0.0, 0.0};
    for (int j = K.node_off[n], j1 = K.node_off[n + 1]; j < j1; ++j) {
      const int ent = K.node_ent[j];
      const size_t el = (size_t)(ent >> 3);
      const int sl = ent & 7;
#pragma unroll
      for (int c = 0; c < 3; ++c) f[c] += __ldcg(felem + FTB_FIDX(3 * sl + c, el));  // L2: written by other SMs in this launch
    }
    for (int q = K.halo_off[h]; q < K.halo_off[h + 1]; ++q) {
      const int i = K.halo_slot[q];
      int nb = 0;
      while (i >= P.nb_cum[nb + 1]) ++nb;
      double* dst = p2p_recv(P.peer_nb[nb], P.peer_H[nb], buf, P.n_ranks) + 3 * (size_t)(P.peer_slot_off[nb] + (i - P.nb_cum[nb]));
      dst[0] = f[0]; dst[1] = f[1]; dst[2] = f[2];  // peer store over NVLink
    }
    ++count;
  }
  __threadfence_system();  // one system-scope fence per thread, after all of its peer stores
  if (atomicAdd(K.packed, count) + count == (unsigned)K.n_shared) {
    *K.packed = 0;
    __threadfence_system();
    for (int nb = 0; nb < P.n_nb; ++nb) {
      volatile unsigned long long* fl = &reinterpret_cast<P2PHeader*>(P.peer_nb[nb])->hflag[P.peer_my_index[nb]];
      *fl = seq + 1;
    }
  }
}

__global__ void k_p2p_pack(const P2PArgs P, const double* felem, const int* node_off, const int* node_ent,
                           const int* sendNodeIndex, const DevScalars* sc, int nE) {
  if (sc->last | sc->done) return;
  const unsigned long long seq = *P.seq;
  const int buf = (int)(seq & 1ULL);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P.H) {
    int nb = 0;
    while (i >= P.nb_cum[nb + 1]) ++nb;  // few neighbours: linear search in the cumulative counts
    const int n = sendNodeIndex[i];
    double f[3] = {0.0, 0.0, 0.0};
    for (int j = node_off[n]; j < node_off[n + 1]; ++j) {  // every element of a shared node is a boundary element
      const int ent = node_ent[j];
      const size_t e = (size_t)(ent >> 3);
      const int sl = ent & 7;
#pragma unroll
      for (int c = 0; c < 3; ++c) f[c] += felem[FTB_FIDX(3 * sl + c, e)];
    }
    double* dst = p2p_recv(P.peer_nb[nb], P.peer_H[nb], buf, P.n_ranks) + 3 * (size_t)(P.peer_slot_off[nb] + (i - P.nb_cum[nb]));
    dst[0] = f[0]; dst[1] = f[1]; dst[2] = f[2];  // peer store over NVLink
  }
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    s_last = (atomicAdd(P.blocks_done, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    __threadfence_system();
    for (int nb = threadIdx.x; nb < P.n_nb; nb += blockDim.x) {
      volatile unsigned long long* fl = &reinterpret_cast<P2PHeader*>(P.peer_nb[nb])->hflag[P.peer_my_index[nb]];
      *fl = seq + 1;
    }
    if (threadIdx.x == 0) *P.blocks_done = 0;
  }
}

// Diagnostic trace of the partitioned step (FTB200_P2P_TRACE=<file prefix>; off by default, no kernel of the default path
// changes): TRACE_SLOTS time stamps (%globaltimer, ns) per step, written by one-thread marker kernels between the
// launches of the step and by k_adv_p2p around its publish / wait phases; dumped per rank when the context is destroyed.
//   0 step start | 1 boundary elements done | 2 partial sums packed and flagged | 3 k_adv_p2p starts (interior joined)
//   4 dt published to every rank | 5 every rank's dt and every neighbour's flag seen | 6 node kernel done | 7 interior done
constexpr int TRACE_SLOTS = 8, TRACE_STEPS = 2048;
__global__ void k_stamp(unsigned long long* tr, const DevScalars* sc, int slot, int step_offset) {
  const long long st = (long long)sc->step + step_offset;
  if (st >= 0 && st < TRACE_STEPS) tr[st * TRACE_SLOTS + slot] = now_ns();
}

// dt exchange + waits + the scalar update of k_adv<false>
__global__ void k_adv_p2p(const P2PArgs P, DevScalars* sc, double* mp, int nPID, double* dt_hist, unsigned long long* trace,
                          const int published) {
  __shared__ double s_ndt;
  __shared__ int s_live, s_ok;
  unsigned long long* tr = nullptr;
  if (trace && threadIdx.x == 0 && sc->step < TRACE_STEPS) { tr = trace + (size_t)sc->step * TRACE_SLOTS; tr[3] = now_ns(); }
  if (threadIdx.x == 0) {
    int live = 1;
    if (sc->done) live = 0;
    else if (sc->last) { sc->done = 1; live = 0; }
    sc->active = live;
    s_live = live;
    s_ok = 1;
  }
  __syncthreads();
  if (!s_live) return;
  const unsigned long long seq = *P.seq;
  const int buf = (int)(seq & 1ULL);
  P2PHeader* self = reinterpret_cast<P2PHeader*>(P.self);
  // publish this rank's dt to every rank (its own window included): ONE 8-byte store per rank, no fence and no flag
  const unsigned long long mybits = *(volatile unsigned long long*)&sc->dtmin_bits;
  const int slot = (int)(seq & 3ULL);
  if (!published)
    for (int r = threadIdx.x; r < P.n_ranks; r += blockDim.x)
      *(volatile unsigned long long*)&reinterpret_cast<P2PHeader*>(P.peer_rank[r])->dtslot[slot][P.rank] = mybits;
  if (tr) tr[4] = now_ns();
  // wait for every rank's dt and every neighbour's partials of this step
  const unsigned long long t0 = now_ns();
  for (int r = threadIdx.x; r < P.n_ranks + P.n_nb; r += blockDim.x) {
    volatile unsigned long long* fl = r < P.n_ranks ? (volatile unsigned long long*)&self->dtslot[slot][r] : &self->hflag[r - P.n_ranks];
    while (r < P.n_ranks ? (*fl == P2P_DT_EMPTY) : (*fl < seq + 1)) {
      __nanosleep(100);
      if (now_ns() - t0 > 5000000000ULL) { s_ok = 0; break; }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (tr) tr[5] = now_ns();
  if (threadIdx.x == 0) {
    if (!s_ok) { sc->status |= 64; sc->last = 1; }  // a peer never arrived: stop instead of hanging
    double dtmin = 1e300;
    for (int r = 0; r < P.n_ranks; ++r) {
      const double d = *(volatile double*)&self->dtslot[slot][r];
      if (d < dtmin) dtmin = d;
      *(volatile unsigned long long*)&self->dtslot[(slot + 2) & 3][r] = P2P_DT_EMPTY;  // re-arm (see P2PHeader)
    }
    sc->dtmin_bits = (unsigned long long)__double_as_longlong(dtmin);
    *P.seq = seq + 1;
    s_ndt = adv_step(sc, dt_hist);
  }
  __syncthreads();
  prony_update(mp, nPID, s_ndt, threadIdx.x, blockDim.x);
}


constexpr int NODE_TILE = 128;  // internal node order is padded to whole tiles of this many nodes

// =============================================================================================
// Injury criteria of the brain drivers (examples/ex5/ex5.cpp:1311-1430), device side.  k_elem<..., WITH_INJ> leaves
// the per-element quantities of the step; the kernels below do what the reference's loop does across elements:
// running extrema with their element and time, the 95th-percentile values (math.cpp:160-199: the order statistic
// (int)(0.95 n) - 1, found here by a 6-pass (11-bit digits) radix select on order-preserving keys -- a selection, so bit-exact),
// and the element lists of the percentile maxima.
constexpr int INJ_BINS = 2048;    // 11-bit digits: 6 passes over the 64-bit keys (the top pass has 9 bits)
constexpr int INJ_PASSES = 6;
static_assert(P2P_IHIST == 2 * INJ_BINS, "peer-memory window layout");
struct InjState {
  double scal[12];  // maxStrain, maxT, minStrain, minT, maxShear, maxShearT, maxPSxSR, maxTimePSxSR, MPS95, t, MPSxSR95, t
  int elems[4];     // reference element ids of the four extrema (ex5.cpp:63,74)
  int upd[2];       // this step raised MPS-95 / MPSxSR-95
  unsigned red_done, sel_done[2];
  unsigned kth0, kth[2];
  unsigned long long prefix[2];
  unsigned hist[2][INJ_BINS];
  int nIncluded, pad;
};
constexpr int INJ_BLOCKS = 592;  // 4 per SM
constexpr int INJ_THREADS = 256;
constexpr int INJ_ITEMS = 8;      // elements per thread and loop trip, loaded before they are used

struct InjCand { double v; int id; };
// a beats b: strictly larger value, or the same value at a lower reference element id (the reference's loop keeps
// the first element that attains the maximum, ex5.cpp:1318-1332)
__device__ __forceinline__ InjCand inj_best(InjCand a, InjCand b) {
  return (b.v > a.v || (b.v == a.v && b.id < a.id)) ? b : a;
}
__device__ __forceinline__ InjCand inj_shfl(InjCand a, int o) {
  InjCand b;
  b.v = __shfl_xor_sync(0xffffffffu, a.v, o);
  b.id = __shfl_xor_sync(0xffffffffu, a.id, o);
  return b;
}

// part: [4][INJ_BLOCKS] values, parti: [4][INJ_BLOCKS] ids.  Metric 1 (minimum strain) is reduced as the maximum of -smin.
__global__ void __launch_bounds__(INJ_THREADS) k_injury_reduce(const ElemArgs A, const int* ref_of, InjState* st, double* part,
                                                                int* parti) {
  const DevScalars* sc = A.sc;
  if (!sc->active) return;
  InjCand c[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) { c[m].v = -1.0; c[m].id = 0x7fffffff; }
  for (int base = blockIdx.x * INJ_THREADS * 4 + threadIdx.x; base < A.nE; base += gridDim.x * INJ_THREADS * 4) {
    double q[4][4];
    int id[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // loads first
      const int e = base + j * INJ_THREADS;
      const bool on = e < A.nE && A.inj_incl[e];
      id[j] = on ? ref_of[e] : -1;
      q[j][0] = on ? A.inj_ps[e] : 0.0; q[j][1] = on ? -A.inj_smin[e] : 0.0;
      q[j][2] = on ? A.inj_shear[e] : 0.0; q[j][3] = on ? A.inj_psxsr[e] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (id[j] < 0) continue;
#pragma unroll
      for (int m = 0; m < 4; ++m) c[m] = inj_best(c[m], InjCand{q[j][m], id[j]});
    }
  }
  __shared__ double sv[4][INJ_THREADS / 32];
  __shared__ int si[4][INJ_THREADS / 32];
  __shared__ int s_last;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c[m] = inj_best(c[m], inj_shfl(c[m], o));
    if ((threadIdx.x & 31) == 0) { sv[m][threadIdx.x >> 5] = c[m].v; si[m][threadIdx.x >> 5] = c[m].id; }
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    const int m = threadIdx.x;
    InjCand b{sv[m][0], si[m][0]};
    for (int w = 1; w < INJ_THREADS / 32; ++w) b = inj_best(b, InjCand{sv[m][w], si[m][w]});
    part[m * INJ_BLOCKS + blockIdx.x] = b.v;
    parti[m * INJ_BLOCKS + blockIdx.x] = b.id;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&st->red_done, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // final stage: the block partials, again by the order-independent "best" rule -> deterministic
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    InjCand b{-1.0, 0x7fffffff};
    for (int k = threadIdx.x; k < (int)gridDim.x; k += INJ_THREADS)
      b = inj_best(b, InjCand{__ldcg(part + m * INJ_BLOCKS + k), __ldcg(parti + m * INJ_BLOCKS + k)});
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b = inj_best(b, inj_shfl(b, o));
    if ((threadIdx.x & 31) == 0) { sv[m][threadIdx.x >> 5] = b.v; si[m][threadIdx.x >> 5] = b.id; }
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    const int m = threadIdx.x;
    InjCand b{sv[m][0], si[m][0]};
    for (int w = 1; w < INJ_THREADS / 32; ++w) b = inj_best(b, InjCand{sv[m][w], si[m][w]});
    // running extremum with its element and time: `if (maxStrain < current)`, ex5.cpp:1318-1332,1350-1354
    const double cur = (m == 1) ? -st->scal[2] : st->scal[2 * m];
    if (b.id != 0x7fffffff && cur < b.v) {
      st->scal[2 * m] = (m == 1) ? -b.v : b.v;
      st->scal[2 * m + 1] = sc->Time;
      st->elems[m] = b.id;
    }
    if (m == 0) st->red_done = 0;
  }
}

__device__ __forceinline__ unsigned long long inj_key(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);  // ascending keys == ascending doubles
}
__device__ __forceinline__ double inj_unkey(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFULL) : ~k;
  return __longlong_as_double((long long)b);
}

// one digit (most significant first) of the two selections: blockIdx.y = 0 MPS, 1 MPSxSR.  Pass p looks at bits
// [shift, shift + width) with shift = 55, 44, 33, 22, 11, 0.  The strains of a step share their exponent, so the
// leading digits are nearly constant: the histogram votes are aggregated per warp (__match_any_sync) before they
// reach shared memory -- one atomic per distinct digit per warp instead of 32 colliding ones.
__global__ void __launch_bounds__(INJ_THREADS) k_injury_select(const ElemArgs A, InjState* st, const int pass, double* hist95,
                                                                double* histx95, const int hist_only) {
  const DevScalars* sc = A.sc;
  if (!sc->active) return;
  const int arr = blockIdx.y;
  const double* data = arr ? A.inj_psxsr : A.inj_ps;
  __shared__ unsigned h[INJ_BINS];
  __shared__ int s_last;
  for (int i = threadIdx.x; i < INJ_BINS; i += INJ_THREADS) h[i] = 0;
  __syncthreads();
  const int shift = 55 - 11 * pass;
  const unsigned long long prefix = st->prefix[arr];
  const int lane = threadIdx.x & 31;
  const int nLoop = (A.nE + gridDim.x * INJ_THREADS * INJ_ITEMS - 1) / (gridDim.x * INJ_THREADS * INJ_ITEMS);
  for (int it = 0; it < nLoop; ++it) {  // whole warps stay in the loop: __match_any_sync needs a full mask
    const int base = (it * gridDim.x + blockIdx.x) * INJ_THREADS * INJ_ITEMS + threadIdx.x;
    double v[INJ_ITEMS];
    uint8_t in[INJ_ITEMS];
#pragma unroll
    for (int j = 0; j < INJ_ITEMS; ++j) {  // all loads in flight before the first vote
      const int e = base + j * INJ_THREADS;
      in[j] = e < A.nE ? A.inj_incl[e] : (uint8_t)0;
      v[j] = e < A.nE ? data[e] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < INJ_ITEMS; ++j) {
      unsigned d = 0xFFFFFFFFu;  // not a candidate
      if (in[j]) {
        const unsigned long long k = inj_key(v[j]);
        if (pass == 0 || ((k ^ prefix) >> (shift + 11)) == 0) d = (unsigned)(k >> shift) & (INJ_BINS - 1);
      }
      const unsigned peers = __match_any_sync(0xffffffffu, d);
      if (d != 0xFFFFFFFFu && lane == __ffs(peers) - 1) atomicAdd(&h[d], (unsigned)__popc(peers));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < INJ_BINS; i += INJ_THREADS)
    if (h[i]) atomicAdd(&st->hist[arr][i], h[i]);
  if (hist_only) return;  // several partitions: the histograms are summed across ranks first, then k_injury_pick
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&st->sel_done[arr], 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int i = threadIdx.x; i < INJ_BINS; i += INJ_THREADS) {
    h[i] = __ldcg(&st->hist[arr][i]);
    st->hist[arr][i] = 0;
  }
  __syncthreads();
  // the digit whose bucket holds rank k: 8 bins per thread, exclusive scan of the per-thread sums, then a local walk
  constexpr int PER = INJ_BINS / INJ_THREADS;
  __shared__ unsigned tsum[INJ_THREADS];
  unsigned mine = 0;
#pragma unroll
  for (int j = 0; j < PER; ++j) mine += h[threadIdx.x * PER + j];
  tsum[threadIdx.x] = mine;
  __shared__ unsigned s_k0;
  if (threadIdx.x == 0) s_k0 = pass == 0 ? st->kth0 : st->kth[arr];  // read once, before the owner rewrites it
  __syncthreads();
  for (int o = 1; o < INJ_THREADS; o <<= 1) {  // inclusive scan of the per-thread sums
    const unsigned add = threadIdx.x >= o ? tsum[threadIdx.x - o] : 0u;
    __syncthreads();
    tsum[threadIdx.x] += add;
    __syncthreads();
  }
  const unsigned k0 = s_k0;
  const unsigned excl = tsum[threadIdx.x] - mine;
  // exactly one thread owns rank k0 (the last one also catches an inconsistent count)
  const bool owner = (k0 >= excl && k0 < excl + mine) || (threadIdx.x == INJ_THREADS - 1 && k0 >= tsum[INJ_THREADS - 1]);
  if (owner) {
    unsigned k = k0 - excl;
    unsigned digit = threadIdx.x * PER;
    for (int j = 0; j < PER - 1 && k >= h[digit]; ++j) { k -= h[digit]; ++digit; }
    const unsigned long long np = (pass == 0 ? 0ULL : prefix) | ((unsigned long long)digit << shift);
    st->prefix[arr] = np;
    st->kth[arr] = k;
    st->sel_done[arr] = 0;
    if (pass == INJ_PASSES - 1) {  // ex5.cpp:1372-1377 / :1402-1406
      const double v = inj_unkey(np);
      const long long i = sc->step - 1;
      double* hh = arr ? histx95 : hist95;
      if (hh && i >= 0 && i < sc->hist_cap) hh[i] = v;
      if (v > st->scal[8 + 2 * arr]) {
        st->scal[8 + 2 * arr] = v;
        st->scal[9 + 2 * arr] = sc->Time;
        st->upd[arr] = 1;
      } else {
        st->upd[arr] = 0;
      }
    }
  }
}

// Bucket search of one radix pass on histograms that were summed across the partitions (k_injury_select with
// hist_only, then an all-reduce of st->hist): grid (1, 2), one block per array.  Every rank computes the same digit.
__global__ void __launch_bounds__(INJ_THREADS) k_injury_pick(const DevScalars* sc, InjState* st, const int pass, double* hist95,
                                                              double* histx95) {
  if (!sc->active) return;
  const int arr = blockIdx.y;
  __shared__ unsigned h[INJ_BINS];
  __shared__ unsigned tsum[INJ_THREADS];
  __shared__ unsigned s_k0;
  for (int i = threadIdx.x; i < INJ_BINS; i += INJ_THREADS) {
    h[i] = st->hist[arr][i];
    st->hist[arr][i] = 0;
  }
  __syncthreads();
  constexpr int PER = INJ_BINS / INJ_THREADS;
  unsigned mine = 0;
#pragma unroll
  for (int j = 0; j < PER; ++j) mine += h[threadIdx.x * PER + j];
  tsum[threadIdx.x] = mine;
  if (threadIdx.x == 0) s_k0 = pass == 0 ? st->kth0 : st->kth[arr];
  __syncthreads();
  for (int o = 1; o < INJ_THREADS; o <<= 1) {
    const unsigned add = threadIdx.x >= o ? tsum[threadIdx.x - o] : 0u;
    __syncthreads();
    tsum[threadIdx.x] += add;
    __syncthreads();
  }
  const unsigned k0 = s_k0;
  const unsigned excl = tsum[threadIdx.x] - mine;
  const bool owner = (k0 >= excl && k0 < excl + mine) || (threadIdx.x == INJ_THREADS - 1 && k0 >= tsum[INJ_THREADS - 1]);
  if (!owner) return;
  const int shift = 55 - 11 * pass;
  unsigned k = k0 - excl;
  unsigned digit = threadIdx.x * PER;
  for (int j = 0; j < PER - 1 && k >= h[digit]; ++j) { k -= h[digit]; ++digit; }
  const unsigned long long np = (pass == 0 ? 0ULL : st->prefix[arr]) | ((unsigned long long)digit << shift);
  st->prefix[arr] = np;
  st->kth[arr] = k;
  if (pass == INJ_PASSES - 1) {
    const double v = inj_unkey(np);
    const long long i = sc->step - 1;
    double* hh = arr ? histx95 : hist95;
    if (hh && i >= 0 && i < sc->hist_cap) hh[i] = v;
    if (v > st->scal[8 + 2 * arr]) {
      st->scal[8 + 2 * arr] = v;
      st->scal[9 + 2 * arr] = sc->Time;
      st->upd[arr] = 1;
    } else {
      st->upd[arr] = 0;
    }
  }
}

// Sum of the per-rank histograms of one radix pass through the peer-memory windows (the percentile is a GLOBAL order
// statistic, math.cpp:160-199 gathers all ranks): every rank stores its 2 x 2048 counters into its slot of every rank's
// window (buffer = pass parity), raises a flag, waits for all ranks' flags of this (step, pass) and replaces its own
// histogram by the sum in rank order.  Two buffers suffice: a rank reaches pass p + 2 only after every rank has
// announced pass p + 1, i.e. has finished reading pass p.  One block; k_injury_pick follows.
__global__ void __launch_bounds__(INJ_THREADS) k_injury_xchg(const P2PArgs P, const DevScalars* sc, InjState* st, const int pass) {
  if (!sc->active) return;
  __shared__ int s_ok;
  const unsigned long long want = (*P.seq) * 8ULL + (unsigned long long)pass + 1ULL;  // seq was advanced by k_adv_p2p of this step
  const int buf = pass & 1;
  const size_t slot = (size_t)2 * INJ_BINS;
  const unsigned* mine = &st->hist[0][0];
  for (int r = 0; r < P.n_ranks; ++r) {
    unsigned* dst = reinterpret_cast<unsigned*>(P.peer_rank[r] + p2p_ihist_off()) + ((size_t)buf * P.n_ranks + P.rank) * slot;
    for (int i = threadIdx.x; i < 2 * INJ_BINS; i += INJ_THREADS) dst[i] = mine[i];
  }
  if (threadIdx.x == 0) s_ok = 1;
  __threadfence_system();
  __syncthreads();
  for (int r = threadIdx.x; r < P.n_ranks; r += INJ_THREADS) {
    __threadfence_system();  // release by the thread that raises the flag: the block's stores (ordered by the barrier) first
    *(volatile unsigned long long*)&reinterpret_cast<P2PHeader*>(P.peer_rank[r])->iflag[P.rank] = want;
  }
  const unsigned long long t0 = now_ns();
  for (int r = threadIdx.x; r < P.n_ranks; r += INJ_THREADS) {
    volatile unsigned long long* fl = &reinterpret_cast<P2PHeader*>(P.self)->iflag[r];
    while (*fl < want) {
      __nanosleep(100);
      if (now_ns() - t0 > 5000000000ULL) { s_ok = 0; break; }
    }
  }
  __threadfence_system();
  __syncthreads();
  __threadfence_system();  // acquire in every reading thread too (the flags were observed by other threads of the block)
  if (!s_ok) { if (threadIdx.x == 0) atomicOr(&const_cast<DevScalars*>(sc)->status, 64); return; }
  const volatile unsigned* all = reinterpret_cast<const volatile unsigned*>(P.self + p2p_ihist_off()) + (size_t)buf * P.n_ranks * slot;
  for (int i = threadIdx.x; i < 2 * INJ_BINS; i += INJ_THREADS) {
    unsigned sum = 0;
    for (int r = 0; r < P.n_ranks; ++r) sum += all[(size_t)r * slot + i];  // written by the peers: system-scope loads
    (&st->hist[0][0])[i] = sum;
  }
}

// element lists of the percentile maxima, rebuilt in the step that raised them (ex5.cpp:1376-1398, :1406-1428)
__global__ void __launch_bounds__(256) k_injury_lists(const ElemArgs A, const InjState* st) {
  if (!A.sc->active) return;
  const int u0 = st->upd[0], u1 = st->upd[1];
  if (!(u0 | u1)) return;
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= A.nE || !A.inj_incl[e]) return;
  unsigned f = A.inj_flags[e];
  if (u0) f = (f & ~FTB_INJ_LIST95) | (A.inj_ps[e] >= st->scal[8] ? FTB_INJ_LIST95 : 0u);
  if (u1) f = (f & ~FTB_INJ_LISTX95) | (A.inj_psxsr[e] >= st->scal[10] ? FTB_INJ_LISTX95 : 0u);
  A.inj_flags[e] = (uint8_t)f;
}

// CalculateMaximumPrincipalStrain for every element from the current displacements (legacy / on-demand path) and
// the reference-configuration volumes of computePartVolume (Elements.cpp:30-38); outputs in reference element order
__global__ void k_principal(const ElemArgs A, const int* ref_of, double* smax, double* smin, double* shear, double* vol0) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= A.nE) return;
  const size_t E = (size_t)A.nE;
  double X[8][3], U[8][3];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int nd = A.conn[(size_t)k * E + e];
#pragma unroll
    for (int c = 0; c < 3; ++c) { X[k][c] = A.X[c][nd]; U[k][c] = A.u[c][nd]; }
  }
  const size_t re = (size_t)ref_of[e];
  if (A.etype && A.etype[e]) {
    double Xt[4][3], Ut[4][3], ft[4][3], dd;
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) { Xt[k][c] = X[k][c]; Ut[k][c] = U[k][c]; }
    if (vol0) vol0[re] = tet4_volume(Xt);
    if (!smax) return;
    double cs1[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    tet4_element<0, false>(Xt, Ut, 0, A.mp, false, NoHistory(), StrainSink{cs1}, ft, &dd);
    double a1, b1, c1;
    principal_strains(cs1, &a1, &b1, &c1, 1);
    smax[re] = a1; smin[re] = b1; shear[re] = c1;
    return;
  }
  if (vol0) {
    double xm[7][3], n[8], g[7];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int k = 0; k < 8; ++k) n[k] = X[k][c];
      hex_modes(n, g);
#pragma unroll
      for (int m = 0; m < 7; ++m) xm[m][c] = g[m];
    }
    vol0[re] = hex_volume_modes(xm);
  }
  if (!smax) return;
  const double* mp = A.mp + (size_t)A.pid[e] * FTB_MP_STRIDE;
  double cs[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  double fe[8][3], d;
  LocalScratch S;
  hex8_element<0, false>(X, U, 0, mp, false, NoHistory(), StrainSink{cs}, S, fe, &d);  // material 0: kinematics only
  double a, b, c;
  principal_strains(cs, &a, &b, &c);
  smax[re] = a; smin[re] = b; shear[re] = c;
}

// ---------------------------------------------------------------------------------------------
// roofline denominators
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 12345.678) out[0] = s;  // never true; keeps the chains alive
}
__global__ void __launch_bounds__(256) k_copy_peak(const double2* __restrict__ src, double2* __restrict__ dst, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

}  // namespace ftb
