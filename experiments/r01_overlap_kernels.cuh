// experiments/r01_overlap_kernels.cuh -- NOT BUILT.  Three designs that overlap the fp64-bound element phase with the
// HBM-bound node phase (round 1: k_node_ovl, k_elem_pipe/k_node_pipe, k_step).  All passed the parity suite at commit
// 00d5a01 and all measured SLOWER than the serial two-kernel step (0.416-0.649 ms vs 0.264 ms at 100^3),
// so they were taken out of the shipped library in round 2.  Kept as a record of what was tried; the host-side drivers
// that launched them are in git history (femtech_b200/csrc/ftb200_capi.cu at 00d5a01).

// ---------------------------------------------------------------------------------------------
// Overlapped step (single partition, one uniform run of hexahedra): the memory-bound half of the node work runs BESIDE
// the fp64-bound element kernel instead of after it.
//   k_node_ovl  (helper stream, persistent, at most one or two blocks per SM): the FINISH phase of k_node -- deterministic
//               gather of f_int in ascending element id, a = (f_e - f_i)/m, second kick, energy partials -- tile by tile,
//               each tile as soon as the element chunks it depends on are announced (elem_announce).  It needs only the
//               times of the step being integrated (sc->nt_*: k_adv has not yet moved them to sc->t_*), never the new dt.
//   k_node<false, true, ...> (main stream, after k_adv): the START phase -- first kick and drift with the new dt,
//               boundary conditions -- the only node work left on the critical path.
// The element kernel never waits for anything, and k_node_ovl holds a bounded number of blocks (grid <= 2 x SMs), so the
// pair cannot deadlock whatever order the hardware schedules them in.  Same arithmetic, same summation order and the same
// 128-node energy partials as the serial step: the results are bit-identical (tests/test_gpu_overlap.py).
__device__ __forceinline__ unsigned long long ovl_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
struct OvlArgs {
  NodeArgs N;
  const unsigned* ctr;       // finished warps per element chunk (this step)
  const unsigned* target;    // warps per element chunk
  const unsigned short* lo;  // per node tile: first and last element chunk it reads
  const unsigned short* hi;
  int nTiles;
};
template <bool ENERGY>
__global__ void __launch_bounds__(NODE_BLOCK, FTB_NODE_MINBLOCKS) k_node_ovl(const OvlArgs P) {
  const NodeArgs& A = P.N;
  DevScalars* sc = A.sc;
  if (sc->done | sc->last) return;  // the step does not run (the element kernel and k_adv apply the same test)
  const double dt1 = sc->nt_half - sc->nt_n, dt2 = sc->nt_np1 - sc->nt_half;
  __shared__ int s_ok;
  __shared__ double sw[2][3][NODE_BLOCK / 32];
  int upto = 0;  // thread 0: element chunks [0, upto) are known to be complete (tiles come in ascending order)
  int par = 0;
  for (int tile = blockIdx.x; tile < P.nTiles; tile += gridDim.x, par ^= 1) {
    const int n = tile * NODE_BLOCK + threadIdx.x;
    // everything that does not depend on the element kernel is requested before the wait
    unsigned fl = 0;
    double vv[3], aa[3], dd[3], fprev[3], fext[3], m = 1.0;
    int ent[8];
    if (n < A.nN) {
      fl = A.flags[n];
#pragma unroll
      for (int q = 0; q < 8; ++q) ent[q] = __ldg(A.ell + (size_t)q * A.nN + n);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        vv[c] = A.v[c][n];
        aa[c] = A.a[c][n];
        fext[c] = A.fe[c] ? A.fe[c][n] : 0.0;
        if (ENERGY) { dd[c] = A.du[c][n]; fprev[c] = A.fi[c][n]; }
      }
      m = A.m[n];
    }
    if (threadIdx.x == 0) {
      int ok = 1;
      const int c1 = P.hi[tile];
      if (c1 >= upto) {
        for (int c = upto; c <= c1 && ok; ++c) {
          const unsigned want = P.target[c];
          const volatile unsigned* q = P.ctr + c;
          if (*q < want) {
            const unsigned long long t0 = ovl_now_ns();
            while (*q < want) {
              __nanosleep(100);
              if (ovl_now_ns() - t0 > 2000000000ULL) { ok = 0; break; }  // 2 s: the element kernel is gone
            }
          }
        }
        upto = c1 + 1;
        __threadfence();  // acquire: the forces announced by the counters are visible to the loads below
      }
      s_ok = ok;
    }
    __syncthreads();
    if (!s_ok) {
      if (threadIdx.x == 0) atomicOr(&sc->status, 32);
      return;
    }
    double wke = 0.0, wint = 0.0, wext = 0.0;
    if (n < A.nN) {
      double f[3] = {0.0, 0.0, 0.0};
      double fv[8][3];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int en = ent[q] < 0 ? 0 : ent[q];
#pragma unroll
        for (int c = 0; c < 3; ++c)  // written by the concurrently running element kernel: L2 loads, never the read-only path
          fv[q][c] = (ent[q] >= 0) ? __ldcg(A.felem + FTB_FIDX(3 * (en & 7) + c, en >> 3)) : 0.0;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int c = 0; c < 3; ++c) f[c] += fv[q][c];
      if (fl & FTB_FLAG_OVERFLOW)
        for (int j = A.node_off[n] + 8, j1 = A.node_off[n + 1]; j < j1; ++j) {
          const int en = __ldg(A.node_ent + j);
#pragma unroll
          for (int c = 0; c < 3; ++c) f[c] += __ldcg(A.felem + FTB_FIDX(3 * (en & 7) + c, en >> 3));
        }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const bool b = (fl >> c) & 1u;
        const double fnet = fext[c] - f[c];
        const double a_old = aa[c];
        if (!b) aa[c] = fnet / m;
        if (!b) {
          const double vhalf = vv[c] + dt1 * a_old;
          vv[c] = vhalf + dt2 * aa[c];
        }
        if (ENERGY && !(fl & FTB_FLAG_NOTOWNED)) {
          wke += m * vv[c] * vv[c];
          if (b) wext += dd[c] * (fprev[c] + f[c] + m * (aa[c] + a_old));
          wint += dd[c] * (fprev[c] + f[c]);
          wext += dd[c] * (fext[c] + fext[c]);
        }
        if (A.store_fi) A.fi[c][n] = f[c];
        A.v[c][n] = vv[c];
        A.a[c][n] = aa[c];
      }
    }
    if (ENERGY) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        wke += __shfl_down_sync(0xffffffffu, wke, o);
        wint += __shfl_down_sync(0xffffffffu, wint, o);
        wext += __shfl_down_sync(0xffffffffu, wext, o);
      }
      if ((threadIdx.x & 31) == 0) {
        sw[par][0][threadIdx.x >> 5] = wke;
        sw[par][1][threadIdx.x >> 5] = wint;
        sw[par][2][threadIdx.x >> 5] = wext;
      }
      __syncthreads();  // the other parity's buffer is free again by the time any warp gets here next
      if (threadIdx.x == 0) {
        double s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
        for (int w = 0; w < NODE_BLOCK / 32; ++w) {
          s0 += sw[par][0][w];
          s1 += sw[par][1][w];
          s2 += sw[par][2][w];
        }
        A.epart[tile] = s0;  // the serial kernel's layout: partial of node block `tile`
        A.epart[P.nTiles + tile] = s1;
        A.epart[2 * P.nTiles + tile] = s2;
      }
    } else {
      __syncthreads();  // s_ok is rewritten by thread 0 at the top of the next tile
    }
  }
}
// =============================================================================================
// Pipelined resident loop: two PERSISTENT kernels per step running concurrently on two streams.
//
//   k_elem_pipe (fp64 bound)  tiles of 64 elements, in chunk order: gathers X,u,v,a,flags of the 8 nodes,
//                             performs the first kick + drift + BC of the step on the fly (so the node
//                             arrays are only read), element forces -> felem, element dt -> min.  Its last
//                             block does the scalar bookkeeping of the time loop (old k_adv).
//   k_node_pipe (HBM bound)   tiles of 128 nodes, in group order: node group g holds the nodes whose
//                             elements all lie in chunks <= g, so it may run as soon as k_elem_pipe has
//                             finished chunk g of the SAME step -- while later chunks are still being
//                             computed.  Repeats the (bitwise identical) drift, gathers f_int through the CSR
//                             map, a = f/m, second kick, energy partials; writes u, v, a.
//
// Dependencies are tracked with per-chunk completion counters in device memory (release: stores ->
// __syncthreads -> __threadfence -> atomicAdd; acquire: volatile poll -> __threadfence -> __syncthreads ->
// ld.cg loads).  k_elem_pipe of step m+1 waits, per tile, for the node groups of step m that its
// elements touch (`need`), so the fp64 pipe never drains at a step boundary except for the dt reduction.
// Both grids are sized to be co-resident on every SM and take tiles from an atomic ticket, so any
// resident subset of blocks makes progress (no scheduling-order deadlock).
constexpr int PIPE_MAXC = 64;
constexpr int NODE_TILE = 128;

struct StepScal {
  double t_n, t_np1, t_half, dt;
};
struct PipeCtl {
  unsigned elem_done[2][PIPE_MAXC];
  unsigned node_done[2][PIPE_MAXC];
  unsigned elem_prefix[2];
  unsigned node_prefix[2];
  unsigned elem_ticket[2];
  unsigned node_ticket[2];
  unsigned elem_blocks_done, node_blocks_done;
  unsigned elem_target[PIPE_MAXC];  // tiles per element chunk
  unsigned node_target[PIPE_MAXC];  // tiles per node group
  unsigned need[PIPE_MAXC];         // element chunk c needs node groups <= need[c] of the previous step
  int C;
  int nTilesE, nTilesN;
  long long elem_step, node_step, stop_step;  // index of the next step of either kernel / first step not to run
  StepScal scal[2];                 // scal[m & 1]: times of step m
};

__device__ __forceinline__ void pipe_advance(unsigned* prefix, const unsigned* done, const unsigned* target, int C) {
  for (;;) {
    const unsigned p = *(volatile unsigned*)prefix;
    if ((int)p >= C) break;
    if (*(volatile const unsigned*)&done[p] != target[p]) break;
    atomicCAS(prefix, p, p + 1);
  }
}
__device__ __forceinline__ unsigned long long pipe_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// wait until *prefix > need_gt.  Bounded: if the two kernels are not co-resident (they must be, see the
// grid sizing in ftb200_shape_functions) the wait gives up after 2 s, flags status bit 32 and stops the
// loop instead of hanging the device.
__device__ __forceinline__ bool pipe_wait(const unsigned* prefix, unsigned need_gt, DevScalars* sc, PipeCtl* ctl) {
  if (*(volatile const unsigned*)prefix <= need_gt) {
    const unsigned long long t0 = pipe_now_ns();
    while (*(volatile const unsigned*)prefix <= need_gt) {
      __nanosleep(40);
      if (pipe_now_ns() - t0 > 2000000000ULL || (*(volatile int*)&sc->status & 32)) {
        atomicOr(&sc->status, 32);
        *(volatile long long*)&ctl->stop_step = 0;
        return false;
      }
    }
  }
  __threadfence();
  return true;
}

// first kick + drift + BC of one dof (Benchmarking-Parallel.cpp:115-122,131-135,184-244).  Explicit
// intrinsics: k_elem_pipe and k_node_pipe must produce the SAME bits.
__device__ __forceinline__ double pipe_drift(const double u, const double v, const double a, const bool b,
                                             const unsigned kind, const double dt1, const double dt, const double T,
                                             const double* __restrict__ bc_rate) {
  double un = u;
  if (!b) un = __fma_rn(dt, __fma_rn(dt1, a, v), u);
  if (kind) un = __dmul_rn(T, bc_rate[kind]);
  return un;
}

struct PipeElemArgs {
  ElemArgs E;
  const double* v[3];
  const double* a[3];
  const uint16_t* flags;
  const uint8_t* tile_chunk;  // element tile -> chunk
  PipeCtl* ctl;
  double* dt_hist;
  int nPID;
};

#ifndef FTB_PIPE_ELEM_REGS
#define FTB_PIPE_ELEM_REGS 136
#endif
// 136 registers: three warps of this kernel use 13056 of the 16384 registers of an SM sub-partition and
// leave room for one warp of k_node_pipe (<= 96 registers) -- the two kernels must be co-resident.
template <int MATSEL>
__global__ void __maxnreg__(FTB_PIPE_ELEM_REGS) k_elem_pipe(const PipeElemArgs P) {
  const ElemArgs& A = P.E;
  PipeCtl* ctl = P.ctl;
  DevScalars* sc = A.sc;
  const long long m = *(volatile long long*)&ctl->elem_step;
  if (m >= *(volatile long long*)&ctl->stop_step) return;
  const int p = (int)(m & 1);
  const StepScal S = ctl->scal[p];
  const double dt1 = S.t_half - S.t_n;
  __shared__ double sm_cols[72][ELEM_BLOCK];
  __shared__ unsigned s_tile;
  __shared__ int s_last, s_ok;
  const size_t E = (size_t)A.nE;
  int status = 0;
  unsigned long long bmin = 0x7FF0000000000000ULL;
  for (;;) {
    if (threadIdx.x == 0) s_tile = atomicAdd(&ctl->elem_ticket[p], 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    if (tile >= (unsigned)ctl->nTilesE) break;
    const int c = P.tile_chunk[tile];
    const int e = (int)(tile * ELEM_BLOCK + threadIdx.x);
    int nd[8];
    if (e < A.nE) {
#pragma unroll
      for (int k = 0; k < 8; ++k) nd[k] = __ldg(A.conn + (size_t)k * E + e);
    }
    // the node groups of the previous step that this chunk touches must be complete
    if (threadIdx.x == 0) s_ok = pipe_wait(&ctl->node_prefix[p ^ 1], ctl->need[c], sc, ctl) ? 1 : 0;
    __syncthreads();
    if (!s_ok) return;
    double dte = 1e300;
    if (e < A.nE) {
      double X[8][3], U[8][3];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const unsigned fl = P.flags[nd[k]];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          X[k][cc] = __ldg(A.X[cc] + nd[k]);
          // plain (L1-allocating) loads are safe: a 128-byte line holds 16 consecutive nodes of ONE node
          // tile, i.e. of one group, and this tile has waited for every group it touches
          const double u = A.u[cc][nd[k]];
          const double v = P.v[cc][nd[k]];
          const double a = P.a[cc][nd[k]];
          U[k][cc] = pipe_drift(u, v, a, (fl >> cc) & 1u, (fl >> (4 + 2 * cc)) & 3u, dt1, S.dt, S.t_np1, sc->bc_rate);
        }
      }
      const int pp = __ldg(A.pid + e);
      const double* mp = A.mp + (size_t)pp * FTB_MP_STRIDE;
      const int mat = (MATSEL >= 0) ? MATSEL : (int)mp[MP_MATID];
      double fe[8][3];
      DevHist h{A.hist, E, (size_t)e};
      SmemScratch Sc{&sm_cols[0][threadIdx.x]};
      double d;
      status |= hex8_element<MATSEL, true>(X, U, mat, mp, true, h, NoOutput(), Sc, fe, &d);
      dte = d;
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) __stcg(A.felem + FTB_FIDX(3 * k + cc, e), fe[k][cc]);
      if (__ldg(A.eflag + e)) dte = 1e300;  // element skipped, StableTimeStep.cpp:13-19
    }
    unsigned long long b = dt_to_bits(dte);
    bmin = b < bmin ? b : bmin;
    __syncthreads();  // every store of the tile has been issued
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned old = atomicAdd(&ctl->elem_done[p][c], 1u);
      if (old + 1 == ctl->elem_target[c]) pipe_advance(&ctl->elem_prefix[p], ctl->elem_done[p], ctl->elem_target, ctl->C);
    }
  }
  // element dt: one atomic per warp for the whole kernel
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, bmin, o);
    bmin = t < bmin ? t : bmin;
  }
  if ((threadIdx.x & 31) == 0) atomicMin(&sc->dtmin_bits, bmin);
  if (status) atomicOr(&sc->status, status);
  // ---- last block: scalar bookkeeping of the time loop (Benchmarking-Parallel.cpp:106-112,168; StableTimeStep.cpp:33-38)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(&ctl->elem_blocks_done, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __shared__ double s_ndt;
  if (threadIdx.x == 0) {
    __threadfence();
    double dtmin = __longlong_as_double((long long)*(volatile unsigned long long*)&sc->dtmin_bits);
    if (dtmin > 1e20) dtmin = 1e20;
    sc->dtmin_bits = 0x7FF0000000000000ULL;
    // step m is complete as far as the elements are concerned
    sc->t_n = S.t_n; sc->t_np1 = S.t_np1; sc->t_half = S.t_half; sc->dt = S.dt;
    sc->Time = S.t_np1;
    if (P.dt_hist && sc->step < sc->hist_cap) P.dt_hist[sc->step] = S.dt;
    sc->step += 1;
    sc->steps_left -= 1;
    int stop = 0;
    if (dtmin < sc->failure_dt) { sc->status |= 16; stop = 1; }  // TerminateFemTech(19)
    const double ndt = sc->reduction * dtmin;
    StepScal N;
    N.t_n = S.t_np1; N.dt = ndt; N.t_np1 = S.t_np1 + ndt; N.t_half = 0.5 * (N.t_np1 + N.t_n);
    ctl->scal[p ^ 1] = N;
    sc->ndt = ndt; sc->nt_n = N.t_n; sc->nt_np1 = N.t_np1; sc->nt_half = N.t_half;
    if (!(sc->Time < sc->tMax) || sc->steps_left <= 0) stop = 1;
    if (stop) ctl->stop_step = m + 1;
    // recycle the counters of the other parity (their readers are done: every tile above waited for them)
    for (int i = 0; i < ctl->C; ++i) { ctl->elem_done[p ^ 1][i] = 0; ctl->node_done[p ^ 1][i] = 0; }
    ctl->elem_prefix[p ^ 1] = 0; ctl->node_prefix[p ^ 1] = 0;
    // tickets are recycled by the kernel that draws them (every block of THIS launch has left its loop);
    // a straggler of k_node_pipe(m-1) may still draw from node_ticket[p ^ 1]
    ctl->elem_ticket[p] = 0;
    pipe_advance(&ctl->elem_prefix[p ^ 1], ctl->elem_done[p ^ 1], ctl->elem_target, ctl->C);
    pipe_advance(&ctl->node_prefix[p ^ 1], ctl->node_done[p ^ 1], ctl->node_target, ctl->C);
    ctl->elem_blocks_done = 0;
    s_ndt = ndt;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < P.nPID; q += ELEM_BLOCK) {  // Prony factors of the next dt
    double* mq = const_cast<double*>(A.mp) + (size_t)q * FTB_MP_STRIDE;
    if ((int)mq[MP_MATID] == 5) {
      const double rt1 = s_ndt / mq[MP_T1], rt2 = s_ndt / mq[MP_T2];
      const double c11 = exp(-rt1), c12 = exp(-rt2);
      mq[MP_C11] = c11; mq[MP_C12] = c12;
      mq[MP_C21] = mq[MP_G1] * (1 - c11) / rt1;
      mq[MP_C22] = mq[MP_G2] * (1 - c12) / rt2;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    *(volatile long long*)&ctl->elem_step = m + 1;
  }
}

struct PipeNodeArgs {
  NodeArgs N;
  const int* ell;             // [8][nN] first eight CSR entries of every node (-1 = none), plane q = q-th entry
  const uint8_t* tile_group;  // node tile -> group
  PipeCtl* ctl;
  double* etile;              // [3][nTilesN] energy partials per node tile
  double* ehist;
  int energy;
};

__global__ void __maxnreg__(96) k_node_pipe(const PipeNodeArgs P) {
  const NodeArgs& A = P.N;
  PipeCtl* ctl = P.ctl;
  DevScalars* sc = A.sc;
  const long long m = *(volatile long long*)&ctl->node_step;
  if (m >= *(volatile long long*)&ctl->stop_step) return;
  const int p = (int)(m & 1);
  const StepScal S = ctl->scal[p];
  const double dt1 = S.t_half - S.t_n, dt2 = S.t_np1 - S.t_half;
  __shared__ unsigned s_tile;
  __shared__ int s_last, s_ok;
  __shared__ double sw[3][NODE_TILE / 32];
  const size_t E = (size_t)A.nE;
  const int nT = ctl->nTilesN;
  for (;;) {
    if (threadIdx.x == 0) s_tile = atomicAdd(&ctl->node_ticket[p], 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    if (tile >= (unsigned)nT) break;
    const int g = P.tile_group[tile];
    const int n = (int)(tile * NODE_TILE + threadIdx.x);
    // state of the previous step (written by an earlier kernel) can be loaded before the wait
    const unsigned fl = A.flags[n];
    double uu[3], vv[3], aa[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { uu[c] = __ldcg(A.u[c] + n); vv[c] = __ldcg(A.v[c] + n); aa[c] = __ldcg(A.a[c] + n); }
    const double mass = A.m[n];
    // static gather map in ELL form (eight entries per node, coalesced planes): no dependent offset load,
    // so everything above and these entries are ONE round of independent loads issued before the wait
    int ent[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) ent[q] = __ldg(P.ell + (size_t)q * A.nN + n);
    if (threadIdx.x == 0) s_ok = pipe_wait(&ctl->elem_prefix[p], (unsigned)g, sc, ctl) ? 1 : 0;
    __syncthreads();
    if (!s_ok) return;
    // 24 independent loads in flight per thread, then the sum in ascending element order
    // (GetForce_3D.cpp:15,39-44); adding an exact 0.0 for a missing entry does not change the result
    double fv[8][3];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const size_t e = (size_t)((ent[q] < 0 ? 0 : ent[q]) >> 3);
      const int sl = (ent[q] < 0 ? 0 : ent[q]) & 7;
#pragma unroll
      for (int c = 0; c < 3; ++c) fv[q][c] = (ent[q] >= 0) ? A.felem[FTB_FIDX(3 * sl + c, e)] : 0.0;
    }
    double f[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int c = 0; c < 3; ++c) f[c] += fv[q][c];
    if (fl & FTB_FLAG_OVERFLOW)
    for (int j = A.node_off[n] + 8, j1 = A.node_off[n + 1]; j < j1; ++j) {
      const int en = __ldg(A.node_ent + j);
      const size_t e = (size_t)(en >> 3);
      const int sl = en & 7;
#pragma unroll
      for (int c = 0; c < 3; ++c) f[c] += A.felem[FTB_FIDX(3 * sl + c, e)];
    }
    double wke = 0.0, wint = 0.0, wext = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const bool b = (fl >> c) & 1u;
      const unsigned kind = (fl >> (4 + 2 * c)) & 3u;
      const double u_old = uu[c], a_old = kind ? 0.0 : aa[c];
      const double un = pipe_drift(u_old, vv[c], aa[c], b, kind, dt1, S.dt, S.t_np1, sc->bc_rate);
      double vn = vv[c], an = aa[c];
      if (kind) { vn = sc->bc_rate[kind]; an = 0.0; }  // ApplyBoundaryConditions, :184-244
      const double fext = A.fe[c] ? A.fe[c][n] : 0.0;
      const double fnet = fext - f[c];                 // GetForce_3D.cpp:11,49-51
      if (!b) {
        const double vhalf = __fma_rn(dt1, aa[c], vv[c]);
        an = fnet / mass;                              // CalculateAcclerations.cpp:7-11
        vn = vhalf + dt2 * an;                         // :146-151
      }
      if (P.energy && !(fl & FTB_FLAG_NOTOWNED)) {     // CheckEnergy.cpp:19-52
        const double dd = un - u_old;
        const double fprev = A.fi[c][n];
        wke += mass * vn * vn;
        if (b) wext += dd * (fprev + f[c] + mass * (an + a_old));
        wint += dd * (fprev + f[c]);
        wext += dd * (fext + fext);
      }
      A.u[c][n] = un; A.v[c][n] = vn; A.a[c][n] = an;
      if (A.store_fi) A.fi[c][n] = f[c];
    }
    if (P.energy) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        wke += __shfl_down_sync(0xffffffffu, wke, o);
        wint += __shfl_down_sync(0xffffffffu, wint, o);
        wext += __shfl_down_sync(0xffffffffu, wext, o);
      }
      if ((threadIdx.x & 31) == 0) { sw[0][threadIdx.x >> 5] = wke; sw[1][threadIdx.x >> 5] = wint; sw[2][threadIdx.x >> 5] = wext; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (P.energy) {
        double s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
        for (int w = 0; w < NODE_TILE / 32; ++w) { s0 += sw[0][w]; s1 += sw[1][w]; s2 += sw[2][w]; }
        P.etile[tile] = s0; P.etile[nT + tile] = s1; P.etile[2 * nT + tile] = s2;
      }
      __threadfence();
      const unsigned old = atomicAdd(&ctl->node_done[p][g], 1u);
      if (old + 1 == ctl->node_target[g]) pipe_advance(&ctl->node_prefix[p], ctl->node_done[p], ctl->node_target, ctl->C);
    }
  }
  // ---- last block: fixed-order energy reduction (K8) and step counter
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(&ctl->node_blocks_done, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  if (P.energy) {
    __shared__ double sh[3][NODE_TILE];
    __threadfence();
    double s[3] = {0, 0, 0};
    for (int i = threadIdx.x; i < nT; i += NODE_TILE) {
      s[0] += __ldcg(P.etile + i); s[1] += __ldcg(P.etile + nT + i); s[2] += __ldcg(P.etile + 2 * nT + i);
    }
    sh[0][threadIdx.x] = s[0]; sh[1][threadIdx.x] = s[1]; sh[2][threadIdx.x] = s[2];
    __syncthreads();
    for (int o = NODE_TILE / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) {
        sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
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
      if (P.ehist && m >= 0 && m < sc->hist_cap) {
        P.ehist[4 * m + 0] = sc->Wint; P.ehist[4 * m + 1] = sc->Wext; P.ehist[4 * m + 2] = WKE; P.ehist[4 * m + 3] = sc->Etot;
      }
    }
  }
  if (threadIdx.x == 0) {
    ctl->node_blocks_done = 0;
    ctl->node_ticket[p] = 0;  // every block of this launch has left its loop
    __threadfence();
    *(volatile long long*)&ctl->node_step = m + 1;
  }
}

// tuning probe: mark every chunk / group complete so that one of the pipe kernels can be timed alone
__global__ void k_pipe_debug_arm(PipeCtl* ctl, DevScalars* sc) {
  for (int q = 0; q < 2; ++q) {
    for (int i = 0; i < ctl->C; ++i) { ctl->elem_done[q][i] = ctl->elem_target[i]; ctl->node_done[q][i] = ctl->node_target[i]; }
    ctl->elem_prefix[q] = ctl->C; ctl->node_prefix[q] = ctl->C; ctl->elem_ticket[q] = 0; ctl->node_ticket[q] = 0;
  }
  ctl->elem_blocks_done = 0; ctl->node_blocks_done = 0;
  ctl->stop_step = 0x7FFFFFFFFFFFFFFFLL;
  sc->steps_left = 1 << 30;
  sc->tMax = 1e300;
}

// (re)arm the pipeline counters at the start of a run call
__global__ void k_pipe_begin(DevScalars* sc, PipeCtl* ctl, double tMax, long long steps) {
  sc->tMax = tMax;
  sc->steps_left = steps;
  const bool none = (!(sc->Time < tMax) || steps <= 0 || (sc->status & 16));
  ctl->stop_step = none ? ctl->elem_step : 0x7FFFFFFFFFFFFFFFLL;
  ctl->node_step = ctl->elem_step;
  const int p = (int)(ctl->elem_step & 1);
  for (int q = 0; q < 2; ++q) {
    for (int i = 0; i < ctl->C; ++i) { ctl->elem_done[q][i] = 0; ctl->node_done[q][i] = (q == (p ^ 1)) ? ctl->node_target[i] : 0; }
    ctl->elem_prefix[q] = 0; ctl->node_prefix[q] = 0; ctl->elem_ticket[q] = 0; ctl->node_ticket[q] = 0;
    pipe_advance(&ctl->elem_prefix[q], ctl->elem_done[q], ctl->elem_target, ctl->C);
    pipe_advance(&ctl->node_prefix[q], ctl->node_done[q], ctl->node_target, ctl->C);
  }
  ctl->elem_blocks_done = 0; ctl->node_blocks_done = 0;
  // times of the first step of this run: the scalars left by explicit_begin / the previous run
  StepScal S; S.t_n = sc->nt_n; S.t_np1 = sc->nt_np1; S.t_half = sc->nt_half; S.dt = sc->ndt;
  ctl->scal[p] = S;
}

// =============================================================================================
// Fused step kernel (single GPU, resident loop): ONE launch per time step.  Every warp is an independent
// worker that alternates between
//   * an ELEMENT tile (32 elements, fp64 bound): gathers X,u,v,a,flags, does the first kick + drift + BC of
//     the step on the fly, element forces -> felem, element dt, and
//   * a NODE tile (32 nodes, HBM bound) whose dependency group is complete: drift (bitwise identical),
//     CSR/ELL gather of f_int, a = f/m, second kick, energy partials, writes u, v, a.
// While a warp waits on the memory of its node tile, the other warps of the SM keep the fp64 pipe busy, so the
// HBM-bound node work is hidden under the fp64-bound element work without a second kernel.  Node group g
// (nodes whose elements all lie in element chunks <= g) becomes ready when chunk g is complete; completion
// is published per warp with __threadfence + atomicAdd, readiness is a prefix counter.  No warp ever waits
// while element tiles remain, all tickets are dynamic, and the step's previous state is complete at launch,
// so there is no cross-launch dependency and no co-residency requirement.  The warp that finishes last
// performs the scalar update of the time loop; the energy partials (one per node tile, fixed order) are
// reduced by k_energy_tiles on a second stream, off the critical path.
constexpr int STEP_MAXC = 64;
struct StepCtl {
  unsigned elem_ticket, node_ticket;
  unsigned elem_prefix;
  unsigned warps_done;
  unsigned elem_done[STEP_MAXC];
  unsigned elem_target[STEP_MAXC];  // 32-element tiles per chunk
  int C, nTilesE, nTilesN;
  unsigned energy_blocks_done;
  int pad;
  long long energy_step;           // steps whose energy partials have been reduced (k_energy_tiles only)
};
struct StepArgs {
  ElemArgs E;
  NodeArgs N;
  const int* ell;
  const uint8_t* etile_chunk;  // 32-element tile -> chunk
  const uint8_t* ntile_group;  // 32-node tile -> group
  StepCtl* ctl;
  double* etile;               // [2][3][nTilesN] energy partials, double buffered by step parity
  double* dt_hist;
  int nPID, energy;
};

template <int MATSEL>
__global__ void __launch_bounds__(ELEM_BLOCK, (MATSEL == 5 || MATSEL < 0) ? 4 : ELEM_MINBLOCKS) k_step(const StepArgs P) {
  const ElemArgs& A = P.E;
  const NodeArgs& Nd = P.N;
  StepCtl* ctl = P.ctl;
  DevScalars* sc = A.sc;
  const int lane = threadIdx.x & 31;
  if (sc->last | sc->done) {  // dead iteration of a graph replay
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      if (!sc->done && sc->last) sc->done = 1;
      sc->active = 0;
    }
    return;
  }
  // times of THIS step (left by the previous launch / explicit_begin)
  const double t_n = sc->nt_n, t_np1 = sc->nt_np1, t_half = sc->nt_half, dt = sc->ndt;
  const double dt1 = t_half - t_n, dt2 = t_np1 - t_half;
  const int parity = (int)(sc->step & 1);
  __shared__ double sm_cols[72][ELEM_BLOCK];
  const size_t E = (size_t)A.nE;
  const int nTilesE = ctl->nTilesE, nTilesN = ctl->nTilesN;
  int status = 0;
  unsigned long long bmin = 0x7FF0000000000000ULL;
  unsigned et = 0;
  if (lane == 0) et = atomicAdd(&ctl->elem_ticket, 1u);
  et = __shfl_sync(0xffffffffu, et, 0);
  // every warp holds ONE claimed node tile (plain atomicAdd ticket: no compare-and-swap storm) and runs it as soon as
  // its dependency group is complete; holders keep taking element tiles meanwhile, so the wait never blocks progress
  unsigned nt = 0;
  if (lane == 0) nt = atomicAdd(&ctl->node_ticket, 1u);
  nt = __shfl_sync(0xffffffffu, nt, 0);
  bool nodes_left = nt < (unsigned)nTilesN;
  unsigned long long t_wait0 = 0;
  for (;;) {
    const bool have_elem = et < (unsigned)nTilesE;
    if (have_elem) {
      unsigned et_next = 0;
      if (lane == 0) et_next = atomicAdd(&ctl->elem_ticket, 1u);  // next ticket: latency hidden by this tile
      const int c = P.etile_chunk[et];
      const int e = (int)(et * 32 + lane);
      double dte = 1e300;
      if (e < A.nE) {
        int nd[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) nd[k] = __ldg(A.conn + (size_t)k * E + e);
        double X[8][3], U[8][3];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const unsigned fl = Nd.flags[nd[k]];
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) {
            X[k][cc] = __ldg(A.X[cc] + nd[k]);
            U[k][cc] = pipe_drift(Nd.u[cc][nd[k]], Nd.v[cc][nd[k]], Nd.a[cc][nd[k]], (fl >> cc) & 1u, (fl >> (4 + 2 * cc)) & 3u,
                                  dt1, dt, t_np1, sc->bc_rate);
          }
        }
        const int pp = __ldg(A.pid + e);
        const double* mp = A.mp + (size_t)pp * FTB_MP_STRIDE;
        const int mat = (MATSEL >= 0) ? MATSEL : (int)mp[MP_MATID];
        double fe[8][3];
        DevHist h{A.hist, E, (size_t)e};
        SmemScratch Sc{&sm_cols[0][threadIdx.x]};
        double d;
        status |= hex8_element<MATSEL, true>(X, U, mat, mp, true, h, NoOutput(), Sc, fe, &d);
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) __stcg(A.felem + FTB_FIDX(3 * k + cc, e), fe[k][cc]);
        dte = __ldg(A.eflag + e) ? 1e300 : d;  // element skipped, StableTimeStep.cpp:13-19
      }
      const unsigned long long b = dt_to_bits(dte);
      bmin = b < bmin ? b : bmin;
      __threadfence();
      __syncwarp();
      if (lane == 0) {
        const unsigned old = atomicAdd(&ctl->elem_done[c], 1u);
        if (old + 1 == ctl->elem_target[c]) pipe_advance(&ctl->elem_prefix, ctl->elem_done, ctl->elem_target, ctl->C);
      }
      et = __shfl_sync(0xffffffffu, et_next, 0);
    }
    // ---- node tiles whose group is complete: one per element tile while elements remain, drain afterwards
    bool did_node = false;
    if (nodes_left) {
      const unsigned cand = nt;
      int state = 0;  // 0 not ready yet, 1 ready
      if (lane == 0) state = ((unsigned)P.ntile_group[cand] < *(volatile unsigned*)&ctl->elem_prefix) ? 1 : 0;
      state = __shfl_sync(0xffffffffu, state, 0);
      if (state == 1) {
        did_node = true;
        __threadfence();  // acquire: the element forces of the group are visible
        const int n = (int)(cand * 32 + lane);
        const unsigned fl = Nd.flags[n];
        double uu[3], vv[3], aa[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) { uu[c] = Nd.u[c][n]; vv[c] = Nd.v[c][n]; aa[c] = Nd.a[c][n]; }
        const double mass = Nd.m[n];
        int ent[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) ent[q] = __ldg(P.ell + (size_t)q * Nd.nN + n);
        double fv[8][3];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const size_t e = (size_t)((ent[q] < 0 ? 0 : ent[q]) >> 3);
          const int sl = (ent[q] < 0 ? 0 : ent[q]) & 7;
#pragma unroll
          for (int c = 0; c < 3; ++c) fv[q][c] = (ent[q] >= 0) ? __ldcg(A.felem + FTB_FIDX(3 * sl + c, e)) : 0.0;
        }
        double f[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int q = 0; q < 8; ++q)  // ascending element id (GetForce_3D.cpp:15,39-44); + 0.0 for a missing entry is exact
#pragma unroll
          for (int c = 0; c < 3; ++c) f[c] += fv[q][c];
        if (fl & FTB_FLAG_OVERFLOW)
          for (int j = Nd.node_off[n] + 8, j1 = Nd.node_off[n + 1]; j < j1; ++j) {
            const int en = __ldg(Nd.node_ent + j);
#pragma unroll
            for (int c = 0; c < 3; ++c) f[c] += __ldcg(A.felem + FTB_FIDX(3 * (en & 7) + c, en >> 3));
          }
        double wke = 0.0, wint = 0.0, wext = 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const bool bnd = (fl >> c) & 1u;
          const unsigned kind = (fl >> (4 + 2 * c)) & 3u;
          const double u_old = uu[c], a_old = kind ? 0.0 : aa[c];
          const double un = pipe_drift(u_old, vv[c], aa[c], bnd, kind, dt1, dt, t_np1, sc->bc_rate);
          double vn = vv[c], an = aa[c];
          if (kind) { vn = sc->bc_rate[kind]; an = 0.0; }  // ApplyBoundaryConditions, :184-244
          const double fext = Nd.fe[c] ? Nd.fe[c][n] : 0.0;
          const double fnet = fext - f[c];                  // GetForce_3D.cpp:11,49-51
          if (!bnd) {
            const double vhalf = __fma_rn(dt1, aa[c], vv[c]);
            an = fnet / mass;                               // CalculateAcclerations.cpp:7-11
            vn = vhalf + dt2 * an;                          // :146-151
          }
          if (P.energy && !(fl & FTB_FLAG_NOTOWNED)) {      // CheckEnergy.cpp:19-52
            const double dd = un - u_old;
            const double fprev = Nd.fi[c][n];
            wke += mass * vn * vn;
            if (bnd) wext += dd * (fprev + f[c] + mass * (an + a_old));
            wint += dd * (fprev + f[c]);
            wext += dd * (fext + fext);
          }
          Nd.u[c][n] = un; Nd.v[c][n] = vn; Nd.a[c][n] = an;
          if (Nd.store_fi) Nd.fi[c][n] = f[c];
        }
        if (P.energy) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            wke += __shfl_down_sync(0xffffffffu, wke, o);
            wint += __shfl_down_sync(0xffffffffu, wint, o);
            wext += __shfl_down_sync(0xffffffffu, wext, o);
          }
          if (lane == 0) {
            double* et3 = P.etile + (size_t)parity * 3 * nTilesN;
            et3[cand] = wke; et3[nTilesN + cand] = wint; et3[2 * nTilesN + cand] = wext;
          }
        }
        if (lane == 0) nt = atomicAdd(&ctl->node_ticket, 1u);
        nt = __shfl_sync(0xffffffffu, nt, 0);
        nodes_left = nt < (unsigned)nTilesN;
      }
    }
    if (!have_elem) {
      if (!nodes_left) break;
      if (!did_node) {  // drain phase: the remaining groups wait for the last element tiles (bounded)
        if (t_wait0 == 0) t_wait0 = pipe_now_ns_early();
        __nanosleep(200);
        if (pipe_now_ns_early() - t_wait0 > 2000000000ULL) { atomicOr(&sc->status, 32); break; }
      } else {
        t_wait0 = 0;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, bmin, o);
    bmin = t < bmin ? t : bmin;
  }
  int last = 0;
  if (lane == 0) {
    atomicMin(&sc->dtmin_bits, bmin);  // min is order independent: deterministic
    if (status) atomicOr(&sc->status, status);
    __threadfence();
    const unsigned total = gridDim.x * (ELEM_BLOCK / 32);
    last = (atomicAdd(&ctl->warps_done, 1u) == total - 1) ? 1 : 0;
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (!last) return;
  // ---- the warp that finishes last: scalar update of the loop (Benchmarking-Parallel.cpp:106-112,168)
  double ndt = 0.0;
  if (lane == 0) {
    __threadfence();
    ctl->warps_done = 0; ctl->elem_ticket = 0; ctl->node_ticket = 0; ctl->elem_prefix = 0;
    for (int i = 0; i < ctl->C; ++i) ctl->elem_done[i] = 0;
    sc->active = 1;
    ndt = adv_step(sc, P.dt_hist);
  }
  ndt = __shfl_sync(0xffffffffu, ndt, 0);
  prony_update(const_cast<double*>(A.mp), P.nPID, ndt, lane, 32);
}

// K8 for the fused step: deterministic two-level reduction of the per-node-tile partials (fixed ranges, fixed
// order), running on a second stream concurrently with the next step
constexpr int ENERGY_BLOCKS = 32;
__global__ void __launch_bounds__(256) k_energy_tiles(DevScalars* sc, StepCtl* ctl, const double* etile, double* eblock,
                                                      double* ehist) {
  const long long k = *(volatile long long*)&ctl->energy_step;  // launches are serialised on their stream: one step each
  if (k >= *(volatile long long*)&sc->step) return;             // dead iteration: nothing new to reduce
  const int nT = ctl->nTilesN;
  const double* et3 = etile + (size_t)(k & 1) * 3 * nT;
  __shared__ double sh[3][256];
  __shared__ int s_last;
  const int per = (nT + ENERGY_BLOCKS - 1) / ENERGY_BLOCKS;
  const int lo = blockIdx.x * per, hi = min(nT, lo + per);
  double s[3] = {0, 0, 0};
  for (int i = lo + threadIdx.x; i < hi; i += 256) { s[0] += et3[i]; s[1] += et3[nT + i]; s[2] += et3[2 * nT + i]; }
  sh[0][threadIdx.x] = s[0]; sh[1][threadIdx.x] = s[1]; sh[2][threadIdx.x] = s[2];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; sh[2][threadIdx.x] += sh[2][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    eblock[blockIdx.x] = sh[0][0]; eblock[ENERGY_BLOCKS + blockIdx.x] = sh[1][0]; eblock[2 * ENERGY_BLOCKS + blockIdx.x] = sh[2][0];
    __threadfence();
    s_last = (atomicAdd(&ctl->energy_blocks_done, 1u) == ENERGY_BLOCKS - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    double t0 = 0, t1 = 0, t2 = 0;
    for (int b = 0; b < ENERGY_BLOCKS; ++b) {
      t0 += __ldcg(eblock + b); t1 += __ldcg(eblock + ENERGY_BLOCKS + b); t2 += __ldcg(eblock + 2 * ENERGY_BLOCKS + b);
    }
    const double WKE = 0.5 * t0;
    sc->Wint += 0.5 * t1;
    sc->Wext += 0.5 * t2;
    sc->WKE = WKE;
    sc->Etot = fabs(WKE + sc->Wint - sc->Wext);
    if (ehist && k >= 0 && k < sc->hist_cap) {
      ehist[4 * k + 0] = sc->Wint; ehist[4 * k + 1] = sc->Wext; ehist[4 * k + 2] = WKE; ehist[4 * k + 3] = sc->Etot;
    }
    ctl->energy_blocks_done = 0;
    __threadfence();
    *(volatile long long*)&ctl->energy_step = k + 1;
  }
}

