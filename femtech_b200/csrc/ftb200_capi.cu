// ftb200_capi.cu -- the C-ABI (include/ftb200.h) over the CUDA kernels.
// Host logic only: allocation, layout conversion at the boundary, CSR / halo map construction,
// kernel sequencing (streams, CUDA graphs) and error mapping to the reference's abort codes.
#include "../../include/ftb200.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "ftb200_kernels.cuh"
#include "ftb200_brick.cuh"

using namespace ftb;

namespace {
constexpr int GRAPH_STEPS = 25;

struct ProfEvents {
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  std::vector<std::pair<size_t, size_t>> elem, node;  // (start,stop) indices
};
}  // namespace

struct ftb200_ctx {
  int rank = 0, nranks = 1, device = 0;
  cudaStream_t stream = nullptr, stream2 = nullptr, stream_lo = nullptr;  // main (high priority), helper (high), interior elements (low)
  bool own_stream = true;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  long long launches = 0;
  char err[512] = {0};
  int nN = 0, nE = 0, nPID = 0;
  int nNp = 0;  // padded internal node count (node groups are aligned to NODE_TILE)
  bool mesh_ok = false, mat_ok = false, shape_ok = false, begun = false, bc_ok = false;
  // host copies of the inputs
  std::vector<double> h_X;
  std::vector<int> h_conn, h_pid, h_matid;
  // mixed C3D8 / C3D4 meshes (ftb200_upload_mesh_mixed)
  std::vector<uint8_t> h_etype;   // 1 = tetrahedron, reference element order; empty = all hexahedra
  uint8_t* etype = nullptr;       // internal order
  int* gpoff = nullptr;           // [nE+1] Gauss points before each element, reference order
  bool has_tet = false;
  long long nGP = 0;              // Gauss points of the mesh (8 nE without tetrahedra)
  int nEb_hex = 0, nEi_hex = 0;   // hexahedra among the boundary / interior elements (they come first in each class)
  struct ElemRange { int e0, e1, tet, mat, affine; };
  std::vector<ElemRange> ranges;  // internal element order = runs of equal (class, element type, material, affine geometry)
  double* ring_host = nullptr;   // step ring (ftb200_step_ring): mapped pinned host memory, its device alias, records
  double* ring_dev = nullptr;
  long long ring_cap = 0;
  bool use_affine = true;  // FTB200_AFFINE=0: parallelepiped hexahedra go through the general kernel too
  bool use_nh = true;      // FTB200_NH=0: neo-Hookean / HGO parallelepipeds through k_elem_affine<MAT, false> instead of k_elem_affine_cj<MAT>
  long long nE_affine = 0;
  // rigid-body prescribed motion (ftb200_set_rigid_bc)
  DevRigid* rigid = nullptr;
  double *rigid_tab = nullptr, *aprev[3] = {nullptr, nullptr, nullptr};
  int rigid_count = 0;
  // injury criteria (ftb200_injury_begin)
  bool injury = false;
  double *inj_ps = nullptr, *inj_psxsr = nullptr, *inj_smin = nullptr, *inj_shear = nullptr, *inj_part = nullptr, *inj_hist = nullptr;
  uint8_t *inj_flags = nullptr, *inj_incl = nullptr;
  int* inj_parti = nullptr;
  InjState* inj_state = nullptr;
  double inj_thr[4] = {0.15, 0.30, 120.0, 28.0};
  std::vector<double> h_props;
  std::vector<int> h_sendProcessID, h_sendCum, h_sendNodeIndex;
  // device arrays
  double *X[3] = {0, 0, 0}, *u[3] = {0, 0, 0}, *v[3] = {0, 0, 0}, *a[3] = {0, 0, 0}, *fi[3] = {0, 0, 0};
  double *du[3] = {0, 0, 0}, *fnet[3] = {0, 0, 0}, *fe[3] = {0, 0, 0};
  bool has_fe = false;
  double* m = nullptr;
  uint16_t* flags = nullptr;
  int *conn = nullptr, *pid = nullptr, *ref_of = nullptr;
  int *d_nref = nullptr, *d_nint = nullptr;  // internal node -> caller's id (-1 = padding) and back
  std::vector<int> h_nint;
  int* d_ell = nullptr;  // fixed-width node -> (element, slot) map [8][nNp]
  // single-partition step: the energy reduction of step n runs on the helper stream under the element kernel of n + 1
  cudaEvent_t ev_nodes_done = nullptr, ev_energy_done = nullptr;
  bool energy_async = true, energy_pending = false, energy_async_now = true;  // _now: off for single-step calls (nothing to overlap)
  uint8_t* eflag = nullptr;
  double *felem = nullptr, *hist = nullptr, *mp = nullptr;
  int *node_off = nullptr, *node_ent = nullptr;
  DevScalars* sc = nullptr;
  double *dthist = nullptr, *ehist = nullptr, *epart = nullptr, *out3 = nullptr;
  long long hist_cap = 0;
  int node_blocks = 0;
  double* d_stage[3] = {0, 0, 0};  // 3N doubles each
  int* d_istage = nullptr;         // 3N ints
  double* d_big = nullptr;         // lazily sized scratch for legacy energy / gp outputs
  size_t d_big_bytes = 0;
  unsigned long long* d_detmin = nullptr;
  int* d_nonpos = nullptr;
  int uniform_mat = -1;
  bool has_visco = false;
  int energy = 0;
  // halo
  int halo_count = 0, nshared = 0, nE_boundary = 0;
  int *d_sendNodeIndex = nullptr, *halo_nodes = nullptr, *halo_off = nullptr, *halo_slot = nullptr,
      *halo_node_idx = nullptr;
  const double* halo_recv_cur = nullptr;
  double *d_xsend = nullptr, *d_xrecv = nullptr;  // exchange windows of the host-buffer (MPI host) entry points
  // peer-memory transport
  char* p2p_window = nullptr;
  size_t p2p_bytes = 0;
  P2PArgs p2p;
  bool p2p_ready = false;
  unsigned long long* d_seq = nullptr;
  // exchange fused into the element kernels (PackArgs): device copy of the arguments, counters, and the switch that
  // elem_args consults while launch_step_p2p launches the step's element kernels
  PackArgs* d_pk = nullptr;
  unsigned* d_pk_ctr = nullptr;
  bool p2p_fused = false, pk_on = false;
  // launch attributes of the partitioned step (launch_k): 0 none, 1 priority la_prio, 2 programmatic event la_event
  int la_kind = 0, la_prio = 0, la_err = 0;
  cudaEvent_t la_event = nullptr, ev_prog = nullptr;
  int prio_hi = 0, prio_lo = 0;
  int p2p_order = 0;  // FTB200_P2P_ORDER: how the boundary elements get ahead of the interior (launch_step_p2p)
  unsigned long long* trace = nullptr;  // FTB200_P2P_TRACE: [TRACE_STEPS][TRACE_SLOTS] time stamps of the partitioned step
  std::string trace_prefix;
  unsigned* d_p2p_blocks = nullptr;
  std::vector<void*> p2p_opened;
  cudaGraphExec_t p2p_graph = nullptr;
  int p2p_graph_energy = -1;
  long long p2p_graph_launches = 0;
  double Time0 = 0.0;
  // brick-fused step (ftb200_brick.cuh): built by shape_functions when the mesh qualifies
  bool brick_ok = false, brick_want = false, felem_stale = false;  // FTB200_BRICK=1 selects the brick-fused step
  int brick_dims[3] = {10, 5, 5};
  int nB = 0, nIntTot = 0, nSurf = 0, surf_blocks = 0;
  long long nSlots = 0;
  BrickHdr* b_hdr = nullptr;
  uint16_t *b_conn16 = nullptr, *b_map16 = nullptr;
  int *b_halo = nullptr, *b_pe = nullptr, *s_ell = nullptr, *s_ovoff = nullptr, *s_ovent = nullptr;
  unsigned* b_flags32 = nullptr;
  int brick_grid = 0;
  double* b_part[3] = {nullptr, nullptr, nullptr};
  // graph cache
  cudaGraphExec_t graph = nullptr;
  int graph_energy = -1;
  // profiling
  bool profile = false;
  ProfEvents prof;
  double prof_elem_ms = 0, prof_node_ms = 0;
  long long prof_elem_n = 0, prof_node_n = 0;
};

namespace {

int fail(ftb200_ctx* c, int code, const char* fmt, ...) {
  if (c) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(c->err, sizeof(c->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

#define CK(call)                                                                                        \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess)                                                                              \
      return fail(ctx, e_ == cudaErrorMemoryAllocation ? FTB200_ERR_ALLOC : FTB200_ERR_CUDA, "%s: %s (%s:%d)", #call, \
                  cudaGetErrorString(e_), __FILE__, __LINE__);                                          \
  } while (0)

// Every kernel goes through launch_k (it counts the launches: `gpu_launches` of the bench line).
template <class... P, class... A>
inline void launch_k(ftb200_ctx* ctx, void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t strm, A&&... args) {
  if (ctx->la_kind) {
    // launch attributes of the partitioned step (launch_step_p2p): an explicit priority, or a programmatic event that
    // fires when every block of this grid has STARTED (the interior elements wait for it on their own stream)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = strm;
    cudaLaunchAttribute at[1];
    if (ctx->la_kind == 1) {
      at[0].id = cudaLaunchAttributePriority;
      at[0].val.priority = ctx->la_prio;
    } else {
      at[0].id = cudaLaunchAttributeProgrammaticEvent;
      at[0].val.programmaticEvent.event = ctx->la_event;
      at[0].val.programmaticEvent.flags = 0;
      at[0].val.programmaticEvent.triggerAtBlockStart = 1;
    }
    cfg.attrs = at; cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, static_cast<P>(args)...);
    if (e != cudaSuccess && !ctx->la_err) ctx->la_err = (int)e;
  } else {
    kern<<<grid, block, smem, strm>>>(static_cast<P>(args)...);
  }
  ctx->launches++;
}
#define LAUNCH(kern, grid, block, strm, ...)                                              \
  do {                                                                                    \
    auto kfn_ = kern;                                                                     \
    launch_k(ctx, kfn_, dim3(grid), dim3(block), (size_t)0, (strm), __VA_ARGS__);         \
  } while (0)

// material-5 element kernels: the two-stage history buffer is dynamic shared memory beyond the 48 KB static limit
#define LAUNCH_HIST(kern, grid, block, strm, ...)                                                        \
  do {                                                                                                   \
    auto kfn_ = kern;                                                                                    \
    /* the attribute is per device: a process may hold contexts on several GPUs, so no process-wide cache */ \
    cudaFuncSetAttribute(kfn_, cudaFuncAttributeMaxDynamicSharedMemorySize, HIST_STAGE_BYTES);           \
    launch_k(ctx, kfn_, dim3(grid), dim3(block), (size_t)HIST_STAGE_BYTES, (strm), __VA_ARGS__);         \
  } while (0)

inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

// The captured step graphs bake device pointers and flags into their kernel arguments: whoever reallocates one of those
// arrays (history, external force, injury, rigid-body state) or flips such a flag drops them; the next run rebuilds.
void drop_graphs(ftb200_ctx* ctx) {
  if (ctx->graph) { cudaGraphExecDestroy(ctx->graph); ctx->graph = nullptr; }
  if (ctx->p2p_graph) { cudaGraphExecDestroy(ctx->p2p_graph); ctx->p2p_graph = nullptr; }
}

template <class T>
int dalloc(ftb200_ctx* ctx, T** p, size_t n) {
  if (*p) { cudaFree(*p); *p = nullptr; }
  CK(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
  return 0;
}
template <class T>
void dfree(T*& p) {
  if (p) cudaFree(p);
  p = nullptr;
}

// Every copy and fill of the set-up code is ordered on the context's OWN stream.  The streams are non-blocking (several
// contexts may live in one process, some with kernels that wait for a peer), so they do not synchronise with the legacy
// default stream -- and a plain cudaMemcpy from pageable memory returns once the data are staged, before the DMA has
// landed, a plain cudaMemset is asynchronous altogether: a kernel launched on the context's stream right behind them
// could read the old contents (seen: the node list of the rigid-body condition on a first, cold context).
cudaError_t ftb_memcpy(ftb200_ctx* ctx, void* dst, const void* src, size_t n, cudaMemcpyKind kind) {
  cudaError_t e = cudaMemcpyAsync(dst, src, n, kind, ctx->stream);
  return e != cudaSuccess ? e : cudaStreamSynchronize(ctx->stream);
}
cudaError_t ftb_memset(ftb200_ctx* ctx, void* p, int v, size_t n) { return cudaMemsetAsync(p, v, n, ctx->stream); }

int ensure_big(ftb200_ctx* ctx, size_t bytes) {
  if (ctx->d_big_bytes >= bytes) return 0;
  dfree(ctx->d_big);
  CK(cudaMalloc((void**)&ctx->d_big, bytes));
  ctx->d_big_bytes = bytes;
  return 0;
}

ElemArgs elem_args(ftb200_ctx* c, int e0, int e1, int ignore) {
  ElemArgs A;
  A.pk = c->pk_on ? c->d_pk : nullptr;
  A.pk_nEb = c->pk_on ? c->nE_boundary : 0;
  for (int k = 0; k < 3; ++k) { A.X[k] = c->X[k]; A.u[k] = c->u[k]; }
  A.conn = c->conn; A.pid = c->pid; A.eflag = c->eflag; A.mp = c->mp; A.felem = c->felem; A.hist = c->hist;
  A.sc = c->sc; A.nE = c->nE; A.e0 = e0; A.e1 = e1; A.ignore_loop_flags = ignore;
  A.etype = c->etype;
  A.inj_ps = c->inj_ps; A.inj_psxsr = c->inj_psxsr; A.inj_smin = c->inj_smin; A.inj_shear = c->inj_shear;
  A.inj_flags = c->inj_flags; A.inj_incl = c->inj_incl;
  for (int k = 0; k < 4; ++k) A.inj_thr[k] = c->inj_thr[k];
  return A;
}
NodeArgs node_args(ftb200_ctx* c, const double* recv) {
  NodeArgs A;
  for (int k = 0; k < 3; ++k) {
    A.u[k] = c->u[k]; A.v[k] = c->v[k]; A.a[k] = c->a[k]; A.fi[k] = c->fi[k]; A.du[k] = c->du[k];
    A.fe[k] = c->has_fe ? c->fe[k] : nullptr;
  }
  A.m = c->m; A.flags = c->flags; A.felem = c->felem; A.node_off = c->node_off; A.node_ent = c->node_ent;
  A.halo_recv = recv;
  A.halo_off = c->halo_off; A.halo_slot = c->halo_slot;
  A.halo_node_idx = recv ? c->halo_node_idx : nullptr;
  A.epart = c->epart; A.sc = c->sc; A.nN = c->nNp; A.nE = c->nE;
  A.store_fi = c->energy ? 1 : 0;
  for (int k = 0; k < 3; ++k) { A.X[k] = c->X[k]; A.aprev[k] = c->aprev[k]; }
  A.rigid = c->rigid;
  A.ell = c->d_ell;
  A.halo_recv_alt = nullptr;
  A.p2p_seq = nullptr;
  return A;
}

// element kernel dispatch on the (uniform) material of the launch
template <bool WITH_FORCE, bool WITH_DT>
void launch_elem_hex(ftb200_ctx* ctx, cudaStream_t s, int e0, int e1, int ignore, int mat, int affine);

template <bool WITH_FORCE, bool WITH_DT>
void launch_elem_tet(ftb200_ctx* ctx, cudaStream_t s, int e0, int e1, int ignore) {
  if (e1 <= e0) return;
  const ElemArgs A = elem_args(ctx, e0, e1, ignore);
  const int grid = cdiv(e1 - e0, TET_BLOCK);
  if (WITH_FORCE && WITH_DT && ctx->injury && !ignore) LAUNCH((k_elem_tet<true, true, true>), grid, TET_BLOCK, s, A);
  else LAUNCH((k_elem_tet<WITH_FORCE, WITH_DT, false>), grid, TET_BLOCK, s, A);
}

// [e0, e1) of the internal order is cut into runs of equal (element type, material): every run gets the kernel
// specialised for its material (a uniform mesh is one run = one launch, as before)
template <bool WITH_FORCE, bool WITH_DT>
void launch_elem(ftb200_ctx* ctx, cudaStream_t s, int e0, int e1, int ignore) {
  if (e1 <= e0) return;
  for (const auto& r : ctx->ranges) {
    const int a = std::max(e0, r.e0), b = std::min(e1, r.e1);
    if (b <= a) continue;
    if (r.tet) launch_elem_tet<WITH_FORCE, WITH_DT>(ctx, s, a, b, ignore);
    else launch_elem_hex<WITH_FORCE, WITH_DT>(ctx, s, a, b, ignore, r.mat, r.affine);
  }
}

template <bool WITH_FORCE, bool WITH_DT>
void launch_elem_hex(ftb200_ctx* ctx, cudaStream_t s, int e0, int e1, int ignore, int mat, int affine) {
  if (e1 <= e0) return;
  const ElemArgs A = elem_args(ctx, e0, e1, ignore);
  const int grid = cdiv(e1 - e0, ELEM_BLOCK);
  if (!WITH_FORCE) { LAUNCH((k_elem<-1, false, true>), grid, ELEM_BLOCK, s, A); return; }
  if (WITH_FORCE && WITH_DT && affine && mat > 0) {  // a run of parallelepipeds of material 1, 4 or 5: k_elem_affine
    const bool inj = ctx->injury && !ignore;
    switch (mat) {
      case 1:
        if (inj) LAUNCH((k_elem_affine<1, true>), grid, ELEM_BLOCK, s, A);
        else if (ctx->use_nh && A.pk) LAUNCH((k_elem_affine_cj<1, true>), grid, ELEM_BLOCK, s, A);
        else if (ctx->use_nh) LAUNCH(k_elem_affine_cj<1>, grid, ELEM_BLOCK, s, A);
        else LAUNCH((k_elem_affine<1, false>), grid, ELEM_BLOCK, s, A);
        return;
      case 4:
        if (inj) LAUNCH((k_elem_affine<4, true>), grid, ELEM_BLOCK, s, A);
        else if (ctx->use_nh && A.pk) LAUNCH((k_elem_affine_cj<4, true>), grid, ELEM_BLOCK, s, A);
        else if (ctx->use_nh) LAUNCH(k_elem_affine_cj<4>, grid, ELEM_BLOCK, s, A);
        else LAUNCH((k_elem_affine<4, false>), grid, ELEM_BLOCK, s, A);
        return;
      case 5:
        if (inj) LAUNCH_HIST((k_elem_affine<5, true>), grid, ELEM_BLOCK, s, A);
        else LAUNCH_HIST((k_elem_affine<5, false>), grid, ELEM_BLOCK, s, A);
        return;
      default: break;
    }
  }
  if (WITH_FORCE && WITH_DT && ctx->injury && !ignore) {  // a step of the loop with the injury criteria on
    switch (mat) {
      case 1: LAUNCH((k_elem<1, true, true, true>), grid, ELEM_BLOCK, s, A); break;
      case 4: LAUNCH((k_elem<4, true, true, true>), grid, ELEM_BLOCK, s, A); break;
      case 5: LAUNCH_HIST((k_elem<5, true, true, true>), grid, ELEM_BLOCK, s, A); break;
      default: LAUNCH((k_elem<-1, true, true, true>), grid, ELEM_BLOCK, s, A); break;
    }
    return;
  }
  switch (mat) {
    case 1: LAUNCH((k_elem<1, WITH_FORCE, WITH_DT>), grid, ELEM_BLOCK, s, A); break;
    case 4: LAUNCH((k_elem<4, WITH_FORCE, WITH_DT>), grid, ELEM_BLOCK, s, A); break;
    case 5: LAUNCH_HIST((k_elem<5, WITH_FORCE, WITH_DT>), grid, ELEM_BLOCK, s, A); break;
    default: LAUNCH((k_elem<-1, WITH_FORCE, WITH_DT>), grid, ELEM_BLOCK, s, A); break;
  }
}

cudaEvent_t prof_event(ftb200_ctx* ctx, size_t* idx) {
  ProfEvents& P = ctx->prof;
  if (P.used == P.pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    P.pool.push_back(e);
  }
  *idx = P.used;
  return P.pool[P.used++];
}

// CalculateInjuryCriterions (ex5.cpp:1311-1430) across elements, after k_adv has advanced Time: running extrema,
// the two 95th-percentile selections (INJ_PASSES radix passes), the element lists.  INJ_LAUNCHES kernels.
constexpr int INJ_LAUNCHES = 2 + INJ_PASSES;
void launch_injury(ftb200_ctx* ctx, cudaStream_t s) {
  const ElemArgs A = elem_args(ctx, 0, ctx->nE, 0);
  LAUNCH(k_injury_reduce, std::min(INJ_BLOCKS, cdiv(ctx->nE, INJ_THREADS * 4)), INJ_THREADS, s, A, ctx->ref_of, ctx->inj_state, ctx->inj_part, ctx->inj_parti);
  double* h0 = ctx->inj_hist;
  double* h1 = ctx->inj_hist ? ctx->inj_hist + ctx->hist_cap : nullptr;
  if (ctx->nranks > 1) return;  // several partitions: the caller drives the passes (ftb200_injury_select_hist/_pick/_lists)
  for (int pass = 0; pass < INJ_PASSES; ++pass) LAUNCH(k_injury_select, dim3(std::min(INJ_BLOCKS, cdiv(ctx->nE, INJ_THREADS * INJ_ITEMS)), 2), INJ_THREADS, s, A, ctx->inj_state, pass, h0, h1, 0);
  LAUNCH(k_injury_lists, cdiv(ctx->nE, 256), 256, s, A, ctx->inj_state);
}

// one loop iteration: K_elem -> K_adv -> K_node (-> K_energy) (-> injury criteria)
void launch_step(ftb200_ctx* ctx, const double* recv) {
  cudaStream_t s = ctx->stream;
  size_t i0 = 0, i1 = 0;
  if (ctx->profile) cudaEventRecord(prof_event(ctx, &i0), s);
  launch_elem<true, true>(ctx, s, 0, ctx->nE, 0);
  if (ctx->profile) { cudaEventRecord(prof_event(ctx, &i1), s); ctx->prof.elem.push_back({i0, i1}); }
  const NodeArgs N = node_args(ctx, recv);
  const bool en_async = ctx->energy && ctx->energy_async && ctx->energy_async_now && ctx->nranks == 1 && !recv && !ctx->profile;
  // k_adv moves sc->step / sc->active, which the energy reduction of the previous step (helper stream) still reads
  if (ctx->energy_pending) { cudaStreamWaitEvent(s, ctx->ev_energy_done, 0); ctx->energy_pending = false; }
  LAUNCH((k_adv<false>), 1, 128, s, ctx->sc, ctx->mp, ctx->nPID, 0.0, ctx->dthist);
  if (ctx->rigid) LAUNCH(k_rigid_step, 1, 32, s, ctx->sc, ctx->rigid, 0);
  if (ctx->profile) cudaEventRecord(prof_event(ctx, &i0), s);
  if (ctx->energy) LAUNCH((k_node<true, true, true, true>), ctx->node_blocks, NODE_BLOCK, s, N);
  else LAUNCH((k_node<true, true, true, false>), ctx->node_blocks, NODE_BLOCK, s, N);
  if (ctx->profile) { cudaEventRecord(prof_event(ctx, &i1), s); ctx->prof.node.push_back({i0, i1}); }
  if (ctx->energy) {
    if (en_async) {  // K8 of this step overlaps K1 of the next one; joined before the next k_adv / at the end of the run
      cudaEventRecord(ctx->ev_nodes_done, s);
      cudaStreamWaitEvent(ctx->stream2, ctx->ev_nodes_done, 0);
      LAUNCH(k_energy, 1, 256, ctx->stream2, ctx->sc, ctx->epart, ctx->node_blocks, ctx->ehist);
      cudaEventRecord(ctx->ev_energy_done, ctx->stream2);
      ctx->energy_pending = true;
    } else {
      LAUNCH(k_energy, 1, 256, s, ctx->sc, ctx->epart, ctx->node_blocks, ctx->ehist);
    }
  }
  if (ctx->injury) launch_injury(ctx, s);
}

// ---------------------------------------------------------------------------------------------------------------
// Brick decomposition of an all-hexahedra mesh for k_brick / k_surf (ftb200_brick.cuh).  Elements are binned by their
// centroid into boxes of dims[] mean element extents (a structured or voxel mesh gives exact dims[0] x dims[1] x dims[2]
// bricks); a box with more than BRICK_NT elements, BRICK_NIMAX interior or BRICK_NSMAX surface nodes is bisected at the median of its longest
// axis until it fits.  Bricks are numbered box by box, x fastest: neighbours in the launch order share surface nodes in L2.
struct BrickPlan {
  int nB = 0;
  std::vector<int> eoff;    // [nB + 1] into elems
  std::vector<int> elems;   // caller element ids, ascending inside each brick
};

void brick_split(const ftb200_ctx* ctx, const std::vector<float>& cen, const std::vector<unsigned char>& deg, std::vector<int>& list,
                 size_t lo, size_t hi, std::vector<int>& stamp, std::vector<unsigned char>& cnt, int& stamp_id,
                 std::vector<std::pair<size_t, size_t>>& out) {
  const size_t n = hi - lo;
  bool fits = n <= (size_t)BRICK_NT;
  if (fits) {  // interior nodes (all their elements in the group) and surface nodes of the group
    ++stamp_id;
    int nl = 0, ni = 0;
    for (size_t i = lo; i < hi; ++i)
      for (int k = 0; k < 8; ++k) {
        const int nd = ctx->h_conn[8 * (size_t)list[i] + k];
        if (stamp[nd] != stamp_id) { stamp[nd] = stamp_id; cnt[nd] = 0; ++nl; }
        if (++cnt[nd] == deg[nd]) ++ni;
      }
    fits = ni <= BRICK_NIMAX && nl - ni <= BRICK_NSMAX;
  }
  if (fits || n <= 1) { out.push_back({lo, hi}); return; }
  float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
  for (size_t i = lo; i < hi; ++i)
    for (int c = 0; c < 3; ++c) {
      const float v = cen[3 * (size_t)list[i] + c];
      mn[c] = std::min(mn[c], v); mx[c] = std::max(mx[c], v);
    }
  int ax = 0;
  for (int c = 1; c < 3; ++c)
    if (mx[c] - mn[c] > mx[ax] - mn[ax]) ax = c;
  const size_t mid = lo + n / 2;
  std::nth_element(list.begin() + lo, list.begin() + mid, list.begin() + hi, [&](int a, int b) {
    const float va = cen[3 * (size_t)a + ax], vb = cen[3 * (size_t)b + ax];
    return va < vb || (va == vb && a < b);
  });
  brick_split(ctx, cen, deg, list, lo, mid, stamp, cnt, stamp_id, out);
  brick_split(ctx, cen, deg, list, mid, hi, stamp, cnt, stamp_id, out);
}

BrickPlan plan_bricks(const ftb200_ctx* ctx) {
  const int nE = ctx->nE, nN = ctx->nN;
  BrickPlan P;
  std::vector<float> cen(3 * (size_t)nE);
  double bbmin[3] = {1e300, 1e300, 1e300}, bbmax[3] = {-1e300, -1e300, -1e300}, hsum[3] = {0, 0, 0};
  for (int e = 0; e < nE; ++e) {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, cs[3] = {0, 0, 0};
    for (int k = 0; k < 8; ++k)
      for (int c = 0; c < 3; ++c) {
        const double x = ctx->h_X[3 * (size_t)ctx->h_conn[8 * (size_t)e + k] + c];
        lo[c] = std::min(lo[c], x); hi[c] = std::max(hi[c], x); cs[c] += x;
      }
    for (int c = 0; c < 3; ++c) {
      const double x = cs[c] / 8.0;
      cen[3 * (size_t)e + c] = (float)x;
      bbmin[c] = std::min(bbmin[c], x); bbmax[c] = std::max(bbmax[c], x);
      hsum[c] += hi[c] - lo[c];
    }
  }
  long long nc[3];
  double cell[3];
  for (int c = 0; c < 3; ++c) {
    const double h = std::max(hsum[c] / std::max(nE, 1), 1e-300);
    cell[c] = h * ctx->brick_dims[c];
    nc[c] = std::max<long long>(1, (long long)std::floor((bbmax[c] - bbmin[c]) / cell[c]) + 1);
  }
  // (box, caller id) order
  std::vector<std::pair<long long, int>> key(nE);
  for (int e = 0; e < nE; ++e) {
    long long ix[3];
    for (int c = 0; c < 3; ++c) {
      ix[c] = (long long)std::floor(((double)cen[3 * (size_t)e + c] - bbmin[c]) / cell[c] + 1e-6);
      ix[c] = std::min(std::max(ix[c], 0LL), nc[c] - 1);
    }
    key[e] = {ix[0] + nc[0] * (ix[1] + nc[1] * ix[2]), e};
  }
  std::sort(key.begin(), key.end());
  std::vector<int> list(nE);
  for (int i = 0; i < nE; ++i) list[i] = key[i].second;
  std::vector<int> stamp(nN, 0);
  std::vector<unsigned char> deg(nN, 0), cnt(nN, 0);
  for (size_t i = 0; i < 8 * (size_t)nE; ++i) ++deg[ctx->h_conn[i]];  // (at most 8 per node: checked by the caller)
  int stamp_id = 0;
  std::vector<std::pair<size_t, size_t>> groups;
  for (size_t i = 0; i < (size_t)nE;) {
    size_t j = i;
    while (j < (size_t)nE && key[j].first == key[i].first) ++j;
    brick_split(ctx, cen, deg, list, i, j, stamp, cnt, stamp_id, groups);
    i = j;
  }
  P.nB = (int)groups.size();
  P.eoff.assign(P.nB + 1, 0);
  P.elems.resize(nE);
  size_t w = 0;
  for (int b = 0; b < P.nB; ++b) {
    std::sort(list.begin() + groups[b].first, list.begin() + groups[b].second);
    for (size_t i = groups[b].first; i < groups[b].second; ++i) P.elems[w++] = list[i];
    P.eoff[b + 1] = (int)w;
  }
  return P;
}

// ---- brick-fused step: k_brick -> k_surf -> k_adv (-> k_energy).  The START of a step (first kick, drift, boundary
//      condition) is the prologue of k_brick / k_surf, so the state between two steps is the full-step (u, v, a) and a run
//      needs no separate START launch.
bool use_brick(const ftb200_ctx* ctx) { return ctx->brick_ok && !ctx->rigid && !ctx->injury && ctx->nranks == 1 && !ctx->p2p_ready; }

// energy partials of the brick-fused step: one per warp of k_brick, then one per warp of k_surf
int brick_eparts(const ftb200_ctx* c) { return c->nB * (BRICK_NT / 32) + c->surf_blocks * (SURF_BLOCK / 32); }

BrickArgs brick_args(ftb200_ctx* c) {
  BrickArgs A;
  A.hdr = c->b_hdr; A.conn16 = c->b_conn16; A.pe = c->b_pe; A.flags32 = c->b_flags32; A.halo = c->b_halo; A.map16 = c->b_map16;
  A.nB = c->nB;
  for (int k = 0; k < 3; ++k) {
    A.X[k] = c->X[k]; A.u[k] = c->u[k]; A.v[k] = c->v[k]; A.a[k] = c->a[k]; A.fi[k] = c->fi[k];
    A.fe[k] = c->has_fe ? c->fe[k] : nullptr;
    A.part[k] = c->b_part[k];
  }
  A.m = c->m; A.mp = c->mp;
  A.epart = c->epart; A.nEpart = brick_eparts(c); A.sc = c->sc; A.store_fi = c->energy ? 1 : 0;
  return A;
}
SurfArgs surf_args(ftb200_ctx* c) {
  SurfArgs A;
  A.ell = c->s_ell; A.ov_off = c->s_ovoff; A.ov_ent = c->s_ovent;
  for (int k = 0; k < 3; ++k) {
    A.u[k] = c->u[k]; A.v[k] = c->v[k]; A.a[k] = c->a[k]; A.fi[k] = c->fi[k];
    A.fe[k] = c->has_fe ? c->fe[k] : nullptr;
    A.part[k] = c->b_part[k];
  }
  A.m = c->m; A.flags = c->flags; A.epart = c->epart; A.nEpart = brick_eparts(c); A.eoff = c->nB * (BRICK_NT / 32);
  A.node0 = c->nIntTot; A.nS = c->nSurf; A.sc = c->sc; A.store_fi = c->energy ? 1 : 0;
  return A;
}
template <int MAT, bool EN>
void launch_brick_k(ftb200_ctx* ctx, cudaStream_t s, const BrickArgs& A) {
  auto kfn = k_brick<MAT, EN>;
  cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BRICK_SMEM_BYTES);  // per device
  launch_k(ctx, kfn, dim3(ctx->brick_grid), dim3(BRICK_NT), BRICK_SMEM_BYTES, s, A);
}
void launch_step_brick(ftb200_ctx* ctx) {
  cudaStream_t s = ctx->stream;
  size_t i0 = 0, i1 = 0;
  const BrickArgs A = brick_args(ctx);
  const int mat = ctx->ranges.empty() ? 1 : ctx->ranges[0].mat;
  if (ctx->profile) cudaEventRecord(prof_event(ctx, &i0), s);
  if (mat == 4) { if (ctx->energy) launch_brick_k<4, true>(ctx, s, A); else launch_brick_k<4, false>(ctx, s, A); }
  else { if (ctx->energy) launch_brick_k<1, true>(ctx, s, A); else launch_brick_k<1, false>(ctx, s, A); }
  if (ctx->profile) { cudaEventRecord(prof_event(ctx, &i1), s); ctx->prof.elem.push_back({i0, i1}); }
  if (ctx->profile) cudaEventRecord(prof_event(ctx, &i0), s);
  if (ctx->nSurf > 0) {
    const SurfArgs S = surf_args(ctx);
    if (ctx->energy) LAUNCH((k_surf<true>), ctx->surf_blocks, SURF_BLOCK, s, S);
    else LAUNCH((k_surf<false>), ctx->surf_blocks, SURF_BLOCK, s, S);
  }
  if (ctx->profile) { cudaEventRecord(prof_event(ctx, &i1), s); ctx->prof.node.push_back({i0, i1}); }
  const bool en_async = ctx->energy && ctx->energy_async && ctx->energy_async_now && !ctx->profile;
  if (ctx->energy_pending) { cudaStreamWaitEvent(s, ctx->ev_energy_done, 0); ctx->energy_pending = false; }
  LAUNCH((k_adv<false>), 1, 128, s, ctx->sc, ctx->mp, ctx->nPID, 0.0, ctx->dthist);
  if (ctx->energy) {
    const int nblocks = brick_eparts(ctx);
    if (en_async) {
      cudaEventRecord(ctx->ev_nodes_done, s);
      cudaStreamWaitEvent(ctx->stream2, ctx->ev_nodes_done, 0);
      LAUNCH(k_energy, 1, 256, ctx->stream2, ctx->sc, ctx->epart, nblocks, ctx->ehist);
      cudaEventRecord(ctx->ev_energy_done, ctx->stream2);
      ctx->energy_pending = true;
    } else {
      LAUNCH(k_energy, 1, 256, s, ctx->sc, ctx->epart, nblocks, ctx->ehist);
    }
  }
  ctx->felem_stale = true;
}

void prof_collect(ftb200_ctx* ctx) {
  ProfEvents& P = ctx->prof;
  for (auto& pr : P.elem) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, P.pool[pr.first], P.pool[pr.second]) == cudaSuccess) { ctx->prof_elem_ms += ms; ctx->prof_elem_n++; }
  }
  for (auto& pr : P.node) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, P.pool[pr.first], P.pool[pr.second]) == cudaSuccess) { ctx->prof_node_ms += ms; ctx->prof_node_n++; }
  }
  P.elem.clear(); P.node.clear(); P.used = 0;
}

int upload_aos(ftb200_ctx* ctx, const double* host, double* const dst[3]) {
  const size_t n3 = 3 * (size_t)ctx->nN;
  CK(cudaMemcpyAsync(ctx->d_stage[0], host, n3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(k_aos_to_soa, cdiv(ctx->nNp, 256), 256, ctx->stream, ctx->d_stage[0], dst[0], dst[1], dst[2], ctx->d_nref, ctx->nNp);
  return 0;
}
int download_aos(ftb200_ctx* ctx, double* const src[3], double* host, int stage) {
  const size_t n3 = 3 * (size_t)ctx->nN;
  LAUNCH(k_soa_to_aos, cdiv(ctx->nNp, 256), 256, ctx->stream, src[0], src[1], src[2], ctx->d_stage[stage], ctx->d_nref, ctx->nNp);
  CK(cudaMemcpyAsync(host, ctx->d_stage[stage], n3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  return 0;
}
int upload_boundary(ftb200_ctx* ctx, const int* boundary) {
  const size_t n3 = 3 * (size_t)ctx->nN;
  CK(cudaMemcpyAsync(ctx->d_istage, boundary, n3 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(k_boundary_to_flags, cdiv(ctx->nNp, 256), 256, ctx->stream, ctx->d_istage, ctx->flags, ctx->d_nref, ctx->nNp);
  return 0;
}

void free_all(ftb200_ctx* c) {
  for (int k = 0; k < 3; ++k) {
    dfree(c->X[k]); dfree(c->u[k]); dfree(c->v[k]); dfree(c->a[k]); dfree(c->fi[k]); dfree(c->du[k]);
    dfree(c->fnet[k]); dfree(c->fe[k]); dfree(c->d_stage[k]);
  }
  dfree(c->m); dfree(c->flags); dfree(c->conn); dfree(c->pid); dfree(c->ref_of); dfree(c->eflag);
  dfree(c->felem); dfree(c->hist); dfree(c->mp); dfree(c->node_off); dfree(c->node_ent); dfree(c->sc);
  dfree(c->etype); dfree(c->gpoff);
  dfree(c->rigid); dfree(c->rigid_tab); dfree(c->aprev[0]); dfree(c->aprev[1]); dfree(c->aprev[2]);
  dfree(c->inj_ps); dfree(c->inj_psxsr); dfree(c->inj_smin); dfree(c->inj_shear); dfree(c->inj_part); dfree(c->inj_hist);
  dfree(c->inj_flags); dfree(c->inj_incl); dfree(c->inj_parti); dfree(c->inj_state);
  dfree(c->dthist); dfree(c->ehist); dfree(c->epart); dfree(c->out3); dfree(c->d_istage); dfree(c->d_big);
  c->d_big_bytes = 0;
  dfree(c->d_detmin); dfree(c->d_nonpos);
  for (void* q : c->p2p_opened) cudaIpcCloseMemHandle(q);
  c->p2p_opened.clear();
  dfree(c->p2p_window); dfree(c->d_seq); dfree(c->d_p2p_blocks);
  c->p2p_ready = false;
  if (c->p2p_graph) { cudaGraphExecDestroy(c->p2p_graph); c->p2p_graph = nullptr; }
  dfree(c->d_nref); dfree(c->d_nint); dfree(c->d_ell);
  dfree(c->b_hdr); dfree(c->b_conn16); dfree(c->b_map16); dfree(c->b_halo); dfree(c->b_pe); dfree(c->b_flags32); dfree(c->s_ell); dfree(c->s_ovoff); dfree(c->s_ovent);
  dfree(c->b_part[0]); dfree(c->b_part[1]); dfree(c->b_part[2]);
  c->brick_ok = false;
  if (c->graph) { cudaGraphExecDestroy(c->graph); c->graph = nullptr; }
  dfree(c->d_xsend); dfree(c->d_xrecv);
  dfree(c->d_pk); dfree(c->d_pk_ctr);
  dfree(c->d_sendNodeIndex); dfree(c->halo_nodes); dfree(c->halo_off); dfree(c->halo_slot); dfree(c->halo_node_idx);
  if (c->graph) { cudaGraphExecDestroy(c->graph); c->graph = nullptr; }
}

}  // namespace

extern "C" {

const char* ftb200_build_info(void) {
  return "femtech_b200 C-ABI; CUDA kernels built for sm_100a (fp64, one thread per element / node)";
}

int ftb200_create(int rank, int nranks, int device, ftb200_ctx** out) {
  if (!out) return FTB200_ERR_INPUT;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    fprintf(stderr, "ftb200_create: no usable CUDA device (%s); there is no CPU fallback\n", cudaGetErrorString(e));
    return FTB200_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) return FTB200_ERR_INPUT;
  ftb200_ctx* ctx = new ftb200_ctx();
  ctx->rank = rank; ctx->nranks = nranks; ctx->device = device;
  int prio_lo = 0, prio_hi = 0;
  if (cudaSetDevice(device) != cudaSuccess || cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) != cudaSuccess ||
      cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
      cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
      cudaStreamCreateWithPriority(&ctx->stream_lo, cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_nodes_done, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_energy_done, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_prog, cudaEventDisableTiming) != cudaSuccess) {
    delete ctx;
    return FTB200_ERR_CUDA;
  }
  ctx->prio_hi = prio_hi; ctx->prio_lo = prio_lo;
  if (const char* ev = getenv("FTB200_P2P_ORDER")) ctx->p2p_order = atoi(ev);
  *out = ctx;
  return FTB200_OK;
}

int ftb200_destroy(ftb200_ctx* ctx) {
  if (!ctx) return FTB200_OK;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  if (ctx->trace) {  // diagnostic dump of the partitioned step's time stamps
    std::vector<unsigned long long> h((size_t)TRACE_STEPS * TRACE_SLOTS);
    cudaMemcpy(h.data(), ctx->trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    char path[512];
    snprintf(path, sizeof path, "%s_rank%d.txt", ctx->trace_prefix.c_str(), ctx->rank);
    if (FILE* f = fopen(path, "w")) {
      for (int st = 0; st < TRACE_STEPS; ++st) {
        const unsigned long long* r = &h[(size_t)st * TRACE_SLOTS];
        if (!r[0]) continue;
        fprintf(f, "%d", st);
        for (int k = 0; k < TRACE_SLOTS; ++k) fprintf(f, " %llu", r[k]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
    dfree(ctx->trace);
    ctx->trace = nullptr;
  }
  free_all(ctx);
  for (auto e : ctx->prof.pool) cudaEventDestroy(e);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  if (ctx->stream_lo) cudaStreamDestroy(ctx->stream_lo);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->ev_nodes_done) cudaEventDestroy(ctx->ev_nodes_done);
  if (ctx->ev_energy_done) cudaEventDestroy(ctx->ev_energy_done);
  if (ctx->ev_prog) cudaEventDestroy(ctx->ev_prog);
  if (ctx->ring_host) cudaFreeHost(ctx->ring_host);
  delete ctx;
  return FTB200_OK;
}

const char* ftb200_last_error(const ftb200_ctx* ctx) { return ctx ? ctx->err : "null context"; }
long long ftb200_launch_count(const ftb200_ctx* ctx) { return ctx ? ctx->launches : 0; }

int ftb200_set_stream(ftb200_ctx* ctx, void* cuda_stream) {
  if (!ctx) return FTB200_ERR_INPUT;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->graph) { cudaGraphExecDestroy(ctx->graph); ctx->graph = nullptr; }
  if (cuda_stream) {
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
  } else if (!ctx->own_stream) {
    int plo = 0, phi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&plo, &phi));
    CK(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, phi));
    ctx->own_stream = true;
  }
  return FTB200_OK;
}

int ftb200_upload_mesh(ftb200_ctx* ctx, const double* coordinates, const int* connectivity, const int* pid, int nNodes,
                       int nElements) {
  if (!ctx || !coordinates || !connectivity || !pid || nNodes <= 0 || nElements <= 0)
    return fail(ctx, FTB200_ERR_INPUT, "upload_mesh: null pointer or empty mesh");
  if ((long long)nElements * 8 >= (1LL << 31)) return fail(ctx, FTB200_ERR_INPUT, "upload_mesh: more than 2^28 elements per GPU");
  for (long long i = 0; i < 8LL * nElements; ++i)
    if (connectivity[i] < 0 || connectivity[i] >= nNodes)
      return fail(ctx, FTB200_ERR_INPUT, "upload_mesh: connectivity[%lld] = %d out of range", i, connectivity[i]);
  ctx->nN = nNodes; ctx->nE = nElements;
  ctx->h_X.assign(coordinates, coordinates + 3 * (size_t)nNodes);
  ctx->h_conn.assign(connectivity, connectivity + 8 * (size_t)nElements);
  ctx->h_pid.assign(pid, pid + nElements);
  ctx->h_etype.clear(); ctx->has_tet = false;
  ctx->mesh_ok = true; ctx->shape_ok = false; ctx->begun = false;
  return FTB200_OK;
}

long long ftb200_affine_element_count(ftb200_ctx* ctx) { return (ctx && ctx->shape_ok) ? ctx->nE_affine : -1; }
int ftb200_brick_info(ftb200_ctx* ctx, long long* out8) {
  if (!ctx || !ctx->shape_ok || !out8) return fail(ctx, FTB200_ERR_INPUT, "brick_info: call shape_functions first");
  out8[0] = ctx->brick_ok ? ctx->nB : 0; out8[1] = ctx->brick_ok ? ctx->nIntTot : 0; out8[2] = ctx->brick_ok ? ctx->nSurf : 0;
  out8[3] = ctx->brick_ok ? ctx->nSlots : 0;
  for (int c = 0; c < 3; ++c) out8[4 + c] = ctx->brick_dims[c];
  out8[7] = use_brick(ctx) ? 1 : 0;
  return FTB200_OK;
}
int ftb200_brick_maps(ftb200_ctx* ctx, int* brick_of_element, int* interior_brick_of_node) {
  if (!ctx || !ctx->shape_ok || !ctx->brick_ok) return fail(ctx, FTB200_ERR_INPUT, "brick_maps: no brick decomposition");
  CK(cudaSetDevice(ctx->device));
  std::vector<BrickHdr> hdr(ctx->nB);
  std::vector<int> ref_of(ctx->nE);
  CK(ftb_memcpy(ctx, hdr.data(), ctx->b_hdr, hdr.size() * sizeof(BrickHdr), cudaMemcpyDeviceToHost));
  CK(ftb_memcpy(ctx, ref_of.data(), ctx->ref_of, ref_of.size() * sizeof(int), cudaMemcpyDeviceToHost));
  std::vector<int> bint(ctx->nNp, -1);
  for (int b = 0; b < ctx->nB; ++b) {
    if (brick_of_element)
      for (int i = 0; i < hdr[b].nEl; ++i) brick_of_element[ref_of[hdr[b].e0 + i]] = b;
    for (int i = 0; i < hdr[b].nInt; ++i) bint[hdr[b].ibase + i] = b;
  }
  if (interior_brick_of_node)
    for (int n = 0; n < ctx->nN; ++n) interior_brick_of_node[n] = bint[ctx->h_nint[n]];
  return FTB200_OK;
}
long long ftb200_gauss_point_count(ftb200_ctx* ctx) { return ctx ? (ctx->shape_ok ? ctx->nGP : (ctx->has_tet ? -1 : 8LL * ctx->nE)) : -1; }

int ftb200_upload_mesh_mixed(ftb200_ctx* ctx, const double* coordinates, const int* connectivity, const int* eptr, const int* pid,
                             int nNodes, int nElements) {
  if (!ctx || !coordinates || !connectivity || !eptr || !pid || nNodes <= 0 || nElements <= 0)
    return fail(ctx, FTB200_ERR_INPUT, "upload_mesh_mixed: null pointer or empty mesh");
  if ((long long)nElements * 8 >= (1LL << 31)) return fail(ctx, FTB200_ERR_INPUT, "upload_mesh_mixed: more than 2^28 elements per GPU");
  std::vector<int> conn8(8 * (size_t)nElements);
  std::vector<uint8_t> et(nElements);
  bool any = false;
  for (int e = 0; e < nElements; ++e) {
    const int n = eptr[e + 1] - eptr[e];
    if (n != 8 && n != 4)  // the reference terminates on other element types in 3-D (ShapeFunctions.cpp, code 3)
      return fail(ctx, FTB200_ERR_INPUT, "upload_mesh_mixed: element %d has %d nodes; C3D8 and C3D4 are supported", e, n);
    for (int k = 0; k < 8; ++k) {
      const int v = connectivity[eptr[e] + (k < n ? k : 0)];  // unused slots of a tetrahedron repeat its first node
      if (v < 0 || v >= nNodes) return fail(ctx, FTB200_ERR_INPUT, "upload_mesh_mixed: node %d of element %d out of range", v, e);
      conn8[8 * (size_t)e + k] = v;
    }
    et[e] = n == 4;
    any = any || n == 4;
  }
  int rc = ftb200_upload_mesh(ctx, coordinates, conn8.data(), pid, nNodes, nElements);
  if (rc) return rc;
  if (any) { ctx->h_etype = et; ctx->has_tet = true; }
  return FTB200_OK;
}

int ftb200_upload_materials(ftb200_ctx* ctx, const int* materialID, const double* properties, int nPID) {
  if (!ctx || !materialID || !properties || nPID <= 0) return fail(ctx, FTB200_ERR_INPUT, "upload_materials: bad arguments");
  for (int p = 0; p < nPID; ++p)
    if (materialID[p] < 0 || materialID[p] > 5)  // StressUpdate.cpp:24-26
      return fail(ctx, FTB200_ERR_MATERIAL, "Unknown material type %d for part %d", materialID[p], p);
  ctx->nPID = nPID;
  ctx->h_matid.assign(materialID, materialID + nPID);
  ctx->h_props.assign(properties, properties + 9 * (size_t)nPID);
  ctx->mat_ok = true; ctx->shape_ok = false;
  return FTB200_OK;
}

int ftb200_upload_comm(ftb200_ctx* ctx, int sendProcessCount, const int* sendProcessID, const int* sendNeighbourCountCum,
                       const int* sendNodeIndex) {
  if (!ctx || sendProcessCount < 0) return fail(ctx, FTB200_ERR_INPUT, "upload_comm: bad arguments");
  ctx->h_sendProcessID.clear(); ctx->h_sendCum.assign(1, 0); ctx->h_sendNodeIndex.clear();
  if (sendProcessCount > 0) {
    if (!sendProcessID || !sendNeighbourCountCum || !sendNodeIndex) return fail(ctx, FTB200_ERR_INPUT, "upload_comm: null pointer");
    ctx->h_sendProcessID.assign(sendProcessID, sendProcessID + sendProcessCount);
    ctx->h_sendCum.assign(sendNeighbourCountCum, sendNeighbourCountCum + sendProcessCount + 1);
    ctx->h_sendNodeIndex.assign(sendNodeIndex, sendNodeIndex + sendNeighbourCountCum[sendProcessCount]);
  }
  ctx->shape_ok = false;
  return FTB200_OK;
}

int ftb200_shape_functions(ftb200_ctx* ctx, double* min_detJ) {
  if (!ctx || !ctx->mesh_ok || !ctx->mat_ok) return fail(ctx, FTB200_ERR_INPUT, "shape_functions: upload mesh and materials first");
  CK(cudaSetDevice(ctx->device));
  const int nN = ctx->nN, nE = ctx->nE, nPID = ctx->nPID;
  for (int e = 0; e < nE; ++e)
    if (ctx->h_pid[e] < 0 || ctx->h_pid[e] >= nPID) return fail(ctx, FTB200_ERR_INPUT, "pid[%d] = %d out of range", e, ctx->h_pid[e]);
  free_all(ctx);
  // ---- internal element order: elements touching a shared node first (their forces feed the halo),
  //      the caller's order kept inside each group --------------------------------------------------
  ctx->halo_count = (int)ctx->h_sendNodeIndex.size();
  std::vector<char> is_shared(nN, 0);
  for (int i = 0; i < ctx->halo_count; ++i) {
    const int n = ctx->h_sendNodeIndex[i];
    if (n < 0 || n >= nN) return fail(ctx, FTB200_ERR_INPUT, "sendNodeIndex[%d] = %d out of range", i, n);
    is_shared[n] = 1;
  }
  std::vector<int> ref_of(nE), int_of(nE);
  {
    std::vector<char> isb(nE, 0);
    int nb = 0;
    if (ctx->halo_count)
      for (int e = 0; e < nE; ++e) {
        for (int k = 0; k < 8; ++k)
          if (is_shared[ctx->h_conn[8 * (size_t)e + k]]) { isb[e] = 1; break; }
        nb += isb[e];
      }
    ctx->nE_boundary = nb;
    // inside each class: hexahedra first, tetrahedra after them (their own kernel); inside each type by material id, so
    // that every run of the internal order is uniform and gets the kernel specialised for it; caller's order kept
    const bool mixed = ctx->has_tet;
    // hexahedra of materials 1, 4, 5 whose reference geometry is a parallelepiped (hex8_is_affine, exact test) form their
    // own runs behind the other elements of the same material: they are integrated by k_elem_affine
    std::vector<char> aff(nE, 0);
    ctx->nE_affine = 0;
    if (const char* ev = getenv("FTB200_AFFINE")) ctx->use_affine = atoi(ev) != 0;
    if (const char* ev = getenv("FTB200_NH")) ctx->use_nh = atoi(ev) != 0;
    if (ctx->use_affine)
      for (int e = 0; e < nE; ++e) {
        if (mixed && ctx->h_etype[e]) continue;
        const int mid = ctx->h_matid[ctx->h_pid[e]];
        if (mid != 1 && mid != 4 && mid != 5) continue;
        double Xe[8][3];
        for (int k = 0; k < 8; ++k)
          for (int c = 0; c < 3; ++c) Xe[k][c] = ctx->h_X[3 * (size_t)ctx->h_conn[8 * (size_t)e + k] + c];
        aff[e] = hex8_is_affine(Xe) ? 1 : 0;
        ctx->nE_affine += aff[e];
      }
    constexpr int NM = 12;  // (material id 0..5, StressUpdate.cpp:7-27) x (general, affine)
    auto bucket = [&](int e) {
      const int mid = ctx->h_matid[ctx->h_pid[e]];
      const int mslot = ((mid >= 0 && mid < 6) ? mid : 0) * 2 + aff[e];  // unknown ids go with the generic kernel, which reports them
      return ((isb[e] ? 0 : 1) * 2 + ((mixed && ctx->h_etype[e]) ? 1 : 0)) * NM + mslot;
    };
    std::vector<int> count(4 * NM, 0), start(4 * NM + 1, 0);
    for (int e = 0; e < nE; ++e) count[bucket(e)]++;
    for (int b = 0; b < 4 * NM; ++b) start[b + 1] = start[b] + count[b];
    ctx->nEb_hex = start[NM]; ctx->nEi_hex = start[3 * NM] - start[2 * NM];
    ctx->ranges.clear();
    for (int b = 0; b < 4 * NM; ++b)
      if (count[b]) {
        const int tet = (b / NM) & 1, mid = (b % NM) >> 1, af = (b % NM) & 1;
        // neighbouring buckets of the same type that would run the same kernel are merged (materials 0, 2, 3 -> generic)
        const int kmat = (mid == 1 || mid == 4 || mid == 5) ? mid : -1;
        if (!ctx->ranges.empty() && ctx->ranges.back().e1 == start[b] && ctx->ranges.back().tet == tet &&
            (tet || (ctx->ranges.back().mat == kmat && ctx->ranges.back().affine == af)) && (start[b] != nb))
          ctx->ranges.back().e1 = start[b + 1];
        else
          ctx->ranges.push_back({start[b], start[b + 1], tet, kmat, af});
      }
    std::vector<int> cur(start.begin(), start.end() - 1);
    for (int e = 0; e < nE; ++e) {
      const int t = cur[bucket(e)]++;
      ref_of[t] = e; int_of[e] = t;
    }
  }
  // ---- brick mode (ftb200_brick.cuh): a single-partition mesh of parallelepiped hexahedra of one material (1 or 4)
  //      whose nodes have at most 8 elements.  Elements are renumbered brick by brick (the whole mesh is one run of
  //      the internal order, so any order inside it is allowed), nodes as [interior nodes by brick | surface nodes by
  //      owning brick].  Every other mesh keeps the two-kernel step.
  BrickPlan plan;
  bool brick = false;
  {
    if (const char* ev = getenv("FTB200_BRICK")) ctx->brick_want = atoi(ev) != 0;
    if (const char* ev = getenv("FTB200_BRICK_DIMS")) {
      int d[3];
      if (sscanf(ev, "%d,%d,%d", &d[0], &d[1], &d[2]) == 3 && d[0] > 0 && d[1] > 0 && d[2] > 0 && (long long)d[0] * d[1] * d[2] <= BRICK_NT)
        for (int c = 0; c < 3; ++c) ctx->brick_dims[c] = d[c];
    }
    bool ok = ctx->brick_want && ctx->nranks == 1 && ctx->halo_count == 0 && !ctx->has_tet && nE > 0 && ctx->nE_affine == nE;
    if (ok) {
      const int m0 = ctx->h_matid[ctx->h_pid[0]];
      ok = (m0 == 1 || m0 == 4);
      for (int e = 0; ok && e < nE; ++e) ok = ctx->h_matid[ctx->h_pid[e]] == m0;
    }
    if (ok) {
      std::vector<unsigned char> deg(nN, 0);
      for (size_t i = 0; ok && i < 8 * (size_t)nE; ++i) ok = ++deg[ctx->h_conn[i]] <= 8;
    }
    if (ok) {
      plan = plan_bricks(ctx);
      brick = true;
      for (int t = 0; t < nE; ++t) { ref_of[t] = plan.elems[t]; int_of[plan.elems[t]] = t; }
    }
  }
  // ---- internal node order: the caller's order (brick mode: interior nodes brick by brick, then the surface nodes
  //      by owning brick), padded to whole node tiles ------------------------------------------------------------
  const int nNp = std::max(cdiv(nN, NODE_TILE) * NODE_TILE, NODE_TILE);
  ctx->nNp = nNp;
  std::vector<int> nref(nNp, -1), nint(nN, -1);
  std::vector<int> b_ibase, b_nint;
  int nIntTot = 0;
  if (brick) {
    std::vector<int> first(nN, -1);
    std::vector<char> multi(nN, 0);
    for (int b = 0; b < plan.nB; ++b)  // ascending brick id: first[] is the lowest brick of a node = its owner
      for (int i = plan.eoff[b]; i < plan.eoff[b + 1]; ++i)
        for (int k = 0; k < 8; ++k) {
          const int nd = ctx->h_conn[8 * (size_t)plan.elems[i] + k];
          if (first[nd] < 0) first[nd] = b;
          else if (first[nd] != b) multi[nd] = 1;
        }
    b_ibase.assign(plan.nB + 1, 0); b_nint.assign(plan.nB, 0);
    std::vector<int> scount(plan.nB + 1, 0);
    for (int n = 0; n < nN; ++n)
      if (first[n] >= 0) { if (multi[n]) scount[first[n]]++; else b_nint[first[n]]++; }
    for (int b = 0; b < plan.nB; ++b) b_ibase[b + 1] = b_ibase[b] + b_nint[b];
    nIntTot = b_ibase[plan.nB];
    std::vector<int> curI(b_ibase.begin(), b_ibase.end() - 1), curS(plan.nB + 1, 0);
    { int acc = nIntTot; for (int b = 0; b < plan.nB; ++b) { curS[b] = acc; acc += scount[b]; } curS[plan.nB] = acc; }
    int orphan = curS[plan.nB];
    for (int n = 0; n < nN; ++n) {
      if (first[n] < 0) nint[n] = orphan++;          // node without elements: a surface node with no partial
      else if (multi[n]) nint[n] = curS[first[n]]++;
      else nint[n] = curI[first[n]]++;
    }
    for (int n = 0; n < nN; ++n) nref[nint[n]] = n;
  } else {
    for (int n = 0; n < nN; ++n) { nref[n] = n; nint[n] = n; }
  }
  ctx->h_nint = nint;
  // ---- shared nodes (internal ids), slots in ascending neighbour order ---------------------------
  std::vector<int> node_h(nNp, -1), halo_nodes, sendIdxInt(ctx->halo_count);
  for (int i = 0; i < ctx->halo_count; ++i) sendIdxInt[i] = nint[ctx->h_sendNodeIndex[i]];
  for (int i = 0; i < nNp; ++i)
    if (nref[i] >= 0 && is_shared[nref[i]]) { node_h[i] = (int)halo_nodes.size(); halo_nodes.push_back(i); }
  ctx->nshared = (int)halo_nodes.size();
  std::vector<int> hoff(ctx->nshared + 1, 0), hslot(ctx->halo_count);
  for (int i = 0; i < ctx->halo_count; ++i) hoff[node_h[sendIdxInt[i]] + 1]++;
  for (int h = 0; h < ctx->nshared; ++h) hoff[h + 1] += hoff[h];
  {
    std::vector<int> cur(hoff.begin(), hoff.end() - 1);
    for (int i = 0; i < ctx->halo_count; ++i) hslot[cur[node_h[sendIdxInt[i]]]++] = i;  // slot index is neighbour-major
  }
  // ---- CSR node -> (element, slot), ascending CALLER element id (GetForce_3D.cpp:15,39-44) ---------
  std::vector<int> off(nNp + 1, 0), ent(8 * (size_t)nE);
  auto nen = [&](int e) { return (ctx->has_tet && ctx->h_etype[e]) ? 4 : 8; };
  for (int e = 0; e < nE; ++e)
    for (int k = 0; k < nen(e); ++k) off[nint[ctx->h_conn[8 * (size_t)e + k]] + 1]++;
  for (int n = 0; n < nNp; ++n) off[n + 1] += off[n];
  {
    std::vector<int> cur(off.begin(), off.end() - 1);
    for (int e = 0; e < nE; ++e)
      for (int k = 0; k < nen(e); ++k) ent[cur[nint[ctx->h_conn[8 * (size_t)e + k]]]++] = int_of[e] * 8 + k;
  }
  // ---- SoA planes ----------------------------------------------------------------------------
  std::vector<int> connT(8 * (size_t)nE), pidI(nE);
  for (int t = 0; t < nE; ++t) {
    const int e = ref_of[t];
    for (int k = 0; k < 8; ++k) connT[(size_t)k * nE + t] = nint[ctx->h_conn[8 * (size_t)e + k]];
    pidI[t] = ctx->h_pid[e];
  }
  std::vector<double> Xs(3 * (size_t)nNp, 0.0);
  for (int i = 0; i < nNp; ++i)
    if (nref[i] >= 0)
      for (int c = 0; c < 3; ++c) Xs[(size_t)c * nNp + i] = ctx->h_X[3 * (size_t)nref[i] + c];
  // per-part parameter blocks
  std::vector<double> mp((size_t)nPID * FTB_MP_STRIDE, 0.0);
  std::vector<char> used(nPID, 0);
  for (int e = 0; e < nE; ++e) used[ctx->h_pid[e]] = 1;
  ctx->uniform_mat = -2; ctx->has_visco = false;
  for (int p = 0; p < nPID; ++p) {
    double* q = &mp[(size_t)p * FTB_MP_STRIDE];
    for (int k = 0; k < 9; ++k) q[k] = ctx->h_props[9 * (size_t)p + k];
    const double rho = q[MP_RHO], mu = q[MP_MU], lambda = q[MP_LAMBDA];
    const double nu = 0.5 * lambda / (lambda + mu);          // CalculateTimeStep.cpp:15
    q[MP_CE] = sqrt(lambda * (1.0 / nu - 1.0) / rho);       // :17
    q[MP_KBULK] = lambda + 2.0 * mu / 3.0;                   // HGOIsotropic.cpp:44
    q[MP_MATID] = (double)ctx->h_matid[p];
    if (used[p]) {
      if (ctx->h_matid[p] == 5) ctx->has_visco = true;
      if (ctx->uniform_mat == -2) ctx->uniform_mat = ctx->h_matid[p];
      else if (ctx->uniform_mat != ctx->h_matid[p]) ctx->uniform_mat = -1;
    }
  }
  if (ctx->uniform_mat != 1 && ctx->uniform_mat != 4 && ctx->uniform_mat != 5) ctx->uniform_mat = -1;
  // ---- device allocation ---------------------------------------------------------------------
  int rc;
  for (int k = 0; k < 3; ++k) {
    if ((rc = dalloc(ctx, &ctx->X[k], nNp)) || (rc = dalloc(ctx, &ctx->u[k], nNp)) || (rc = dalloc(ctx, &ctx->v[k], nNp)) ||
        (rc = dalloc(ctx, &ctx->a[k], nNp)) || (rc = dalloc(ctx, &ctx->fi[k], nNp)) || (rc = dalloc(ctx, &ctx->du[k], nNp)) ||
        (rc = dalloc(ctx, &ctx->fnet[k], nNp)) || (rc = dalloc(ctx, &ctx->d_stage[k], 3 * (size_t)std::max(nN, nNp))))
      return rc;
    CK(ftb_memcpy(ctx, ctx->X[k], &Xs[(size_t)k * nNp], nNp * sizeof(double), cudaMemcpyHostToDevice));
    CK(ftb_memset(ctx, ctx->u[k], 0, nNp * sizeof(double)));
    CK(ftb_memset(ctx, ctx->v[k], 0, nNp * sizeof(double)));
    CK(ftb_memset(ctx, ctx->a[k], 0, nNp * sizeof(double)));
    CK(ftb_memset(ctx, ctx->fi[k], 0, nNp * sizeof(double)));
    CK(ftb_memset(ctx, ctx->du[k], 0, nNp * sizeof(double)));
    CK(ftb_memset(ctx, ctx->fnet[k], 0, nNp * sizeof(double)));
  }
  ctx->node_blocks = cdiv(nNp, NODE_BLOCK);
  int epart_blocks = std::max(ctx->node_blocks, cdiv(nN, NODE_BLOCK));
  if (brick) {
    // per-brick local numbering: interior nodes first (internal id - ibase), then the surface nodes the brick touches
    // in ascending internal id; connectivity in local ids, local node -> (element, slot) map in ascending reference
    // element id, one partial slot per (brick, surface local node); per surface node the slots in ascending brick id
    const int nB = plan.nB;
    const int nS = nN - nIntTot;
    std::vector<BrickHdr> hdr(nB);
    std::vector<uint16_t> conn16((size_t)nB * 8 * BRICK_NT, 0), map16((size_t)nB * 8 * BRICK_NLMAX, 0xFFFFu);
    std::vector<int> halo((size_t)nB * BRICK_NSMAX, 0), sell(8 * (size_t)std::max(nS, 1), -1), scnt(std::max(nS, 1), 0);
    std::vector<std::pair<int, int>> sover;
    std::vector<int> loc(nNp, -1), surf;
    std::vector<unsigned char> lcnt(BRICK_NLMAX);
    long long slot = 0;
    for (int b = 0; b < nB; ++b) {
      const int e0 = plan.eoff[b], nEl = plan.eoff[b + 1] - e0;
      surf.clear();
      for (int i = 0; i < nEl; ++i)
        for (int k = 0; k < 8; ++k) {
          const int g = nint[ctx->h_conn[8 * (size_t)plan.elems[e0 + i] + k]];
          if (g >= nIntTot && loc[g] != -2 - b) { loc[g] = -2 - b; surf.push_back(g); }
        }
      std::sort(surf.begin(), surf.end());
      const int nInt_b = b_nint[b], nLoc = nInt_b + (int)surf.size();
      if (nEl > BRICK_NT || nInt_b > BRICK_NIMAX || (int)surf.size() > BRICK_NSMAX) return fail(ctx, FTB200_ERR_INPUT, "brick %d does not fit (%d elements, %d nodes)", b, nEl, nLoc);
      for (size_t i = 0; i < surf.size(); ++i) { loc[surf[i]] = nInt_b + (int)i; halo[(size_t)b * BRICK_NSMAX + i] = surf[i]; }
      std::fill(lcnt.begin(), lcnt.end(), 0);
      for (int i = 0; i < nEl; ++i)
        for (int k = 0; k < 8; ++k) {
          const int g = nint[ctx->h_conn[8 * (size_t)plan.elems[e0 + i] + k]];
          const int l = g < nIntTot ? g - b_ibase[b] : loc[g];
          conn16[((size_t)b * BRICK_NT + i) * 8 + k] = (uint16_t)l;
          map16[((size_t)b * 8 + lcnt[l]++) * BRICK_NLMAX + l] = (uint16_t)(k * BRICK_NT + i);
        }
      hdr[b] = BrickHdr{e0, nEl, b_ibase[b], nInt_b, nLoc, (int)slot, 0, 0};
      for (size_t i = 0; i < surf.size(); ++i) {
        const int sidx = surf[i] - nIntTot;
        const int q = scnt[sidx]++;
        if (q < 8) sell[(size_t)q * nS + sidx] = (int)(slot + (long long)i);
        else sover.push_back({sidx, (int)(slot + (long long)i)});
      }
      slot += (long long)surf.size();
      if (slot > 0x7fffff00LL) return fail(ctx, FTB200_ERR_INPUT, "too many surface partials for 32-bit slots");
    }
    if ((rc = dalloc(ctx, &ctx->b_pe, nE)) || (rc = dalloc(ctx, &ctx->b_flags32, nNp))) return rc;
    {
      int sms = 0;
      CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
      ctx->brick_grid = std::min(nB, 2 * std::max(sms, 1));  // persistent blocks, two per SM
    }
    ctx->nB = nB; ctx->nIntTot = nIntTot; ctx->nSurf = nS; ctx->nSlots = slot;
    ctx->surf_blocks = cdiv(nS, SURF_BLOCK);
    epart_blocks = std::max(epart_blocks, nB * (BRICK_NT / 32) + ctx->surf_blocks * (SURF_BLOCK / 32));
    if ((rc = dalloc(ctx, &ctx->b_hdr, nB)) || (rc = dalloc(ctx, &ctx->b_conn16, conn16.size())) ||
        (rc = dalloc(ctx, &ctx->b_map16, map16.size())) || (rc = dalloc(ctx, &ctx->b_halo, halo.size())) ||
        (rc = dalloc(ctx, &ctx->s_ell, sell.size())) || (rc = dalloc(ctx, &ctx->b_part[0], (size_t)slot + 1)) ||
        (rc = dalloc(ctx, &ctx->b_part[1], (size_t)slot + 1)) || (rc = dalloc(ctx, &ctx->b_part[2], (size_t)slot + 1)))
      return rc;
    CK(ftb_memcpy(ctx, ctx->b_hdr, hdr.data(), hdr.size() * sizeof(BrickHdr), cudaMemcpyHostToDevice));
    CK(ftb_memcpy(ctx, ctx->b_conn16, conn16.data(), conn16.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    CK(ftb_memcpy(ctx, ctx->b_map16, map16.data(), map16.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    CK(ftb_memcpy(ctx, ctx->b_halo, halo.data(), halo.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(ftb_memcpy(ctx, ctx->s_ell, sell.data(), sell.size() * sizeof(int), cudaMemcpyHostToDevice));
    for (int c = 0; c < 3; ++c) CK(ftb_memset(ctx, ctx->b_part[c], 0, ((size_t)slot + 1) * sizeof(double)));
    if (!sover.empty()) {
      std::stable_sort(sover.begin(), sover.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b2) { return a.first < b2.first; });
      std::vector<int> ooff(nS + 1, 0), oent(sover.size());
      for (auto& pr : sover) ooff[pr.first + 1]++;
      for (int i = 0; i < nS; ++i) ooff[i + 1] += ooff[i];
      for (size_t i = 0; i < sover.size(); ++i) oent[i] = sover[i].second;
      if ((rc = dalloc(ctx, &ctx->s_ovoff, ooff.size())) || (rc = dalloc(ctx, &ctx->s_ovent, oent.size()))) return rc;
      CK(ftb_memcpy(ctx, ctx->s_ovoff, ooff.data(), ooff.size() * sizeof(int), cudaMemcpyHostToDevice));
      CK(ftb_memcpy(ctx, ctx->s_ovent, oent.data(), oent.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    ctx->brick_ok = true;
  }
  if ((rc = dalloc(ctx, &ctx->m, nNp)) || (rc = dalloc(ctx, &ctx->flags, nNp)) || (rc = dalloc(ctx, &ctx->conn, 8 * (size_t)nE)) ||
      (rc = dalloc(ctx, &ctx->pid, nE)) || (rc = dalloc(ctx, &ctx->ref_of, nE)) || (rc = dalloc(ctx, &ctx->eflag, nE)) ||
      (rc = dalloc(ctx, &ctx->felem, 24 * 32 * (size_t)cdiv(nE, 32))) || (rc = dalloc(ctx, &ctx->mp, mp.size())) ||
      (rc = dalloc(ctx, &ctx->node_off, nNp + 1)) || (rc = dalloc(ctx, &ctx->node_ent, 8 * (size_t)nE)) ||
      (rc = dalloc(ctx, &ctx->sc, 1)) || (rc = dalloc(ctx, &ctx->epart, 3 * (size_t)epart_blocks)) ||
      (rc = dalloc(ctx, &ctx->out3, 16)) || (rc = dalloc(ctx, &ctx->d_istage, 3 * (size_t)nN)) ||
      (rc = dalloc(ctx, &ctx->d_detmin, 1)) || (rc = dalloc(ctx, &ctx->d_nonpos, 1)) ||
      (rc = dalloc(ctx, &ctx->d_nref, nNp)) || (rc = dalloc(ctx, &ctx->d_nint, nN)) || (rc = dalloc(ctx, &ctx->d_ell, 8 * (size_t)nNp)))
    return rc;
  CK(ftb_memset(ctx, ctx->m, 0, nNp * sizeof(double)));
  CK(ftb_memset(ctx, ctx->eflag, 0, nE));
  CK(ftb_memset(ctx, ctx->felem, 0, 24 * 32 * (size_t)cdiv(nE, 32) * sizeof(double)));  // tiles of 32 elements (FTB_FIDX)
  CK(ftb_memset(ctx, ctx->sc, 0, sizeof(DevScalars)));
  CK(ftb_memcpy(ctx, ctx->conn, connT.data(), connT.size() * sizeof(int), cudaMemcpyHostToDevice));
  CK(ftb_memcpy(ctx, ctx->pid, pidI.data(), nE * sizeof(int), cudaMemcpyHostToDevice));
  CK(ftb_memcpy(ctx, ctx->ref_of, ref_of.data(), nE * sizeof(int), cudaMemcpyHostToDevice));
  ctx->nGP = 8LL * nE;
  if (ctx->has_tet) {  // element types (internal order) and the packed Gauss-point offsets (reference order)
    std::vector<uint8_t> etI(nE);
    std::vector<int> gpoff(nE + 1, 0);
    for (int t = 0; t < nE; ++t) etI[t] = ctx->h_etype[ref_of[t]];
    for (int e = 0; e < nE; ++e) gpoff[e + 1] = gpoff[e] + (ctx->h_etype[e] ? 1 : 8);
    ctx->nGP = gpoff[nE];
    if ((rc = dalloc(ctx, &ctx->etype, nE)) || (rc = dalloc(ctx, &ctx->gpoff, nE + 1))) return rc;
    CK(ftb_memcpy(ctx, ctx->etype, etI.data(), nE, cudaMemcpyHostToDevice));
    CK(ftb_memcpy(ctx, ctx->gpoff, gpoff.data(), (nE + 1) * sizeof(int), cudaMemcpyHostToDevice));
  }
  CK(ftb_memcpy(ctx, ctx->mp, mp.data(), mp.size() * sizeof(double), cudaMemcpyHostToDevice));
  CK(ftb_memcpy(ctx, ctx->node_off, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice));
  CK(ftb_memcpy(ctx, ctx->node_ent, ent.data(), ent.size() * sizeof(int), cudaMemcpyHostToDevice));
  CK(ftb_memcpy(ctx, ctx->d_nref, nref.data(), nNp * sizeof(int), cudaMemcpyHostToDevice));
  CK(ftb_memcpy(ctx, ctx->d_nint, nint.data(), nN * sizeof(int), cudaMemcpyHostToDevice));
  std::vector<char> overflow(nNp, 0);
  {
    std::vector<int> ell(8 * (size_t)nNp, -1);
    for (int i = 0; i < nNp; ++i) {
      const int deg = off[i + 1] - off[i];
      for (int q = 0; q < std::min(deg, 8); ++q) ell[(size_t)q * nNp + i] = ent[off[i] + q];
      if (deg > 8) overflow[i] = 1;
    }
    CK(ftb_memcpy(ctx, ctx->d_ell, ell.data(), ell.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  if (ctx->has_visco) {  // Hn_1, Hn_2, S0n zero at t = 0 (ShapeFunctions.cpp:245-252)
    const size_t n = (size_t)144 * 32 * cdiv(nE, 32);  // tiles of 32 elements (FTB_HIDX)
    if ((rc = dalloc(ctx, &ctx->hist, n))) return rc;
    CK(ftb_memset(ctx, ctx->hist, 0, n * sizeof(double)));
  }
  // flags: padding nodes are fully constrained and never counted; shared / not-owned bits
  // (CheckEnergy.cpp:21-33: a shared node is counted by the lowest rank sharing it)
  {
    std::vector<uint16_t> fl(nNp, 0);
    for (int i = 0; i < nNp; ++i) {
      if (nref[i] < 0) fl[i] = 7u | FTB_FLAG_NOTOWNED;
      if (overflow[i]) fl[i] |= FTB_FLAG_OVERFLOW;
    }
    for (int i : halo_nodes) fl[i] |= FTB_FLAG_SHARED;
    for (size_t p = 0; p < ctx->h_sendProcessID.size(); ++p)
      if (ctx->h_sendProcessID[p] < ctx->rank)
        for (int i = ctx->h_sendCum[p]; i < ctx->h_sendCum[p + 1]; ++i) fl[sendIdxInt[i]] |= FTB_FLAG_NOTOWNED;
    CK(ftb_memcpy(ctx, ctx->flags, fl.data(), nNp * sizeof(uint16_t), cudaMemcpyHostToDevice));
  }
  if (ctx->halo_count) {
    if ((rc = dalloc(ctx, &ctx->d_sendNodeIndex, ctx->halo_count)) || (rc = dalloc(ctx, &ctx->halo_nodes, ctx->nshared)) ||
        (rc = dalloc(ctx, &ctx->halo_off, ctx->nshared + 1)) || (rc = dalloc(ctx, &ctx->halo_slot, ctx->halo_count)) ||
        (rc = dalloc(ctx, &ctx->halo_node_idx, nNp)))
      return rc;
    CK(ftb_memcpy(ctx, ctx->d_sendNodeIndex, sendIdxInt.data(), ctx->halo_count * sizeof(int), cudaMemcpyHostToDevice));
    CK(ftb_memcpy(ctx, ctx->halo_nodes, halo_nodes.data(), ctx->nshared * sizeof(int), cudaMemcpyHostToDevice));
    CK(ftb_memcpy(ctx, ctx->halo_off, hoff.data(), hoff.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(ftb_memcpy(ctx, ctx->halo_slot, hslot.data(), hslot.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(ftb_memcpy(ctx, ctx->halo_node_idx, node_h.data(), nNp * sizeof(int), cudaMemcpyHostToDevice));
  }
  if (const char* ev = getenv("FTB200_ENERGY_ASYNC")) ctx->energy_async = atoi(ev) != 0;
  // ---- validate the reference configuration: detJ0 > 0 at every Gauss point --------------------
  {
    const unsigned long long inf = 0x7FF0000000000000ULL;
    CK(ftb_memcpy(ctx, ctx->d_detmin, &inf, sizeof(inf), cudaMemcpyHostToDevice));
    CK(ftb_memset(ctx, ctx->d_nonpos, 0, sizeof(int)));
    // the per-element masses land in the first 8 planes of felem (scratch until the first force call)
    LAUNCH(k_mass_elem, cdiv(nE, 128), 128, ctx->stream, elem_args(ctx, 0, nE, 1), ctx->felem, ctx->d_detmin, ctx->d_nonpos);
    LAUNCH(k_mass_gather, cdiv(nNp, 256), 256, ctx->stream, ctx->felem, ctx->node_off, ctx->node_ent, ctx->m, nNp, nE);
    unsigned long long bits = 0;
    int nonpos = 0;
    CK(cudaMemcpyAsync(&bits, ctx->d_detmin, sizeof(bits), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&nonpos, ctx->d_nonpos, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    double dmin;
    memcpy(&dmin, &bits, 8);
    if (min_detJ) *min_detJ = nonpos ? -1.0 : dmin;
    if (nonpos) return fail(ctx, FTB200_ERR_INPUT, "%d elements have a non-positive reference Jacobian", nonpos);
  }
  ctx->shape_ok = true; ctx->begun = false; ctx->bc_ok = false; ctx->has_fe = false;
  return FTB200_OK;
}

int ftb200_lumped_mass(ftb200_ctx* ctx, double* mass_out) {
  if (!ctx || !ctx->shape_ok) return fail(ctx, FTB200_ERR_INPUT, "lumped_mass: call shape_functions first");
  CK(cudaSetDevice(ctx->device));
  const int nN = ctx->nNp, nE = ctx->nE;
  const unsigned long long inf = 0x7FF0000000000000ULL;
  CK(cudaMemcpyAsync(ctx->d_detmin, &inf, sizeof(inf), cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(k_mass_elem, cdiv(nE, 128), 128, ctx->stream, elem_args(ctx, 0, nE, 1), ctx->felem, ctx->d_detmin, ctx->d_nonpos);
  LAUNCH(k_mass_gather, cdiv(nN, 256), 256, ctx->stream, ctx->felem, ctx->node_off, ctx->node_ent, ctx->m, nN, nE);
  CK(cudaGetLastError());
  if (mass_out) {
    double* src[3] = {ctx->m, ctx->m, ctx->m};  // the same value on the three dofs of a node (Mass3D.cpp:146-151)
    int rc = download_aos(ctx, src, mass_out, 0);
    if (rc) return rc;
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return FTB200_OK;
}

int ftb200_get_mass(ftb200_ctx* ctx, double* mass_out) {
  if (!ctx || !ctx->shape_ok || !mass_out) return fail(ctx, FTB200_ERR_INPUT, "get_mass: bad arguments");
  CK(cudaSetDevice(ctx->device));
  double* src[3] = {ctx->m, ctx->m, ctx->m};
  int rc = download_aos(ctx, src, mass_out, 0);
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  return FTB200_OK;
}

// ------------------------------------------------------------------------------------- legacy path
static int legacy_force_local(ftb200_ctx* ctx, const double* displacements, const double* fe, double dt) {
  int rc;
  if ((rc = upload_aos(ctx, displacements, ctx->u))) return rc;
  if (fe) {
    // external-force planes live in the padded internal node order like every other nodal plane (nNp >= nN)
    for (int k = 0; k < 3; ++k)
      if (!ctx->fe[k]) {
        if ((rc = dalloc(ctx, &ctx->fe[k], ctx->nNp))) return rc;
        CK(cudaMemsetAsync(ctx->fe[k], 0, ctx->nNp * sizeof(double), ctx->stream));
      }
    if ((rc = upload_aos(ctx, fe, ctx->fe))) return rc;
    if (!ctx->has_fe) drop_graphs(ctx);  // the graphs were captured with null fe planes
    ctx->has_fe = true;
  }
  if (ctx->has_visco) LAUNCH(k_prony, 1, 128, ctx->stream, ctx->mp, ctx->nPID, dt);
  launch_elem<true, false>(ctx, ctx->stream, 0, ctx->nE, 1);
  ctx->felem_stale = false;
  return 0;
}

int ftb200_get_force(ftb200_ctx* ctx, const double* displacements, const double* fe, double dt, double* fi, double* f_net) {
  if (!ctx || !ctx->shape_ok || !displacements) return fail(ctx, FTB200_ERR_INPUT, "get_force: bad arguments / setup incomplete");
  CK(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = legacy_force_local(ctx, displacements, fe, dt))) return rc;
  const NodeArgs N = node_args(ctx, nullptr);
  LAUNCH(k_gather_force, ctx->node_blocks, NODE_BLOCK, ctx->stream, N, ctx->fnet[0], ctx->fnet[1], ctx->fnet[2]);
  if (fi && (rc = download_aos(ctx, ctx->fi, fi, 0))) return rc;
  if (f_net && (rc = download_aos(ctx, ctx->fnet, f_net, 1))) return rc;
  CK(cudaGetLastError());  // a refused launch (bad configuration on this device) must not return stale forces
  int status = 0;
  CK(cudaMemcpyAsync(&status, &ctx->sc->status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (status & 1) return fail(ctx, FTB200_ERR_MATERIAL, "Unknown material type");
  return FTB200_OK;
}

int ftb200_calculate_accelerations(ftb200_ctx* ctx, const int* boundary, double* accelerations) {
  if (!ctx || !ctx->shape_ok || !boundary || !accelerations) return fail(ctx, FTB200_ERR_INPUT, "calculate_accelerations: bad arguments");
  CK(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = upload_boundary(ctx, boundary))) return rc;
  // a = f_net/m on free dofs into scratch planes, then a masked merge into the caller's array
  LAUNCH(k_accel, cdiv(ctx->nNp, 256), 256, ctx->stream, ctx->fnet[0], ctx->fnet[1], ctx->fnet[2], ctx->m, ctx->flags,
         ctx->a[0], ctx->a[1], ctx->a[2], ctx->nNp);
  const size_t n3 = 3 * (size_t)ctx->nN;
  LAUNCH(k_soa_to_aos, cdiv(ctx->nNp, 256), 256, ctx->stream, ctx->a[0], ctx->a[1], ctx->a[2], ctx->d_stage[0], ctx->d_nref, ctx->nNp);
  CK(cudaMemcpyAsync(ctx->d_stage[1], accelerations, n3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(k_merge_free, cdiv((long long)n3, 256), 256, ctx->stream, ctx->d_stage[0], ctx->d_stage[1], ctx->d_istage, (int)n3);
  CK(cudaMemcpyAsync(accelerations, ctx->d_stage[1], n3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return FTB200_OK;
}

int ftb200_stable_time_step(ftb200_ctx* ctx, const double* displacements, const int* boundary, double* dtMin) {
  if (!ctx || !ctx->shape_ok || !displacements || !boundary || !dtMin) return fail(ctx, FTB200_ERR_INPUT, "stable_time_step: bad arguments");
  CK(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = upload_aos(ctx, displacements, ctx->u))) return rc;
  if ((rc = upload_boundary(ctx, boundary))) return rc;
  LAUNCH(k_eflag, cdiv(ctx->nE, 256), 256, ctx->stream, ctx->conn, ctx->flags, ctx->eflag, ctx->nE);
  const unsigned long long inf = 0x7FF0000000000000ULL;
  CK(cudaMemcpyAsync(&ctx->sc->dtmin_bits, &inf, sizeof(inf), cudaMemcpyHostToDevice, ctx->stream));
  launch_elem<false, true>(ctx, ctx->stream, 0, ctx->nE, 1);
  CK(cudaGetLastError());
  unsigned long long bits = 0;
  CK(cudaMemcpyAsync(&bits, &ctx->sc->dtmin_bits, sizeof(bits), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  double d;
  memcpy(&d, &bits, 8);
  *dtMin = d > 1e20 ? 1e20 : d;  // dtMin starts at `huge`
  return FTB200_OK;
}

int ftb200_check_energy(ftb200_ctx* ctx, const double* u, const double* up, const double* v, const double* a, const double* ap,
                        const double* fi, const double* fip, const double* fe, const double* fep, const int* boundary,
                        double out[3]) {
  if (!ctx || !ctx->shape_ok || !u || !up || !v || !a || !ap || !fi || !fip || !boundary || !out)
    return fail(ctx, FTB200_ERR_INPUT, "check_energy: bad arguments");
  CK(cudaSetDevice(ctx->device));
  const size_t n3 = 3 * (size_t)ctx->nN, B = n3 * sizeof(double);
  int rc;
  if ((rc = ensure_big(ctx, 9 * B))) return rc;
  const double* hp[9] = {u, up, v, a, ap, fi, fip, fe, fep};
  double* dp[9];
  for (int i = 0; i < 9; ++i) {
    dp[i] = hp[i] ? ctx->d_big + i * n3 : nullptr;
    if (hp[i]) CK(cudaMemcpyAsync(dp[i], hp[i], B, cudaMemcpyHostToDevice, ctx->stream));
  }
  CK(cudaMemcpyAsync(ctx->d_istage, boundary, n3 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(k_energy_legacy, cdiv(ctx->nN, NODE_BLOCK), NODE_BLOCK, ctx->stream, dp[0], dp[1], dp[2], dp[3], dp[4], dp[5], dp[6], dp[7],
         dp[8], ctx->d_istage, ctx->m, ctx->flags, ctx->d_nint, ctx->epart, ctx->nN);
  LAUNCH(k_sum3, 1, 256, ctx->stream, ctx->epart, cdiv(ctx->nN, NODE_BLOCK), ctx->out3);
  CK(cudaMemcpyAsync(out, ctx->out3, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return FTB200_OK;
}

int ftb200_get_gp_outputs(ftb200_ctx* ctx, double* F, double* detF, double* pk2, double* Eavg) {
  if (!ctx || !ctx->shape_ok) return fail(ctx, FTB200_ERR_INPUT, "get_gp_outputs: setup incomplete");
  CK(cudaSetDevice(ctx->device));
  const size_t nE = ctx->nE;
  const size_t nG = (size_t)ctx->nGP;  // 8 per hexahedron, 1 per tetrahedron (ShapeFunctions.cpp:71-164)
  const size_t nF = F ? 9 * nG : 0, nD = detF ? nG : 0, nP = pk2 ? 6 * nG : 0, nEa = Eavg ? 9 * nE : 0;
  int rc;
  if ((rc = ensure_big(ctx, (nF + nD + nP + nEa + 1) * sizeof(double)))) return rc;
  double* dF = F ? ctx->d_big : nullptr;
  double* dD = detF ? ctx->d_big + nF : nullptr;
  double* dP = pk2 ? ctx->d_big + nF + nD : nullptr;
  double* dE = Eavg ? ctx->d_big + nF + nD + nP : nullptr;
  LAUNCH(k_gp_outputs, cdiv(ctx->nE, 64), 64, ctx->stream, elem_args(ctx, 0, ctx->nE, 1), ctx->ref_of, ctx->gpoff, dF, dD, dP, dE);
  if (F) CK(cudaMemcpyAsync(F, dF, nF * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (detF) CK(cudaMemcpyAsync(detF, dD, nD * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (pk2) CK(cudaMemcpyAsync(pk2, dP, nP * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (Eavg) CK(cudaMemcpyAsync(Eavg, dE, nEa * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return FTB200_OK;
}

// ----------------------------------------------------------------------------------- resident path
int ftb200_set_state(ftb200_ctx* ctx, const double* displacements, const double* velocities, const double* accelerations,
                     const int* boundary) {
  if (!ctx || !ctx->shape_ok) return fail(ctx, FTB200_ERR_INPUT, "set_state: setup incomplete");
  CK(cudaSetDevice(ctx->device));
  int rc;
  if (displacements && (rc = upload_aos(ctx, displacements, ctx->u))) return rc;
  if (velocities) { CK(cudaStreamSynchronize(ctx->stream)); if ((rc = upload_aos(ctx, velocities, ctx->v))) return rc; }
  if (accelerations) { CK(cudaStreamSynchronize(ctx->stream)); if ((rc = upload_aos(ctx, accelerations, ctx->a))) return rc; }
  if (boundary && (rc = upload_boundary(ctx, boundary))) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  return FTB200_OK;
}

int ftb200_get_state(ftb200_ctx* ctx, double* displacements, double* velocities, double* accelerations, int* boundary,
                     double* fi, double* f_net) {
  if (!ctx || !ctx->shape_ok) return fail(ctx, FTB200_ERR_INPUT, "get_state: setup incomplete");
  CK(cudaSetDevice(ctx->device));
  int rc;
  double* const* fi_src = ctx->fi;
  if (fi || f_net) {  // lazily rebuilt from the element forces of the last evaluation
    if (ctx->p2p_ready && ctx->halo_count && !ctx->halo_recv_cur) {
      unsigned long long seq = 0;
      CK(cudaMemcpyAsync(&seq, ctx->d_seq, sizeof(seq), cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      if (seq > 0) ctx->halo_recv_cur = p2p_recv(ctx->p2p_window, ctx->halo_count, (int)((seq - 1) & 1ULL), ctx->nranks);
    }
    NodeArgs N = node_args(ctx, ctx->halo_recv_cur);
    if (ctx->felem_stale) {
      // the brick-fused step keeps the element forces on the SM: evaluate them once more (materials 1 and 4 carry no
      // history).  The gather goes to the du planes (unused by that step): fi holds the sums the energy check continues
      // from, in the brick step's own summation order, and a read-back must not perturb a running calculation
      launch_elem<true, false>(ctx, ctx->stream, 0, ctx->nE, 1);
      ctx->felem_stale = false;
    }
    if (use_brick(ctx)) {
      for (int k = 0; k < 3; ++k) N.fi[k] = ctx->du[k];
      fi_src = ctx->du;
    }
    LAUNCH(k_gather_force, ctx->node_blocks, NODE_BLOCK, ctx->stream, N, ctx->fnet[0], ctx->fnet[1], ctx->fnet[2]);
  }
  double* const* src[5] = {ctx->u, ctx->v, ctx->a, fi_src, ctx->fnet};
  double* dst[5] = {displacements, velocities, accelerations, fi, f_net};
  for (int i = 0; i < 5; ++i)
    if (dst[i]) {
      if ((rc = download_aos(ctx, const_cast<double**>(src[i]), dst[i], 0))) return rc;
      CK(cudaStreamSynchronize(ctx->stream));
    }
  if (boundary) {
    LAUNCH(k_flags_to_boundary, cdiv(ctx->nNp, 256), 256, ctx->stream, ctx->flags, ctx->d_istage, ctx->d_nref, ctx->nNp);
    CK(cudaMemcpyAsync(boundary, ctx->d_istage, 3 * (size_t)ctx->nN * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return FTB200_OK;
}

int ftb200_set_bc(ftb200_ctx* ctx, const int* bc_kind, const double bc_rate[4]) {
  if (!ctx || !ctx->shape_ok || !bc_kind || !bc_rate) return fail(ctx, FTB200_ERR_INPUT, "set_bc: bad arguments");
  CK(cudaSetDevice(ctx->device));
  const size_t n3 = 3 * (size_t)ctx->nN;
  for (size_t i = 0; i < n3; ++i)
    if (bc_kind[i] < 0 || bc_kind[i] > 3) return fail(ctx, FTB200_ERR_INPUT, "set_bc: bc_kind[%zu] = %d not in 0..3", i, bc_kind[i]);
  CK(cudaMemcpyAsync(ctx->d_istage, bc_kind, n3 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(k_set_bc_kinds, cdiv(ctx->nNp, 256), 256, ctx->stream, ctx->d_istage, ctx->flags, ctx->d_nref, ctx->nNp);
  CK(cudaMemcpyAsync(&ctx->sc->bc_rate[0], bc_rate, 4 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->bc_ok = true;
  return FTB200_OK;
}

int ftb200_record_history(ftb200_ctx* ctx, long long capacity) {
  if (!ctx || !ctx->shape_ok || capacity < 0) return fail(ctx, FTB200_ERR_INPUT, "record_history: bad arguments");
  CK(cudaSetDevice(ctx->device));
  dfree(ctx->dthist); dfree(ctx->ehist);
  ctx->hist_cap = capacity;
  if (capacity > 0) {
    int rc;
    if ((rc = dalloc(ctx, &ctx->dthist, (size_t)capacity)) || (rc = dalloc(ctx, &ctx->ehist, 4 * (size_t)capacity))) return rc;
    CK(ftb_memset(ctx, ctx->dthist, 0, capacity * sizeof(double)));
    CK(ftb_memset(ctx, ctx->ehist, 0, 4 * capacity * sizeof(double)));
  }
  CK(ftb_memcpy(ctx, &ctx->sc->hist_cap, &capacity, sizeof(long long), cudaMemcpyHostToDevice));
  drop_graphs(ctx);  // dthist / ehist are kernel arguments of every captured loop (single-partition and peer-memory)
  dfree(ctx->inj_hist);
  if (ctx->injury && capacity > 0) {
    int rc;
    if ((rc = dalloc(ctx, &ctx->inj_hist, 2 * (size_t)capacity))) return rc;
    CK(ftb_memset(ctx, ctx->inj_hist, 0, 2 * capacity * sizeof(double)));
  }
  return FTB200_OK;
}

int ftb200_get_history(ftb200_ctx* ctx, long long first, long long count, double* dt_hist, double* energy_hist4) {
  if (!ctx || first < 0 || count < 0 || first + count > ctx->hist_cap) return fail(ctx, FTB200_ERR_INPUT, "get_history: range");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  if (dt_hist && count) CK(ftb_memcpy(ctx, dt_hist, ctx->dthist + first, count * sizeof(double), cudaMemcpyDeviceToHost));
  if (energy_hist4 && count) CK(ftb_memcpy(ctx, energy_hist4, ctx->ehist + 4 * first, 4 * count * sizeof(double), cudaMemcpyDeviceToHost));
  return FTB200_OK;
}

// explicit_begin in three phases so that a rank with shared nodes can put the cross-rank MIN of the time
// step and the neighbour sum of the initial forces in between (single rank: the three run back to back)
int ftb200_explicit_begin_dt(ftb200_ctx* ctx, double Time0, double reduction, double failure_dt, int energy_every,
                             double** dtmin_dev) {
  if (!ctx || !ctx->shape_ok) return fail(ctx, FTB200_ERR_INPUT, "explicit_begin: setup incomplete");
  if (energy_every != 0 && energy_every != 1)
    return fail(ctx, FTB200_ERR_INPUT, "explicit_begin: energy_every must be 0 or 1 (the running sums need every step)");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  ctx->energy = energy_every;
  ctx->Time0 = Time0;
  ctx->halo_recv_cur = nullptr;
  // the 25-step graph of the standard loop survives a restart: its kernel arguments are device pointers and flags that
  // explicit_begin does not change (build_graph compares the ones that can; injury / rigid-body / history / stream
  // changes drop it themselves).  Rebuilding it cost 1-2 ms of every ExplicitDynamics call that starts from host state.
  // scalars: keep bc_rate / hist_cap, reset the rest
  DevScalars h;
  CK(cudaStreamSynchronize(s));  // a run enqueued earlier may still be writing the scalars and the step ring
  CK(ftb_memcpy(ctx, &h, ctx->sc, sizeof(h), cudaMemcpyDeviceToHost));
  double rate[4];
  memcpy(rate, h.bc_rate, sizeof(rate));
  memset(&h, 0, sizeof(h));
  memcpy(h.bc_rate, rate, sizeof(rate));
  h.hist_cap = ctx->hist_cap;
  h.ring = ctx->ring_dev; h.ring_cap = ctx->ring_cap;
  // records are counted from this explicit_begin: a slot still holding step k of the previous run must not satisfy a
  // wait for step k of this one (the stream is idle here)
  for (long long i = 0; ctx->ring_host && i < ctx->ring_cap; ++i) ctx->ring_host[8 * i + 2] = -1.0;
  h.Time = Time0; h.tMax = 1e300; h.reduction = reduction; h.failure_dt = failure_dt;
  h.dtmin_bits = 0x7FF0000000000000ULL;
  h.energy_every = energy_every;
  CK(cudaMemcpyAsync(ctx->sc, &h, sizeof(h), cudaMemcpyHostToDevice, s));
  const int nb = cdiv(ctx->nNp, 256);
  // ApplyBoundaryConditions(Time0) (Benchmarking-Parallel.cpp:83)
  LAUNCH(k_apply_bc, nb, 256, s, ctx->u[0], ctx->u[1], ctx->u[2], ctx->v[0], ctx->v[1], ctx->v[2], ctx->a[0], ctx->a[1],
         ctx->a[2], ctx->flags, ctx->sc, Time0, ctx->nNp);
  LAUNCH(k_eflag, cdiv(ctx->nE, 256), 256, s, ctx->conn, ctx->flags, ctx->eflag, ctx->nE);
  // dt = reduction * StableTimeStep() (:86) -- before GetForce because material 5 reads dt
  launch_elem<false, true>(ctx, s, 0, ctx->nE, 1);
  if (dtmin_dev) *dtmin_dev = reinterpret_cast<double*>(&ctx->sc->dtmin_bits);
  return FTB200_OK;
}

int ftb200_explicit_begin_force(ftb200_ctx* ctx, double* send_dev) {
  if (!ctx || !ctx->shape_ok || (ctx->halo_count && !send_dev)) return fail(ctx, FTB200_ERR_INPUT, "explicit_begin_force: bad arguments");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  LAUNCH((k_adv<true>), 1, 128, s, ctx->sc, ctx->mp, ctx->nPID, ctx->Time0, ctx->dthist);
  // GetForce() (:88)
  launch_elem<true, false>(ctx, s, 0, ctx->nE, 1);
  ctx->felem_stale = false;
  if (ctx->halo_count) {
    // partial sums go to the f_net planes (scratch in the resident path): fi still holds the previous
    // step's total, which the energy check needs as fi_prev
    LAUNCH(k_gather_shared, cdiv(ctx->nshared, 128), 128, s, ctx->felem, ctx->node_off, ctx->node_ent, ctx->halo_nodes,
           ctx->fnet[0], ctx->fnet[1], ctx->fnet[2], ctx->nshared, ctx->nE);
    LAUNCH(k_halo_pack, cdiv(ctx->halo_count, 256), 256, s, ctx->fnet[0], ctx->fnet[1], ctx->fnet[2], ctx->d_sendNodeIndex, send_dev,
           ctx->halo_count);
  }
  return FTB200_OK;
}

int ftb200_explicit_begin_finish(ftb200_ctx* ctx, const double* recv_dev) {
  if (!ctx || !ctx->shape_ok || (ctx->halo_count && !recv_dev)) return fail(ctx, FTB200_ERR_INPUT, "explicit_begin_finish: bad arguments");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  // CalculateAccelerations() (:91)
  const NodeArgs N = node_args(ctx, ctx->halo_count ? recv_dev : nullptr);
  LAUNCH((k_node<true, false, false, false>), ctx->node_blocks, NODE_BLOCK, s, N);
  ctx->halo_recv_cur = ctx->halo_count ? recv_dev : nullptr;
  int status = 0;
  CK(cudaMemcpyAsync(&status, &ctx->sc->status, sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  ctx->begun = true;
  if (status & 1) return fail(ctx, FTB200_ERR_MATERIAL, "Unknown material type");
  if (status & 16) return fail(ctx, FTB200_ERR_TIMESTEP, "Timestep too small");
  return FTB200_OK;
}

int ftb200_explicit_begin(ftb200_ctx* ctx, double Time0, double reduction, double failure_dt, int energy_every) {
  if (ctx && ctx->halo_count && ctx->nranks > 1)
    return fail(ctx, FTB200_ERR_INPUT, "explicit_begin: this rank has shared nodes; use the explicit_begin_dt/_force/_finish sequence");
  int rc;
  if ((rc = ftb200_explicit_begin_dt(ctx, Time0, reduction, failure_dt, energy_every, nullptr))) return rc;
  if ((rc = ftb200_explicit_begin_force(ctx, nullptr))) return rc;
  return ftb200_explicit_begin_finish(ctx, nullptr);
}

static void join_energy(ftb200_ctx* ctx) {
  if (ctx->energy_pending) { cudaStreamWaitEvent(ctx->stream, ctx->ev_energy_done, 0); ctx->energy_pending = false; }
}

static int build_graph(ftb200_ctx* ctx) {
  const int sig = ctx->energy | (ctx->has_fe ? 2 : 0) | (ctx->energy_async ? 8 : 0) | (ctx->injury ? 64 : 0) | (ctx->rigid ? 128 : 0) |
                  (use_brick(ctx) ? 256 : 0);
  if (ctx->graph && ctx->graph_energy == sig) return 0;
  if (ctx->graph) { cudaGraphExecDestroy(ctx->graph); ctx->graph = nullptr; }
  cudaGraph_t g = nullptr;
  const long long before = ctx->launches;
  CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < GRAPH_STEPS; ++i) {
    if (use_brick(ctx)) launch_step_brick(ctx);
    else launch_step(ctx, nullptr);
  }
  join_energy(ctx);  // the helper stream rejoins before the capture ends
  cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
  ctx->launches = before;  // captured, not launched
  if (e != cudaSuccess) return fail(ctx, FTB200_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
  e = cudaGraphInstantiate(&ctx->graph, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) { ctx->graph = nullptr; return fail(ctx, FTB200_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e)); }
  ctx->graph_energy = sig;
  return 0;
}

// ---- peer-memory multi-GPU step: boundary elements -> pack into the neighbours' windows || interior elements ->
//      dt exchange + arrival waits + scalar update -> node kernel (receive window selected by step parity)
static void launch_step_p2p(ftb200_ctx* ctx) {
  cudaStream_t s = ctx->stream, s2 = ctx->stream_lo;
  const int nEb = ctx->nE_boundary;
  unsigned long long* tr = ctx->trace;
  if (tr) LAUNCH(k_stamp, 1, 1, s, tr, ctx->sc, 0, 0);
  // The partial sums of the shared nodes should cross NVLink while the interior is integrated.  With a separate pack
  // kernel they do not: whatever is launched behind the interior grid gets its first block only when that grid has
  // handed out all of its blocks (the element kernel allocates every register of an SM) -- the trace
  // (profiles/r02_p2p_trace_2gpu_before.txt) shows the marker behind the boundary elements 103 us into the step and the
  // pack kernel done at 123 us.  p2p_order (two-stream form, measured, none helps): 0 stream priorities only, 1 explicit
  // priority attribute on every launch, 2 the interior waits for a programmatic event that fires once every boundary
  // block has started, 3 the interior waits for the boundary grid to finish.  Hence the fused exchange below.
  const int order = nEb > 0 ? ctx->p2p_order : 0;
  // Fused exchange on a mesh that one hexahedron kernel covers (one material, one geometry class): ONE launch over all
  // elements.  Blocks are handed out in element order, so the boundary elements -- first in the internal order -- run
  // first and their epilogue sends under the interior blocks of the same grid; no second stream, no fork / join.
  bool one_launch = false;
  if (ctx->p2p_fused && !ctx->ranges.empty() && ctx->ranges.size() <= 2 && ctx->ranges.front().e0 == 0 && ctx->ranges.back().e1 == ctx->nE) {
    const auto &r0 = ctx->ranges.front(), &r1 = ctx->ranges.back();
    one_launch = !r0.tet && !r1.tet && r0.mat == r1.mat && r0.affine == r1.affine;
  }
  if (one_launch) {
    ctx->pk_on = true;
    launch_elem_hex<true, true>(ctx, s, 0, ctx->nE, 0, ctx->ranges.front().mat, ctx->ranges.front().affine);
    ctx->pk_on = false;
    if (tr) { LAUNCH(k_stamp, 1, 1, s, tr, ctx->sc, 1, 0); LAUNCH(k_stamp, 1, 1, s, tr, ctx->sc, 2, 0); LAUNCH(k_stamp, 1, 1, s, tr, ctx->sc, 7, 0); }
  } else {
  // two-stream form: elements touching shared nodes on the main stream, the interior behind them on its own stream
  cudaEventRecord(ctx->ev_fork, s);
  if (order == 1) { ctx->la_kind = 1; ctx->la_prio = ctx->prio_hi; }
  if (order == 2) { ctx->la_kind = 2; ctx->la_event = ctx->ev_prog; }
  ctx->pk_on = ctx->p2p_fused;  // only this launch carries the exchange epilogue
  launch_elem<true, true>(ctx, s, 0, nEb, 0);
  ctx->pk_on = false;
  ctx->la_kind = 0;
  if (tr) LAUNCH(k_stamp, 1, 1, s, tr, ctx->sc, 1, 0);
  if (order == 3) cudaEventRecord(ctx->ev_prog, s);
  cudaStreamWaitEvent(s2, (order == 2 || order == 3) ? ctx->ev_prog : ctx->ev_fork, 0);
  if (order == 1) { ctx->la_kind = 1; ctx->la_prio = ctx->prio_lo; }
  launch_elem<true, true>(ctx, s2, nEb, ctx->nE, 0);
  ctx->la_kind = 0;
  if (tr) LAUNCH(k_stamp, 1, 1, s2, tr, ctx->sc, 7, 0);
  cudaEventRecord(ctx->ev_join, s2);
  if (order == 1) { ctx->la_kind = 1; ctx->la_prio = ctx->prio_hi; }
  if (ctx->halo_count && !ctx->p2p_fused)
    LAUNCH(k_p2p_pack, cdiv(ctx->halo_count, 128), 128, s, ctx->p2p, ctx->felem, ctx->node_off, ctx->node_ent,
           ctx->d_sendNodeIndex, ctx->sc, ctx->nE);
  if (tr) LAUNCH(k_stamp, 1, 1, s, tr, ctx->sc, 2, 0);
  cudaStreamWaitEvent(s, ctx->ev_join, 0);
  }
  // k_adv_p2p moves sc->step / sc->active, which the energy reduction of the previous step (helper stream) still reads
  if (ctx->energy_pending) { cudaStreamWaitEvent(s, ctx->ev_energy_done, 0); ctx->energy_pending = false; }
  LAUNCH(k_adv_p2p, 1, 128, s, ctx->p2p, ctx->sc, ctx->mp, ctx->nPID, ctx->dthist, tr, 0);
  if (ctx->rigid) LAUNCH(k_rigid_step, 1, 32, s, ctx->sc, ctx->rigid, 0);
  NodeArgs N = node_args(ctx, ctx->halo_count ? p2p_recv(ctx->p2p_window, ctx->halo_count, 0, ctx->nranks) : nullptr);
  if (ctx->halo_count) {
    N.halo_recv_alt = p2p_recv(ctx->p2p_window, ctx->halo_count, 1, ctx->nranks);
    N.p2p_seq = ctx->d_seq;
  }
  if (ctx->energy) LAUNCH((k_node<true, true, true, true>), ctx->node_blocks, NODE_BLOCK, s, N);
  else LAUNCH((k_node<true, true, true, false>), ctx->node_blocks, NODE_BLOCK, s, N);
  ctx->la_kind = 0;
  if (tr) LAUNCH(k_stamp, 1, 1, s, tr, ctx->sc, 6, -1);  // k_adv_p2p has advanced sc->step
  if (ctx->injury) {
    // CalculateInjuryCriterions across partitions inside the loop: running extrema per rank, then the six radix passes of
    // the two GLOBAL 95th-percentile selections, each with its histogram summed over the ranks through the windows
    launch_injury(ctx, s);  // one rank: everything; several ranks: the extrema only
    if (ctx->nranks > 1) {
      const ElemArgs A = elem_args(ctx, 0, ctx->nE, 0);
      double* h0 = ctx->inj_hist;
      double* h1 = ctx->inj_hist ? ctx->inj_hist + ctx->hist_cap : nullptr;
      for (int pass = 0; pass < INJ_PASSES; ++pass) {
        LAUNCH(k_injury_select, dim3(std::min(INJ_BLOCKS, cdiv(ctx->nE, INJ_THREADS * INJ_ITEMS)), 2), INJ_THREADS, s, A, ctx->inj_state,
               pass, nullptr, nullptr, 1);
        LAUNCH(k_injury_xchg, 1, INJ_THREADS, s, ctx->p2p, ctx->sc, ctx->inj_state, pass);
        LAUNCH(k_injury_pick, dim3(1, 2), INJ_THREADS, s, ctx->sc, ctx->inj_state, pass, h0, h1);
      }
      LAUNCH(k_injury_lists, cdiv(ctx->nE, 256), 256, s, A, ctx->inj_state);
    }
  }
  if (ctx->energy) {
    if (ctx->energy_async && !ctx->profile) {  // K8 of this step under the element kernels of the next one, like the single-partition loop
      cudaEventRecord(ctx->ev_nodes_done, s);
      cudaStreamWaitEvent(ctx->stream2, ctx->ev_nodes_done, 0);
      LAUNCH(k_energy, 1, 256, ctx->stream2, ctx->sc, ctx->epart, ctx->node_blocks, ctx->ehist);
      cudaEventRecord(ctx->ev_energy_done, ctx->stream2);
      ctx->energy_pending = true;
    } else {
      LAUNCH(k_energy, 1, 256, s, ctx->sc, ctx->epart, ctx->node_blocks, ctx->ehist);
    }
  }
}

static void join_energy(ftb200_ctx* ctx);
static int run_async_p2p(ftb200_ctx* ctx, double tMax, long long steps) {
  cudaStream_t s = ctx->stream;
  if (!ctx->trace)
    if (const char* ev = getenv("FTB200_P2P_TRACE")) {
      if (*ev && dalloc(ctx, &ctx->trace, (size_t)TRACE_STEPS * TRACE_SLOTS) == 0) {
        ftb_memset(ctx, ctx->trace, 0, sizeof(unsigned long long) * TRACE_STEPS * TRACE_SLOTS);
        ctx->trace_prefix = ev;
      }
    }
  LAUNCH(k_begin_run, 1, 1, s, ctx->sc, tMax, steps);
  {
    const NodeArgs N = node_args(ctx, nullptr);
    if (ctx->rigid) LAUNCH(k_rigid_step, 1, 32, s, ctx->sc, ctx->rigid, 1);
    if (ctx->energy) LAUNCH((k_node<false, true, false, true>), ctx->node_blocks, NODE_BLOCK, s, N);
    else LAUNCH((k_node<false, true, false, false>), ctx->node_blocks, NODE_BLOCK, s, N);
  }
  const bool use_graph = !ctx->profile;  // built on the first run (warm-up), whatever its length
  const int p2p_sig = ctx->energy | (ctx->has_fe ? 2 : 0) | (ctx->injury ? 64 : 0) | (ctx->rigid ? 128 : 0);
  if (use_graph && !(ctx->p2p_graph && ctx->p2p_graph_energy == p2p_sig)) {
    if (ctx->p2p_graph) { cudaGraphExecDestroy(ctx->p2p_graph); ctx->p2p_graph = nullptr; }
    cudaGraph_t g = nullptr;
    const long long before = ctx->launches;
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < GRAPH_STEPS; ++i) launch_step_p2p(ctx);
    join_energy(ctx);  // the helper stream rejoins before the capture ends
    cudaError_t e = cudaStreamEndCapture(s, &g);
    const long long per_graph = ctx->launches - before;
    ctx->launches = before;
    if (e != cudaSuccess) return fail(ctx, FTB200_ERR_CUDA, "p2p graph capture failed: %s", cudaGetErrorString(e));
    if (ctx->la_err) return fail(ctx, FTB200_ERR_CUDA, "p2p graph capture: a launch with attributes failed: %s (FTB200_P2P_ORDER=%d)",
                                 cudaGetErrorString((cudaError_t)ctx->la_err), ctx->p2p_order);
    e = cudaGraphInstantiate(&ctx->p2p_graph, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) { ctx->p2p_graph = nullptr; return fail(ctx, FTB200_ERR_CUDA, "p2p graph instantiate failed: %s", cudaGetErrorString(e)); }
    ctx->p2p_graph_energy = p2p_sig;
    ctx->p2p_graph_launches = per_graph;
  }
  long long left = steps;
  while (left > 0) {
    if (use_graph && left >= GRAPH_STEPS) {
      CK(cudaGraphLaunch(ctx->p2p_graph, s));
      ctx->launches += ctx->p2p_graph_launches;
      left -= GRAPH_STEPS;
    } else {
      launch_step_p2p(ctx);
      left--;
    }
  }
  join_energy(ctx);
  // the receive window of the last step, for a later get_state
  ctx->halo_recv_cur = nullptr;
  CK(cudaGetLastError());
  return FTB200_OK;
}

int ftb200_explicit_run_async(ftb200_ctx* ctx, double tMax, long long steps) {
  if (!ctx || !ctx->begun) return fail(ctx, FTB200_ERR_INPUT, "explicit_run: call explicit_begin first");
  CK(cudaSetDevice(ctx->device));
  if (ctx->nranks > 1 || ctx->p2p_ready) {  // (a single rank that imported its own window runs the partitioned loop too: bench)
    if (!ctx->p2p_ready)
      return fail(ctx, FTB200_ERR_INPUT, "explicit_run: multi-rank runs need the peer-memory windows (p2p_export/import) "
                                         "or the step_begin/step_join/step_end sequence");
    if (steps <= 0) return FTB200_OK;
    return run_async_p2p(ctx, tMax, steps);
  }
  if (steps <= 0) return FTB200_OK;
  cudaStream_t s = ctx->stream;
  LAUNCH(k_begin_run, 1, 1, s, ctx->sc, tMax, steps);
  const bool brick = use_brick(ctx);
  if (brick) LAUNCH(k_brick_prep, cdiv(std::max(ctx->nNp, ctx->nE), 256), 256, s, ctx->flags, ctx->b_flags32, ctx->nNp, ctx->pid, ctx->eflag, ctx->b_pe, ctx->nE);
  if (!brick) {  // (the brick-fused step starts each step itself)
    const NodeArgs N = node_args(ctx, nullptr);
    if (ctx->rigid) LAUNCH(k_rigid_step, 1, 32, s, ctx->sc, ctx->rigid, 1);
    if (ctx->energy) LAUNCH((k_node<false, true, false, true>), ctx->node_blocks, NODE_BLOCK, s, N);
    else LAUNCH((k_node<false, true, false, false>), ctx->node_blocks, NODE_BLOCK, s, N);
  }
  const int per_step = brick ? (2 + (ctx->nSurf > 0 ? 1 : 0) + (ctx->energy ? 1 : 0))
                             : 3 + (ctx->energy ? 1 : 0) + (ctx->injury ? INJ_LAUNCHES : 0) + (ctx->rigid ? 1 : 0);
  long long left = steps;
  const bool use_graph = !ctx->profile;  // built on the first run (warm-up), whatever its length
  if (use_graph) {
    ctx->energy_async_now = true;
    int rc = build_graph(ctx);
    if (rc) return rc;
  }
  ctx->energy_async_now = steps >= 2;
  while (left > 0) {
    if (use_graph && left >= GRAPH_STEPS) {
      CK(cudaGraphLaunch(ctx->graph, s));
      ctx->launches += (long long)per_step * GRAPH_STEPS;
      left -= GRAPH_STEPS;
    } else {
      if (brick) launch_step_brick(ctx);
      else launch_step(ctx, nullptr);
      left--;
    }
  }
  if (brick) ctx->felem_stale = true;
  join_energy(ctx);
  CK(cudaGetLastError());
  return FTB200_OK;
}

int ftb200_explicit_poll_async(ftb200_ctx* ctx, double* out8_pinned) {
  if (!ctx || !ctx->begun || !out8_pinned) return fail(ctx, FTB200_ERR_INPUT, "explicit_poll_async: bad arguments");
  CK(cudaSetDevice(ctx->device));
  LAUNCH(k_scalars_out, 1, 1, ctx->stream, ctx->sc, ctx->out3 + 8);  // staged through device memory
  CK(cudaMemcpyAsync(out8_pinned, ctx->out3 + 8, 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  return FTB200_OK;
}

int ftb200_explicit_poll(ftb200_ctx* ctx, long long* steps_done, double* Time, double* dt, int* status_bits) {
  if (!ctx || !ctx->begun) return fail(ctx, FTB200_ERR_INPUT, "explicit_poll: call explicit_begin first");
  CK(cudaSetDevice(ctx->device));
  DevScalars h;
  CK(cudaMemcpyAsync(&h, ctx->sc, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->profile) prof_collect(ctx);
  if (steps_done) *steps_done = h.step;
  if (Time) *Time = h.Time;
  if (dt) *dt = h.ndt;
  if (status_bits) *status_bits = h.status;
  if (h.status & 64) return fail(ctx, FTB200_ERR_CUDA, "peer-memory exchange timed out: a neighbour rank never delivered its step");
  return FTB200_OK;
}

int ftb200_explicit_run(ftb200_ctx* ctx, double tMax, long long maxSteps, long long* steps_done, double* Time, double* dt) {
  if (!ctx || !ctx->begun) return fail(ctx, FTB200_ERR_INPUT, "explicit_run: call explicit_begin first");
  long long s0 = 0, s1 = 0;
  double T = 0, d = 0;
  int st = 0, rc;
  if ((rc = ftb200_explicit_poll(ctx, &s0, &T, &d, &st))) return rc;
  long long left = maxSteps;
  s1 = s0;
  while (left > 0 && T < tMax && !(st & 16)) {
    const long long chunk = std::min<long long>(left, 200);
    if ((rc = ftb200_explicit_run_async(ctx, tMax, chunk))) return rc;
    long long s2;
    if ((rc = ftb200_explicit_poll(ctx, &s2, &T, &d, &st))) return rc;
    left -= chunk;
    if (s2 == s1) break;
    s1 = s2;
  }
  if (steps_done) *steps_done = s1 - s0;
  if (Time) *Time = T;
  if (dt) *dt = d;
  if (st & 1) return fail(ctx, FTB200_ERR_MATERIAL, "Unknown material type");
  if (st & 16) return fail(ctx, FTB200_ERR_TIMESTEP, "Timestep too small, dt below FailureTimeStep");
  return FTB200_OK;
}

int ftb200_step_ring(ftb200_ctx* ctx, long long capacity, double** host_ring) {
  if (!ctx || !ctx->shape_ok || capacity < 0) return fail(ctx, FTB200_ERR_INPUT, "step_ring: bad arguments");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->ring_host) { cudaFreeHost(ctx->ring_host); ctx->ring_host = nullptr; ctx->ring_dev = nullptr; }
  ctx->ring_cap = 0;
  if (capacity > 0) {
    void* hp = nullptr;
    void* dp = nullptr;
    CK(cudaHostAlloc(&hp, 8 * sizeof(double) * (size_t)capacity, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer(&dp, hp, 0));
    ctx->ring_host = static_cast<double*>(hp);
    ctx->ring_dev = static_cast<double*>(dp);
    ctx->ring_cap = capacity;
    for (long long i = 0; i < 8 * capacity; ++i) ctx->ring_host[i] = (i % 8 == 2) ? -1.0 : 0.0;
  }
  CK(ftb_memcpy(ctx, &ctx->sc->ring, &ctx->ring_dev, sizeof(double*), cudaMemcpyHostToDevice));
  CK(ftb_memcpy(ctx, &ctx->sc->ring_cap, &ctx->ring_cap, sizeof(long long), cudaMemcpyHostToDevice));
  if (host_ring) *host_ring = ctx->ring_host;
  return FTB200_OK;
}

int ftb200_get_energy(ftb200_ctx* ctx, double out[4]) {
  if (!ctx || !ctx->begun || !out) return fail(ctx, FTB200_ERR_INPUT, "get_energy: bad arguments");
  CK(cudaSetDevice(ctx->device));
  DevScalars h;
  CK(cudaMemcpyAsync(&h, ctx->sc, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  out[0] = h.Wint; out[1] = h.Wext; out[2] = h.WKE; out[3] = h.Etot;
  return FTB200_OK;
}

// ------------------------------------------------------------------------------------------ halo
int ftb200_halo_count(const ftb200_ctx* ctx) { return ctx ? ctx->halo_count : 0; }

int ftb200_halo_pack(ftb200_ctx* ctx, int field, double* send_dev) {
  if (!ctx || !ctx->shape_ok || !send_dev || (field != 0 && field != 1)) return fail(ctx, FTB200_ERR_INPUT, "halo_pack: bad arguments");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->halo_count) return FTB200_OK;
  const double *x = field ? ctx->m : ctx->fi[0], *y = field ? ctx->m : ctx->fi[1], *z = field ? ctx->m : ctx->fi[2];
  LAUNCH(k_halo_pack, cdiv(ctx->halo_count, 256), 256, ctx->stream, x, y, z, ctx->d_sendNodeIndex, send_dev, ctx->halo_count);
  return FTB200_OK;
}

int ftb200_halo_add(ftb200_ctx* ctx, int field, const double* recv_dev) {
  if (!ctx || !ctx->shape_ok || !recv_dev || (field != 0 && field != 1)) return fail(ctx, FTB200_ERR_INPUT, "halo_add: bad arguments");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->halo_count) return FTB200_OK;
  if (field == 1) {
    // mass: one value per node; the three received dofs are identical, add component 0
    LAUNCH(k_halo_add, cdiv(ctx->nshared, 256), 256, ctx->stream, ctx->m, ctx->d_stage[2], ctx->d_stage[2] + ctx->nNp,
           ctx->halo_nodes, ctx->halo_off, ctx->halo_slot, recv_dev, ctx->nshared);
  } else {
    LAUNCH(k_halo_add, cdiv(ctx->nshared, 256), 256, ctx->stream, ctx->fi[0], ctx->fi[1], ctx->fi[2], ctx->halo_nodes,
           ctx->halo_off, ctx->halo_slot, recv_dev, ctx->nshared);
  }
  return FTB200_OK;
}

int ftb200_run_begin(ftb200_ctx* ctx, double tMax, long long steps) {
  if (!ctx || !ctx->begun) return fail(ctx, FTB200_ERR_INPUT, "run_begin: call explicit_begin first");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  LAUNCH(k_begin_run, 1, 1, s, ctx->sc, tMax, steps);
  const NodeArgs N = node_args(ctx, nullptr);
  if (ctx->rigid) LAUNCH(k_rigid_step, 1, 32, s, ctx->sc, ctx->rigid, 1);
  if (ctx->energy) LAUNCH((k_node<false, true, false, true>), ctx->node_blocks, NODE_BLOCK, s, N);
  else LAUNCH((k_node<false, true, false, false>), ctx->node_blocks, NODE_BLOCK, s, N);
  return FTB200_OK;
}

int ftb200_step_begin(ftb200_ctx* ctx, double* send_dev, double** dtmin_dev) {
  if (!ctx || !ctx->begun || (ctx->halo_count && !send_dev)) return fail(ctx, FTB200_ERR_INPUT, "step_begin: bad arguments");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream, s2 = ctx->stream_lo;
  const int nEb = ctx->nE_boundary;
  // elements touching a shared node first (main stream), then the partial f_int of the shared nodes into the send
  // window; the interior elements run on the low-priority stream: they only need the state left by the previous node kernel
  CK(cudaEventRecord(ctx->ev_fork, s));
  launch_elem<true, true>(ctx, s, 0, nEb, 0);
  CK(cudaStreamWaitEvent(s2, ctx->ev_fork, 0));
  launch_elem<true, true>(ctx, s2, nEb, ctx->nE, 0);
  CK(cudaEventRecord(ctx->ev_join, s2));
  if (ctx->halo_count) {
    // partial sums go to the f_net planes (scratch in the resident path): fi still holds the previous
    // step's total, which the energy check needs as fi_prev
    LAUNCH(k_gather_shared, cdiv(ctx->nshared, 128), 128, s, ctx->felem, ctx->node_off, ctx->node_ent, ctx->halo_nodes,
           ctx->fnet[0], ctx->fnet[1], ctx->fnet[2], ctx->nshared, ctx->nE);
    LAUNCH(k_halo_pack, cdiv(ctx->halo_count, 256), 256, s, ctx->fnet[0], ctx->fnet[1], ctx->fnet[2], ctx->d_sendNodeIndex, send_dev,
           ctx->halo_count);
  }
  if (dtmin_dev) *dtmin_dev = reinterpret_cast<double*>(&ctx->sc->dtmin_bits);
  return FTB200_OK;
}

int ftb200_step_join(ftb200_ctx* ctx) {
  if (!ctx || !ctx->begun) return fail(ctx, FTB200_ERR_INPUT, "step_join: bad arguments");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
  return FTB200_OK;
}

int ftb200_step_end(ftb200_ctx* ctx, const double* recv_dev) {
  if (!ctx || !ctx->begun || (ctx->halo_count && !recv_dev)) return fail(ctx, FTB200_ERR_INPUT, "step_end: bad arguments");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  LAUNCH((k_adv<false>), 1, 128, s, ctx->sc, ctx->mp, ctx->nPID, 0.0, ctx->dthist);
  if (ctx->rigid) LAUNCH(k_rigid_step, 1, 32, s, ctx->sc, ctx->rigid, 0);
  const NodeArgs N = node_args(ctx, ctx->halo_count ? recv_dev : nullptr);
  if (ctx->energy) LAUNCH((k_node<true, true, true, true>), ctx->node_blocks, NODE_BLOCK, s, N);
  else LAUNCH((k_node<true, true, true, false>), ctx->node_blocks, NODE_BLOCK, s, N);
  if (ctx->energy) LAUNCH(k_energy, 1, 256, s, ctx->sc, ctx->epart, ctx->node_blocks, ctx->ehist);
  if (ctx->injury) launch_injury(ctx, s);  // several partitions: running extrema only, the selection passes follow
  ctx->halo_recv_cur = ctx->halo_count ? recv_dev : nullptr;
  return FTB200_OK;
}
// ---------------------------------------------------------------------------------------------------------------
// Host-buffer variants of the exchange entry points, for a host that moves the shared-node windows itself (the
// reference's MPI_Isend / MPI_Irecv of sendNodeDisplacement / recvNodeDisplacement, GetForce_3D.cpp:54-102,
// Mass3D.cpp:77-125) and has no CUDA of its own: the device windows live inside the context.
static int ensure_xwin(ftb200_ctx* ctx) {
  if (ctx->d_xsend || !ctx->halo_count) return 0;
  int rc;
  if ((rc = dalloc(ctx, &ctx->d_xsend, 3 * (size_t)ctx->halo_count)) || (rc = dalloc(ctx, &ctx->d_xrecv, 3 * (size_t)ctx->halo_count))) return rc;
  return 0;
}
static int xsend_to_host(ftb200_ctx* ctx, double* send_host) {
  if (!ctx->halo_count) return 0;
  CK(cudaMemcpyAsync(send_host, ctx->d_xsend, 3 * (size_t)ctx->halo_count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
static int xrecv_from_host(ftb200_ctx* ctx, const double* recv_host) {
  if (!ctx->halo_count) return 0;
  CK(cudaMemcpyAsync(ctx->d_xrecv, recv_host, 3 * (size_t)ctx->halo_count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}
int ftb200_device_count(void) {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}
int ftb200_halo_pack_host(ftb200_ctx* ctx, int field, double* send_host) {
  if (!ctx || !ctx->shape_ok || (ctx->halo_count && !send_host)) return fail(ctx, FTB200_ERR_INPUT, "halo_pack_host: bad arguments");
  int rc;
  if ((rc = ensure_xwin(ctx)) || (ctx->halo_count && (rc = ftb200_halo_pack(ctx, field, ctx->d_xsend)))) return rc;
  return xsend_to_host(ctx, send_host);
}
int ftb200_halo_add_host(ftb200_ctx* ctx, int field, const double* recv_host) {
  if (!ctx || !ctx->shape_ok || (ctx->halo_count && !recv_host)) return fail(ctx, FTB200_ERR_INPUT, "halo_add_host: bad arguments");
  int rc;
  if (!ctx->halo_count) return FTB200_OK;
  if ((rc = ensure_xwin(ctx)) || (rc = xrecv_from_host(ctx, recv_host)) || (rc = ftb200_halo_add(ctx, field, ctx->d_xrecv))) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  return FTB200_OK;
}
// legacy GetForce on several ranks: local element forces and the packed partial sums of the shared nodes ...
int ftb200_get_force_begin(ftb200_ctx* ctx, const double* displacements, const double* fe, double dt, double* send_host) {
  if (!ctx || !ctx->shape_ok || !displacements || (ctx->halo_count && !send_host)) return fail(ctx, FTB200_ERR_INPUT, "get_force_begin: bad arguments");
  CK(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = ensure_xwin(ctx)) || (rc = legacy_force_local(ctx, displacements, fe, dt))) return rc;
  if (ctx->halo_count) {
    LAUNCH(k_gather_shared, cdiv(ctx->nshared, 128), 128, ctx->stream, ctx->felem, ctx->node_off, ctx->node_ent, ctx->halo_nodes,
           ctx->fnet[0], ctx->fnet[1], ctx->fnet[2], ctx->nshared, ctx->nE);
    LAUNCH(k_halo_pack, cdiv(ctx->halo_count, 256), 256, ctx->stream, ctx->fnet[0], ctx->fnet[1], ctx->fnet[2], ctx->d_sendNodeIndex,
           ctx->d_xsend, ctx->halo_count);
  }
  return xsend_to_host(ctx, send_host);
}
// ... and, once the host has exchanged them, the assembled fi (+ neighbours in ascending neighbour order) and f_net
int ftb200_get_force_end(ftb200_ctx* ctx, const double* recv_host, double* fi, double* f_net) {
  if (!ctx || !ctx->shape_ok || (ctx->halo_count && !recv_host)) return fail(ctx, FTB200_ERR_INPUT, "get_force_end: bad arguments");
  CK(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = xrecv_from_host(ctx, recv_host))) return rc;
  const NodeArgs N = node_args(ctx, ctx->halo_count ? ctx->d_xrecv : nullptr);
  LAUNCH(k_gather_force, ctx->node_blocks, NODE_BLOCK, ctx->stream, N, ctx->fnet[0], ctx->fnet[1], ctx->fnet[2]);
  ctx->halo_recv_cur = ctx->halo_count ? ctx->d_xrecv : nullptr;
  if (fi && (rc = download_aos(ctx, ctx->fi, fi, 0))) return rc;
  if (f_net && (rc = download_aos(ctx, ctx->fnet, f_net, 1))) return rc;
  CK(cudaGetLastError());
  int status = 0;
  CK(cudaMemcpyAsync(&status, &ctx->sc->status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (status & 1) return fail(ctx, FTB200_ERR_MATERIAL, "Unknown material type");
  return FTB200_OK;
}
// the cross-rank MIN of the stable time step: read this rank's candidate, write the global one (StableTimeStep.cpp:33)
int ftb200_get_dtmin(ftb200_ctx* ctx, double* dtmin_local) {
  if (!ctx || !ctx->shape_ok || !dtmin_local) return fail(ctx, FTB200_ERR_INPUT, "get_dtmin: bad arguments");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));  // the interior elements of a split step (no-op otherwise)
  CK(cudaMemcpyAsync(dtmin_local, &ctx->sc->dtmin_bits, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return FTB200_OK;
}
int ftb200_set_dtmin(ftb200_ctx* ctx, double dtmin_global) {
  if (!ctx || !ctx->shape_ok) return fail(ctx, FTB200_ERR_INPUT, "set_dtmin: bad arguments");
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(&ctx->sc->dtmin_bits, &dtmin_global, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));  // the source is the caller's stack
  return FTB200_OK;
}
int ftb200_explicit_begin_force_host(ftb200_ctx* ctx, double* send_host) {
  if (!ctx || !ctx->shape_ok || (ctx->halo_count && !send_host)) return fail(ctx, FTB200_ERR_INPUT, "explicit_begin_force_host: bad arguments");
  int rc;
  if ((rc = ensure_xwin(ctx)) || (rc = ftb200_explicit_begin_force(ctx, ctx->halo_count ? ctx->d_xsend : nullptr))) return rc;
  return xsend_to_host(ctx, send_host);
}
int ftb200_explicit_begin_finish_host(ftb200_ctx* ctx, const double* recv_host) {
  if (!ctx || !ctx->shape_ok || (ctx->halo_count && !recv_host)) return fail(ctx, FTB200_ERR_INPUT, "explicit_begin_finish_host: bad arguments");
  int rc;
  if ((rc = xrecv_from_host(ctx, recv_host))) return rc;
  return ftb200_explicit_begin_finish(ctx, ctx->halo_count ? ctx->d_xrecv : nullptr);
}
// one step of the loop with the exchange on the host: elements + packed partials out, this rank's dt candidate out ...
int ftb200_step_begin_host(ftb200_ctx* ctx, double* send_host, double* dtmin_local) {
  if (!ctx || !ctx->begun || (ctx->halo_count && !send_host) || !dtmin_local) return fail(ctx, FTB200_ERR_INPUT, "step_begin_host: bad arguments");
  int rc;
  if ((rc = ensure_xwin(ctx)) || (rc = ftb200_step_begin(ctx, ctx->halo_count ? ctx->d_xsend : nullptr, nullptr)) ||
      (rc = xsend_to_host(ctx, send_host)))
    return rc;
  return ftb200_get_dtmin(ctx, dtmin_local);
}
// ... neighbours' partials and the global dt in, scalar update + node kernel
int ftb200_step_end_host(ftb200_ctx* ctx, const double* recv_host, double dtmin_global) {
  if (!ctx || !ctx->begun || (ctx->halo_count && !recv_host)) return fail(ctx, FTB200_ERR_INPUT, "step_end_host: bad arguments");
  int rc;
  if ((rc = ftb200_set_dtmin(ctx, dtmin_global)) || (rc = xrecv_from_host(ctx, recv_host))) return rc;
  return ftb200_step_end(ctx, ctx->halo_count ? ctx->d_xrecv : nullptr);
}

int ftb200_p2p_export(ftb200_ctx* ctx, void* handle_out, void** window_out) {
  if (!ctx || !ctx->shape_ok) return fail(ctx, FTB200_ERR_INPUT, "p2p_export: call shape_functions first");
  CK(cudaSetDevice(ctx->device));
  if ((int)ctx->h_sendProcessID.size() > P2P_MAXNB || ctx->nranks > P2P_MAXP)
    return fail(ctx, FTB200_ERR_INPUT, "p2p_export: more than %d neighbours or %d ranks", P2P_MAXNB, P2P_MAXP);
  if (!ctx->p2p_window) {
    ctx->p2p_bytes = p2p_recv_off(ctx->nranks) + 2 * 3 * (size_t)std::max(ctx->halo_count, 1) * sizeof(double);
    CK(cudaMalloc((void**)&ctx->p2p_window, ctx->p2p_bytes));
    CK(ftb_memset(ctx, ctx->p2p_window, 0, ctx->p2p_bytes));
    {  // the dt slots start out empty (P2PHeader)
      std::vector<unsigned long long> empty(4 * P2P_MAXP, P2P_DT_EMPTY);
      CK(cudaMemcpyAsync(ctx->p2p_window + offsetof(P2PHeader, dtslot), empty.data(), empty.size() * sizeof(unsigned long long),
                         cudaMemcpyHostToDevice, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
    }
    int rc;
    if ((rc = dalloc(ctx, &ctx->d_seq, 1)) || (rc = dalloc(ctx, &ctx->d_p2p_blocks, 1))) return rc;
    CK(ftb_memset(ctx, ctx->d_seq, 0, sizeof(unsigned long long)));
    CK(ftb_memset(ctx, ctx->d_p2p_blocks, 0, sizeof(unsigned)));
    CK(cudaDeviceSynchronize());
  }
  if (handle_out) {
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ctx->p2p_window));
    static_assert(sizeof(cudaIpcMemHandle_t) == FTB200_IPC_HANDLE_BYTES, "IPC handle size");
    memcpy(handle_out, &h, sizeof(h));
  }
  if (window_out) *window_out = ctx->p2p_window;
  return FTB200_OK;
}

int ftb200_p2p_import(ftb200_ctx* ctx, const void* all_handles, int handles_are_pointers, const int* peer_slot_offset,
                      const int* peer_my_index, const int* peer_halo_count) {
  if (!ctx || !ctx->p2p_window || !all_handles) return fail(ctx, FTB200_ERR_INPUT, "p2p_import: call p2p_export first");
  const int nnb = (int)ctx->h_sendProcessID.size();
  if (nnb && (!peer_slot_offset || !peer_my_index || !peer_halo_count)) return fail(ctx, FTB200_ERR_INPUT, "p2p_import: null metadata");
  CK(cudaSetDevice(ctx->device));
  // a loop graph captured after an earlier import has the old windows and the old PackArgs block baked into its kernel
  // arguments: drop it (the next run captures again)
  CK(cudaStreamSynchronize(ctx->stream));
  drop_graphs(ctx);
  P2PArgs& P = ctx->p2p;
  memset(&P, 0, sizeof(P));
  P.self = ctx->p2p_window;
  P.n_nb = nnb; P.n_ranks = ctx->nranks; P.rank = ctx->rank; P.H = ctx->halo_count;
  P.seq = ctx->d_seq; P.blocks_done = ctx->d_p2p_blocks;
  for (int r = 0; r < ctx->nranks; ++r) {
    if (r == ctx->rank) { P.peer_rank[r] = ctx->p2p_window; continue; }
    if (handles_are_pointers) {
      P.peer_rank[r] = (char*)((void* const*)all_handles)[r];
    } else {
      cudaIpcMemHandle_t h;
      memcpy(&h, (const char*)all_handles + (size_t)r * FTB200_IPC_HANDLE_BYTES, sizeof(h));
      void* q = nullptr;
      CK(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
      ctx->p2p_opened.push_back(q);
      P.peer_rank[r] = (char*)q;
    }
  }
  for (int i = 0; i < nnb; ++i) {
    const int q = ctx->h_sendProcessID[i];
    if (q < 0 || q >= ctx->nranks) return fail(ctx, FTB200_ERR_INPUT, "p2p_import: neighbour rank %d out of range", q);
    P.peer_nb[i] = P.peer_rank[q];
    P.peer_slot_off[i] = peer_slot_offset[i];
    P.peer_my_index[i] = peer_my_index[i];
    P.peer_H[i] = peer_halo_count[i];
    P.nb_cum[i] = ctx->h_sendCum[i];
  }
  P.nb_cum[nnb] = ctx->h_sendCum[nnb];
  // ---- exchange fused into the element kernels (all-hexahedra meshes; FTB200_P2P_FUSED=0: separate pack kernel) ----
  ctx->p2p_fused = !ctx->has_tet;
  if (const char* ev = getenv("FTB200_P2P_FUSED")) ctx->p2p_fused = ctx->p2p_fused && atoi(ev) != 0;
  if (ctx->p2p_fused) {
    PackArgs K;
    memset(&K, 0, sizeof(K));
    K.P = P;
    K.halo_node_idx = ctx->halo_node_idx; K.halo_off = ctx->halo_off; K.halo_slot = ctx->halo_slot;
    K.node_off = ctx->node_off; K.node_ent = ctx->node_ent;
    int rc;
    if ((rc = dalloc(ctx, &ctx->d_pk_ctr, (size_t)ctx->nshared + 2)) || (rc = dalloc(ctx, &ctx->d_pk, 1))) return rc;
    CK(ftb_memset(ctx, ctx->d_pk_ctr, 0, ((size_t)ctx->nshared + 2) * sizeof(unsigned)));
    K.node_ctr = ctx->d_pk_ctr + 2; K.packed = ctx->d_pk_ctr;
    K.n_shared = ctx->nshared; K.nEb = ctx->nE_boundary;
    CK(ftb_memcpy(ctx, ctx->d_pk, &K, sizeof(K), cudaMemcpyHostToDevice));
    CK(cudaStreamSynchronize(ctx->stream));  // K lives on this stack frame
  }
  // Load every kernel of the loop NOW: with lazy module loading the first launch of a kernel may
  // synchronise the context, which would deadlock against a peer's spinning wait kernel when several
  // ranks live in one process.
  {
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, k_p2p_pack));
    CK(cudaFuncGetAttributes(&fa, k_adv_p2p));
    CK(cudaFuncGetAttributes(&fa, k_injury_xchg));
    CK(cudaFuncGetAttributes(&fa, k_injury_pick));
    CK(cudaFuncGetAttributes(&fa, k_injury_select));
    CK(cudaFuncGetAttributes(&fa, k_injury_reduce));
    CK(cudaFuncGetAttributes(&fa, k_injury_lists));
    CK(cudaFuncGetAttributes(&fa, k_begin_run));
    CK(cudaFuncGetAttributes(&fa, k_energy));
    CK(cudaFuncGetAttributes(&fa, k_gather_force));
    CK(cudaFuncGetAttributes(&fa, k_node<true, true, true, true>));
    CK(cudaFuncGetAttributes(&fa, k_node<true, true, true, false>));
    CK(cudaFuncGetAttributes(&fa, k_node<false, true, false, true>));
    CK(cudaFuncGetAttributes(&fa, k_node<false, true, false, false>));
    CK(cudaFuncGetAttributes(&fa, k_elem<1, true, true>));
    CK(cudaFuncGetAttributes(&fa, k_elem<4, true, true>));
    CK(cudaFuncGetAttributes(&fa, k_elem<5, true, true>));
    CK(cudaFuncGetAttributes(&fa, k_elem<-1, true, true>));
    CK(cudaFuncGetAttributes(&fa, k_elem_affine_cj<1>));
    CK(cudaFuncGetAttributes(&fa, k_elem_affine_cj<4>));
    CK(cudaFuncGetAttributes(&fa, (k_elem_affine_cj<1, true>)));
    CK(cudaFuncGetAttributes(&fa, (k_elem_affine_cj<4, true>)));
    CK(cudaFuncGetAttributes(&fa, (k_elem_affine<1, false>)));
    CK(cudaFuncGetAttributes(&fa, (k_elem_affine<4, false>)));
    CK(cudaFuncGetAttributes(&fa, (k_elem_affine<5, false>)));
    CK(cudaFuncGetAttributes(&fa, k_stamp));
  }
  ctx->p2p_ready = true;
  return FTB200_OK;
}

// ------------------------------------------------------------------------------------ measurement
int ftb200_profile_enable(ftb200_ctx* ctx, int on) {
  if (!ctx) return FTB200_ERR_INPUT;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->profile = on != 0;
  ctx->prof_elem_ms = ctx->prof_node_ms = 0;
  ctx->prof_elem_n = ctx->prof_node_n = 0;
  ctx->prof.elem.clear(); ctx->prof.node.clear(); ctx->prof.used = 0;
  return FTB200_OK;
}
int ftb200_profile_get(ftb200_ctx* ctx, double* elem_ms, double* node_ms, long long* elem_launches, long long* node_launches) {
  if (!ctx) return FTB200_ERR_INPUT;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  prof_collect(ctx);
  if (elem_ms) *elem_ms = ctx->prof_elem_n ? ctx->prof_elem_ms / ctx->prof_elem_n : 0.0;
  if (node_ms) *node_ms = ctx->prof_node_n ? ctx->prof_node_ms / ctx->prof_node_n : 0.0;
  if (elem_launches) *elem_launches = ctx->prof_elem_n;
  if (node_launches) *node_launches = ctx->prof_node_n;
  return FTB200_OK;
}

// ----------------------------------------------------------------------------------- injury criteria
// ----------------------------------------------------------------------------------- rigid-body BC
__global__ void k_rigid_mark(const int* __restrict__ ids, int n, const int* __restrict__ nint, uint16_t* flags, double* ux,
                             double* uy, double* uz, double* vx, double* vy, double* vz, double* ax, double* ay, double* az) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = nint[ids[i]];
  flags[k] = (uint16_t)(flags[k] | 7u | FTB_FLAG_RIGID);  // boundary on x, y, z (ex5.cpp:849-854)
  ux[k] = uy[k] = uz[k] = 0.0; vx[k] = vy[k] = vz[k] = 0.0; ax[k] = ay[k] = az[k] = 0.0;  // :901-908
}

int ftb200_set_rigid_bc(ftb200_ctx* ctx, const int sizes[6], const double* const t[6], const double* const v[6],
                        const int* boundaryID, int boundarySize) {
  if (!ctx || !ctx->shape_ok || !sizes || !t || !v) return fail(ctx, FTB200_ERR_INPUT, "set_rigid_bc: setup incomplete or bad arguments");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  DevRigid h;
  memset(&h, 0, sizeof(h));
  int total = 0;
  for (int k = 0; k < 6; ++k) {
    if (sizes[k] < 2 || !t[k] || !v[k]) return fail(ctx, FTB200_ERR_INPUT, "set_rigid_bc: trace %d needs at least two points", k);
    h.size[k] = sizes[k]; h.off[k] = total; total += sizes[k];
  }
  std::vector<double> tab(2 * (size_t)total);
  for (int k = 0; k < 6; ++k)
    for (int i = 0; i < sizes[k]; ++i) { tab[h.off[k] + i] = t[k][i]; tab[total + h.off[k] + i] = v[k][i]; }
  // node set: given, or the nodes of the elements of rigid parts (material 0), ex5.cpp:819-846
  std::vector<int> ids;
  if (boundaryID) {
    ids.assign(boundaryID, boundaryID + boundarySize);
    for (int id : ids) if (id < 0 || id >= ctx->nN) return fail(ctx, FTB200_ERR_INPUT, "set_rigid_bc: node %d out of range", id);
  } else {
    std::vector<uint8_t> mark(ctx->nN, 0);
    for (int e = 0; e < ctx->nE; ++e)
      if (ctx->h_matid[ctx->h_pid[e]] == 0)
        for (int k = 0; k < 8; ++k) mark[ctx->h_conn[8 * (size_t)e + k]] = 1;
    for (int n = 0; n < ctx->nN; ++n) if (mark[n]) ids.push_back(n);
  }
  int rc;
  dfree(ctx->rigid_tab);
  if ((rc = dalloc(ctx, &ctx->rigid_tab, 2 * (size_t)total))) return rc;
  CK(ftb_memcpy(ctx, ctx->rigid_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
  h.tab_t = ctx->rigid_tab; h.tab_v = ctx->rigid_tab + total;
  h.R[0] = h.Rinv[0] = 1.0;
  if (!ctx->rigid && (rc = dalloc(ctx, &ctx->rigid, 1))) return rc;
  CK(ftb_memcpy(ctx, ctx->rigid, &h, sizeof(h), cudaMemcpyHostToDevice));
  for (int k = 0; k < 3; ++k) {
    if (!ctx->aprev[k] && (rc = dalloc(ctx, &ctx->aprev[k], (size_t)ctx->nNp))) return rc;
    CK(ftb_memset(ctx, ctx->aprev[k], 0, (size_t)ctx->nNp * sizeof(double)));
  }
  ctx->rigid_count = (int)ids.size();
  if (!ids.empty()) {
    if ((rc = ensure_big(ctx, (ids.size() + (size_t)ctx->nN) * sizeof(int) + 16))) return rc;
    int* d_ids = reinterpret_cast<int*>(ctx->d_big);
    int* d_nint = d_ids + ids.size();
    CK(ftb_memcpy(ctx, d_ids, ids.data(), ids.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(ftb_memcpy(ctx, d_nint, ctx->h_nint.data(), (size_t)ctx->nN * sizeof(int), cudaMemcpyHostToDevice));
    LAUNCH(k_rigid_mark, cdiv((int)ids.size(), 256), 256, ctx->stream, d_ids, (int)ids.size(), d_nint, ctx->flags, ctx->u[0], ctx->u[1],
           ctx->u[2], ctx->v[0], ctx->v[1], ctx->v[2], ctx->a[0], ctx->a[1], ctx->a[2]);
    LAUNCH(k_eflag, cdiv(ctx->nE, 256), 256, ctx->stream, ctx->conn, ctx->flags, ctx->eflag, ctx->nE);
    CK(cudaStreamSynchronize(ctx->stream));
  }
  drop_graphs(ctx);
  ctx->bc_ok = true;
  return FTB200_OK;
}

int ftb200_get_rigid_state(ftb200_ctx* ctx, double* y12, double* ydot12, int* boundary_count) {
  if (!ctx || !ctx->rigid) return fail(ctx, FTB200_ERR_INPUT, "get_rigid_state: call set_rigid_bc first");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  DevRigid h;
  CK(ftb_memcpy(ctx, &h, ctx->rigid, sizeof(h), cudaMemcpyDeviceToHost));
  if (y12) for (int i = 0; i < 12; ++i) y12[i] = h.y[i];
  if (ydot12) for (int i = 0; i < 12; ++i) ydot12[i] = h.ydot[i];
  if (boundary_count) *boundary_count = ctx->rigid_count;
  return FTB200_OK;
}


int ftb200_injury_begin(ftb200_ctx* ctx, const int* exclude_pids, int n_exclude, const double* thresholds4) {
  if (!ctx || !ctx->shape_ok || n_exclude < 0 || (n_exclude && !exclude_pids))
    return fail(ctx, FTB200_ERR_INPUT, "injury_begin: setup incomplete or bad arguments");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  const size_t nE = ctx->nE;
  // InitInjuryCriterion (ex5.cpp:1251-1281): elements whose part is not excluded, internal element order
  std::vector<int> ref_of(nE);
  CK(ftb_memcpy(ctx, ref_of.data(), ctx->ref_of, nE * sizeof(int), cudaMemcpyDeviceToHost));
  std::vector<uint8_t> incl(nE, 1);
  int n = 0;
  for (size_t t = 0; t < nE; ++t) {
    const int pide = ctx->h_pid[ref_of[t]];
    for (int j = 0; j < n_exclude; ++j)
      if (pide == exclude_pids[j]) { incl[t] = 0; break; }
    n += incl[t];
  }
  const int index95 = (int)(n * 0.95) - 1;  // math.cpp:189; the reference faults on a negative index
  if (index95 < 0 && ctx->nranks == 1)
    return fail(ctx, FTB200_ERR_INPUT, "injury_begin: %d participating elements, the 95th percentile needs >= 2", n);
  int rc;
  if (!ctx->inj_ps) {
    if ((rc = dalloc(ctx, &ctx->inj_ps, nE)) || (rc = dalloc(ctx, &ctx->inj_psxsr, nE)) || (rc = dalloc(ctx, &ctx->inj_smin, nE)) ||
        (rc = dalloc(ctx, &ctx->inj_shear, nE)) || (rc = dalloc(ctx, &ctx->inj_flags, nE)) || (rc = dalloc(ctx, &ctx->inj_incl, nE)) ||
        (rc = dalloc(ctx, &ctx->inj_part, 4 * (size_t)INJ_BLOCKS)) || (rc = dalloc(ctx, &ctx->inj_parti, 4 * (size_t)INJ_BLOCKS)) ||
        (rc = dalloc(ctx, &ctx->inj_state, 1)))
      return rc;
  }
  // PS_Old starts at zero (the reference mallocs it uninitialised, ex5.cpp:1285; fresh pages read as zero)
  CK(ftb_memset(ctx, ctx->inj_ps, 0, nE * sizeof(double)));
  CK(ftb_memset(ctx, ctx->inj_psxsr, 0, nE * sizeof(double)));
  CK(ftb_memset(ctx, ctx->inj_smin, 0, nE * sizeof(double)));
  CK(ftb_memset(ctx, ctx->inj_shear, 0, nE * sizeof(double)));
  CK(ftb_memset(ctx, ctx->inj_flags, 0, nE));
  CK(ftb_memcpy(ctx, ctx->inj_incl, incl.data(), nE, cudaMemcpyHostToDevice));
  InjState st;
  memset(&st, 0, sizeof(st));
  st.kth0 = (unsigned)std::max(index95, 0);  // several partitions: ftb200_injury_global_count sets the global rank
  st.nIncluded = n;
  CK(ftb_memcpy(ctx, ctx->inj_state, &st, sizeof(st), cudaMemcpyHostToDevice));
  if (thresholds4) for (int k = 0; k < 4; ++k) ctx->inj_thr[k] = thresholds4[k];
  else { ctx->inj_thr[0] = 0.15; ctx->inj_thr[1] = 0.30; ctx->inj_thr[2] = 120.0; ctx->inj_thr[3] = 28.0; }
  dfree(ctx->inj_hist);
  if (ctx->hist_cap > 0) {
    if ((rc = dalloc(ctx, &ctx->inj_hist, 2 * (size_t)ctx->hist_cap))) return rc;
    CK(ftb_memset(ctx, ctx->inj_hist, 0, 2 * ctx->hist_cap * sizeof(double)));
  }
  ctx->injury = true;
  drop_graphs(ctx);
  return FTB200_OK;
}

int ftb200_injury_end(ftb200_ctx* ctx) {
  if (!ctx) return FTB200_ERR_INPUT;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->injury = false;
  drop_graphs(ctx);
  return FTB200_OK;
}

int ftb200_principal_strains(ftb200_ctx* ctx, double* smax, double* smin, double* shear, double* volume0) {
  if (!ctx || !ctx->shape_ok) return fail(ctx, FTB200_ERR_INPUT, "principal_strains: setup incomplete");
  if ((smax || smin || shear) && !(smax && smin && shear)) return fail(ctx, FTB200_ERR_INPUT, "principal_strains: smax, smin, shear go together");
  CK(cudaSetDevice(ctx->device));
  const size_t nE = ctx->nE;
  int rc;
  if ((rc = ensure_big(ctx, 4 * nE * sizeof(double)))) return rc;
  double* d = ctx->d_big;
  LAUNCH(k_principal, cdiv(ctx->nE, 64), 64, ctx->stream, elem_args(ctx, 0, ctx->nE, 1), ctx->ref_of, smax ? d : nullptr, d + nE,
         d + 2 * nE, volume0 ? d + 3 * nE : nullptr);
  if (smax) {
    CK(cudaMemcpyAsync(smax, d, nE * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(smin, d + nE, nE * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(shear, d + 2 * nE, nE * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (volume0) CK(cudaMemcpyAsync(volume0, d + 3 * nE, nE * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return FTB200_OK;
}

int ftb200_injury_get(ftb200_ctx* ctx, double* scalars12, int* extreme_elems4, unsigned char* flags, double* ps, double* psxsr,
                      double* volumes5) {
  if (!ctx || !ctx->inj_state) return fail(ctx, FTB200_ERR_INPUT, "injury_get: call injury_begin first");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  InjState st;
  CK(ftb_memcpy(ctx, &st, ctx->inj_state, sizeof(st), cudaMemcpyDeviceToHost));
  if (scalars12) for (int k = 0; k < 12; ++k) scalars12[k] = st.scal[k];
  if (extreme_elems4) for (int k = 0; k < 4; ++k) extreme_elems4[k] = st.elems[k];
  if (!(flags || ps || psxsr || volumes5)) return FTB200_OK;
  const size_t nE = ctx->nE;
  std::vector<int> ref_of(nE);
  std::vector<uint8_t> f(nE), incl(nE);
  CK(ftb_memcpy(ctx, ref_of.data(), ctx->ref_of, nE * sizeof(int), cudaMemcpyDeviceToHost));
  CK(ftb_memcpy(ctx, f.data(), ctx->inj_flags, nE, cudaMemcpyDeviceToHost));
  CK(ftb_memcpy(ctx, incl.data(), ctx->inj_incl, nE, cudaMemcpyDeviceToHost));
  std::vector<uint8_t> fr(nE);
  for (size_t t = 0; t < nE; ++t) fr[ref_of[t]] = (uint8_t)(f[t] | (incl[t] ? 0x80u : 0u));
  if (flags) memcpy(flags, fr.data(), nE);
  std::vector<double> tmp(nE);
  if (ps) {
    CK(ftb_memcpy(ctx, tmp.data(), ctx->inj_ps, nE * sizeof(double), cudaMemcpyDeviceToHost));
    for (size_t t = 0; t < nE; ++t) ps[ref_of[t]] = tmp[t];
  }
  if (psxsr) {
    CK(ftb_memcpy(ctx, tmp.data(), ctx->inj_psxsr, nE * sizeof(double), cudaMemcpyDeviceToHost));
    for (size_t t = 0; t < nE; ++t) psxsr[ref_of[t]] = tmp[t];
  }
  if (volumes5) {  // ex5.cpp:1049-1066: summed in the reference's element order
    int rc = ftb200_principal_strains(ctx, nullptr, nullptr, nullptr, tmp.data());
    if (rc) return rc;
    for (int k = 0; k < 5; ++k) volumes5[k] = 0.0;
    for (size_t e = 0; e < nE; ++e) {
      if (!(fr[e] & 0x80u)) continue;
      const double eV = tmp[e];
      if (fr[e] & FTB_INJ_MPS_LO) { volumes5[0] += eV; if (fr[e] & FTB_INJ_MPS_HI) volumes5[1] += eV; }
      if (fr[e] & FTB_INJ_PSR) volumes5[2] += eV;
      if (fr[e] & FTB_INJ_PSXSR) volumes5[3] += eV;
      volumes5[4] += eV;
    }
  }
  return FTB200_OK;
}

int ftb200_injury_history(ftb200_ctx* ctx, long long first, long long count, double* mps95, double* mpsxsr95) {
  if (!ctx || !ctx->inj_hist || first < 0 || count < 0 || first + count > ctx->hist_cap)
    return fail(ctx, FTB200_ERR_INPUT, "injury_history: no history recorded (record_history before injury_begin) or bad range");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  if (mps95 && count) CK(ftb_memcpy(ctx, mps95, ctx->inj_hist + first, count * sizeof(double), cudaMemcpyDeviceToHost));
  if (mpsxsr95 && count) CK(ftb_memcpy(ctx, mpsxsr95, ctx->inj_hist + ctx->hist_cap + first, count * sizeof(double), cudaMemcpyDeviceToHost));
  return FTB200_OK;
}

// ---- several partitions: the percentile is a global order statistic (math.cpp:160-199 gathers all ranks) -----
int ftb200_injury_local_count(ftb200_ctx* ctx, long long* n_included) {
  if (!ctx || !ctx->inj_state || !n_included) return fail(ctx, FTB200_ERR_INPUT, "injury_local_count: call injury_begin first");
  CK(cudaSetDevice(ctx->device));
  InjState st;
  CK(ftb_memcpy(ctx, &st, ctx->inj_state, sizeof(st), cudaMemcpyDeviceToHost));
  *n_included = st.nIncluded;
  return FTB200_OK;
}
int ftb200_injury_global_count(ftb200_ctx* ctx, long long n_total) {
  if (!ctx || !ctx->inj_state) return fail(ctx, FTB200_ERR_INPUT, "injury_global_count: call injury_begin first");
  const long long index95 = (long long)(n_total * 0.95) - 1;
  if (index95 < 0) return fail(ctx, FTB200_ERR_INPUT, "injury_global_count: %lld participating elements, the 95th percentile needs >= 2", n_total);
  CK(cudaSetDevice(ctx->device));
  const unsigned k = (unsigned)index95;
  CK(ftb_memcpy(ctx, &ctx->inj_state->kth0, &k, sizeof(k), cudaMemcpyHostToDevice));
  return FTB200_OK;
}
int ftb200_injury_select_hist(ftb200_ctx* ctx, int pass, unsigned** hist_dev, int* hist_len) {
  if (!ctx || !ctx->injury || pass < 0 || pass >= INJ_PASSES) return fail(ctx, FTB200_ERR_INPUT, "injury_select_hist: bad arguments");
  CK(cudaSetDevice(ctx->device));
  const ElemArgs A = elem_args(ctx, 0, ctx->nE, 0);
  LAUNCH(k_injury_select, dim3(std::min(INJ_BLOCKS, cdiv(ctx->nE, INJ_THREADS * INJ_ITEMS)), 2), INJ_THREADS, ctx->stream, A, ctx->inj_state,
         pass, nullptr, nullptr, 1);
  if (hist_dev) *hist_dev = &ctx->inj_state->hist[0][0];
  if (hist_len) *hist_len = 2 * INJ_BINS;
  return FTB200_OK;
}
int ftb200_injury_select_pick(ftb200_ctx* ctx, int pass) {
  if (!ctx || !ctx->injury || pass < 0 || pass >= INJ_PASSES) return fail(ctx, FTB200_ERR_INPUT, "injury_select_pick: bad arguments");
  CK(cudaSetDevice(ctx->device));
  double* h0 = ctx->inj_hist;
  double* h1 = ctx->inj_hist ? ctx->inj_hist + ctx->hist_cap : nullptr;
  LAUNCH(k_injury_pick, dim3(1, 2), INJ_THREADS, ctx->stream, ctx->sc, ctx->inj_state, pass, h0, h1);
  if (pass == INJ_PASSES - 1) LAUNCH(k_injury_lists, cdiv(ctx->nE, 256), 256, ctx->stream, elem_args(ctx, 0, ctx->nE, 0), ctx->inj_state);
  return FTB200_OK;
}
int ftb200_injury_passes(void) { return INJ_PASSES; }

int ftb200_measure_peaks(ftb200_ctx* ctx, int reps, double* fp64_tflops, double* copy_gbs) {
  if (!ctx) return FTB200_ERR_INPUT;
  CK(cudaSetDevice(ctx->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, ctx->device));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  cudaStream_t s = ctx->stream;
  if (reps < 1) reps = 1;
  if (fp64_tflops) {
    double* d = nullptr;
    CK(cudaMalloc((void**)&d, 64));
    const int iters = 1 << 14, blocks = prop.multiProcessorCount * 8;
    double best = 0;
    for (int r = 0; r < reps + 1; ++r) {
      CK(cudaEventRecord(e0, s));
      LAUNCH(k_dfma_peak, blocks, 256, s, d, iters, 1.0);
      CK(cudaEventRecord(e1, s));
      CK(cudaEventSynchronize(e1));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double tf = 2.0 * 8.0 * iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
      if (r > 0 && tf > best) best = tf;
    }
    cudaFree(d);
    *fp64_tflops = best;
  }
  if (copy_gbs) {
    const size_t n2 = (size_t)1 << 26;  // 64 Mi double2 = 1 GiB per buffer, far larger than L2
    double2 *a = nullptr, *b = nullptr;
    CK(cudaMalloc((void**)&a, n2 * sizeof(double2)));
    CK(cudaMalloc((void**)&b, n2 * sizeof(double2)));
    CK(cudaMemsetAsync(a, 0, n2 * sizeof(double2), s));
    double best = 0;
    for (int r = 0; r < reps + 1; ++r) {
      CK(cudaEventRecord(e0, s));
      LAUNCH(k_copy_peak, prop.multiProcessorCount * 16, 256, s, a, b, n2);
      CK(cudaEventRecord(e1, s));
      CK(cudaEventSynchronize(e1));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double gbs = 2.0 * n2 * sizeof(double2) / (ms * 1e-3) / 1e9;
      if (r > 0 && gbs > best) best = gbs;
    }
    cudaFree(a);
    cudaFree(b);
    *copy_gbs = best;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return FTB200_OK;
}

}  // extern "C"
