/*
 * ftmpi.c -- process-per-rank MPI shim over POSIX shared memory.
 * TEST INFRASTRUCTURE ONLY (oracle harness; never part of the product).
 *
 * Why it exists: neither this container nor the GPU box has an MPI
 * installation, but (a) bit-exact multi-rank ParMETIS partitions / node maps
 * and (b) a multi-core CPU baseline of the reference both need the reference
 * to run on P > 1 ranks.
 *
 * Model: `ftmpirun -np P prog args...` creates one sparse file in /dev/shm,
 * forks P children with FTMPI_RANK / FTMPI_SIZE / FTMPI_SHM set, and waits.
 * Without those variables a process is a 1-rank world (private memory).
 * Every ordered pair (src,dst) owns a byte ring; sends are eager copies into
 * the ring, receives drain the ring into a per-source unexpected-message list
 * and match on (communicator context, tag) in FIFO order, which gives MPI's
 * non-overtaking semantics.  MPI_ANY_SOURCE is not supported (the reference
 * and ParMETIS never use it).  Collectives are linear, rooted at rank 0 and
 * combine contributions in ascending rank order, so results are deterministic.
 */
#define _GNU_SOURCE
#include "mpi.h"

#include <errno.h>
#include <fcntl.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdint.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <sys/types.h>
#include <unistd.h>

#define FT_MAXP 64
#define FT_MAXCOMM 4096
#define FT_MAXREQ 65536
#define FT_HDR 32

typedef struct {
  _Atomic uint64_t head; /* consumer position */
  char pad0[56];
  _Atomic uint64_t tail; /* producer position */
  char pad1[56];
} ft_ring_ctl;

typedef struct {
  _Atomic int abort_flag;
  int abort_code;
  char pad[56];
} ft_global;

typedef struct ft_msg {
  int tag, ctx;
  size_t bytes;
  struct ft_msg *next;
  char data[];
} ft_msg;

typedef struct {
  int used, ctx, size, rank;
  int *wr; /* world ranks of members */
} ft_comm;

typedef struct {
  int active; /* 0 free, 1 pending recv, 2 complete */
  void *buf;
  size_t bytes;
  int src_world, tag, ctx;
  MPI_Status st;
} ft_req;

static int g_rank = 0, g_size = 1, g_inited = 0;
static size_t g_ring = 0;
static char *g_base = NULL;
static ft_global *g_glob = NULL;
static ft_msg *g_unexp_head[FT_MAXP], *g_unexp_tail[FT_MAXP];
static ft_comm g_comm[FT_MAXCOMM];
static ft_req g_req[FT_MAXREQ];
static int g_next_ctx = 1;

static size_t chan_stride(void) { return sizeof(ft_ring_ctl) + g_ring; }
static ft_ring_ctl *chan_ctl(int src, int dst) {
  return (ft_ring_ctl *)(g_base + 4096 + chan_stride() * ((size_t)src * g_size + dst));
}
static char *chan_data(int src, int dst) { return (char *)chan_ctl(src, dst) + sizeof(ft_ring_ctl); }

static void ft_die(const char *m) {
  fprintf(stderr, "[ftmpi %d] fatal: %s\n", g_rank, m);
  if (g_glob) { g_glob->abort_code = 86; atomic_store(&g_glob->abort_flag, 1); }
  _exit(86);
}
static void check_abort(void) {
  if (g_glob && atomic_load_explicit(&g_glob->abort_flag, memory_order_relaxed)) _exit(g_glob->abort_code ? g_glob->abort_code : 1);
}

static int dsize(MPI_Datatype t) {
  switch (t) {
    case MPI_CHAR: case MPI_BYTE: return 1;
    case MPI_SHORT: return 2;
    case MPI_INT: case MPI_FLOAT: case MPI_UNSIGNED: return 4;
    case MPI_LONG: case MPI_DOUBLE: case MPI_LONG_LONG_INT: case MPI_UNSIGNED_LONG: return 8;
    case MPI_DOUBLE_INT: return 16; /* struct {double; int;} padded */
    case MPI_2INT: case MPI_FLOAT_INT: return 8;
    default: ft_die("unknown datatype"); return 0;
  }
}

static void ring_copy_in(char *ring, uint64_t pos, const void *src, size_t n) {
  size_t off = pos % g_ring, first = g_ring - off;
  if (first >= n) memcpy(ring + off, src, n);
  else { memcpy(ring + off, src, first); memcpy(ring, (const char *)src + first, n - first); }
}
static void ring_copy_out(const char *ring, uint64_t pos, void *dst, size_t n) {
  size_t off = pos % g_ring, first = g_ring - off;
  if (first >= n) memcpy(dst, ring + off, n);
  else { memcpy(dst, ring + off, first); memcpy((char *)dst + first, ring, n - first); }
}

/* drain one message from src -> me into the unexpected list; 1 if drained */
static int progress_from(int src) {
  ft_ring_ctl *c = chan_ctl(src, g_rank);
  uint64_t head = atomic_load_explicit(&c->head, memory_order_relaxed);
  uint64_t tail = atomic_load_explicit(&c->tail, memory_order_acquire);
  if (head == tail) return 0;
  char hdr[FT_HDR];
  ring_copy_out(chan_data(src, g_rank), head, hdr, FT_HDR);
  int tag, ctx; uint64_t bytes;
  memcpy(&tag, hdr, 4); memcpy(&ctx, hdr + 4, 4); memcpy(&bytes, hdr + 8, 8);
  ft_msg *m = (ft_msg *)malloc(sizeof(ft_msg) + bytes + 1);
  if (!m) ft_die("out of memory");
  m->tag = tag; m->ctx = ctx; m->bytes = bytes; m->next = NULL;
  ring_copy_out(chan_data(src, g_rank), head + FT_HDR, m->data, bytes);
  size_t total = FT_HDR + ((bytes + 15) & ~(size_t)15);
  atomic_store_explicit(&c->head, head + total, memory_order_release);
  if (g_unexp_tail[src]) g_unexp_tail[src]->next = m; else g_unexp_head[src] = m;
  g_unexp_tail[src] = m;
  return 1;
}
static void progress_all(void) { for (int s = 0; s < g_size; ++s) while (progress_from(s)) {} }

static void ft_send(const void *buf, size_t bytes, int dst, int tag, int ctx) {
  size_t total = FT_HDR + ((bytes + 15) & ~(size_t)15);
  if (total > g_ring) ft_die("message larger than ring (raise FTMPI_RING_MB)");
  ft_ring_ctl *c = chan_ctl(g_rank, dst);
  uint64_t tail = atomic_load_explicit(&c->tail, memory_order_relaxed);
  unsigned spin = 0;
  for (;;) {
    uint64_t head = atomic_load_explicit(&c->head, memory_order_acquire);
    if (tail - head + total <= g_ring) break;
    progress_all(); /* keep peers unblocked while we wait for room */
    if ((++spin & 63) == 0) { check_abort(); sched_yield(); }
  }
  char hdr[FT_HDR];
  memset(hdr, 0, FT_HDR);
  uint64_t b64 = bytes;
  memcpy(hdr, &tag, 4); memcpy(hdr + 4, &ctx, 4); memcpy(hdr + 8, &b64, 8);
  ring_copy_in(chan_data(g_rank, dst), tail, hdr, FT_HDR);
  ring_copy_in(chan_data(g_rank, dst), tail + FT_HDR, buf, bytes);
  atomic_store_explicit(&c->tail, tail + total, memory_order_release);
}

static void ft_recv(void *buf, size_t maxbytes, int src, int tag, int ctx, MPI_Status *st) {
  unsigned spin = 0;
  for (;;) {
    ft_msg *prev = NULL;
    for (ft_msg *m = g_unexp_head[src]; m; prev = m, m = m->next) {
      if (m->ctx != ctx) continue;
      if (!(tag == m->tag || (tag == MPI_ANY_TAG && m->tag >= 0))) continue;
      if (m->bytes > maxbytes) ft_die("message truncated");
      memcpy(buf, m->data, m->bytes);
      if (st) { st->MPI_SOURCE = src; st->MPI_TAG = m->tag; st->MPI_ERROR = 0; st->ftmpi_bytes = (int)m->bytes; }
      if (prev) prev->next = m->next; else g_unexp_head[src] = m->next;
      if (g_unexp_tail[src] == m) g_unexp_tail[src] = prev;
      free(m);
      return;
    }
    if (!progress_from(src)) {
      if ((++spin & 63) == 0) { check_abort(); sched_yield(); }
    }
  }
}

static ft_comm *C(MPI_Comm c) {
  if (c < 0 || c >= FT_MAXCOMM || !g_comm[c].used) ft_die("bad communicator");
  return &g_comm[c];
}
static int new_comm_slot(void) {
  for (int i = 1; i < FT_MAXCOMM; ++i) if (!g_comm[i].used) return i;
  ft_die("too many communicators");
  return -1;
}

/* ------------------------------------------------------------------ init */
int MPI_Init(int *argc, char ***argv) {
  (void)argc; (void)argv;
  if (g_inited) return MPI_SUCCESS;
  const char *er = getenv("FTMPI_RANK"), *es = getenv("FTMPI_SIZE"), *shm = getenv("FTMPI_SHM");
  const char *rm = getenv("FTMPI_RING_MB");
  g_ring = (size_t)(rm ? atoi(rm) : 64) << 20;
  if (er && es && shm) {
    g_rank = atoi(er); g_size = atoi(es);
    if (g_size > FT_MAXP) ft_die("too many ranks");
    int fd = open(shm, O_RDWR);
    if (fd < 0) ft_die("cannot open FTMPI_SHM");
    size_t total = 4096 + chan_stride() * (size_t)g_size * g_size;
    g_base = (char *)mmap(NULL, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (g_base == MAP_FAILED) ft_die("mmap failed");
  } else {
    g_rank = 0; g_size = 1;
    size_t total = 4096 + chan_stride();
    g_base = (char *)mmap(NULL, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (g_base == MAP_FAILED) ft_die("mmap failed");
  }
  g_glob = (ft_global *)g_base;
  g_comm[0].used = 1; g_comm[0].ctx = 0; g_comm[0].size = g_size; g_comm[0].rank = g_rank;
  g_comm[0].wr = (int *)malloc(sizeof(int) * g_size);
  for (int i = 0; i < g_size; ++i) g_comm[0].wr[i] = i;
  g_inited = 1;
  return MPI_SUCCESS;
}
int MPI_Initialized(int *flag) { *flag = g_inited; return MPI_SUCCESS; }
int MPI_Finalize(void) { if (g_inited) MPI_Barrier(MPI_COMM_WORLD); fflush(NULL); return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm c, int code) {
  (void)c;
  fprintf(stderr, "[ftmpi %d] MPI_Abort(%d)\n", g_rank, code);
  fflush(NULL);
  if (g_glob) { g_glob->abort_code = code ? code : 1; atomic_store(&g_glob->abort_flag, 1); }
  _exit(code ? code : 1);
}
int MPI_Comm_size(MPI_Comm c, int *s) { *s = C(c)->size; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm c, int *r) { *r = C(c)->rank; return MPI_SUCCESS; }
double MPI_Wtime(void) { struct timeval tv; gettimeofday(&tv, 0); return tv.tv_sec + 1e-6 * tv.tv_usec; }

/* ------------------------------------------------------------------- p2p */
int MPI_Send(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c) {
  ft_comm *cm = C(c);
  ft_send(b, (size_t)n * dsize(t), cm->wr[d], tag, cm->ctx);
  return MPI_SUCCESS;
}
int MPI_Recv(void *b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status *st) {
  ft_comm *cm = C(c);
  MPI_Status loc;
  ft_recv(b, (size_t)n * dsize(t), cm->wr[s], tag, cm->ctx, &loc);
  if (st) { *st = loc; st->MPI_SOURCE = s; }
  return MPI_SUCCESS;
}
static int new_req(void) {
  static int cursor = 0;
  for (int k = 0; k < FT_MAXREQ; ++k) {
    int i = (cursor + k) % FT_MAXREQ;
    if (!g_req[i].active) { cursor = i + 1; return i; }
  }
  ft_die("too many requests");
  return -1;
}
int MPI_Isend(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request *r) {
  MPI_Send(b, n, t, d, tag, c);
  int i = new_req();
  g_req[i].active = 2;
  memset(&g_req[i].st, 0, sizeof(MPI_Status));
  *r = i;
  return MPI_SUCCESS;
}
int MPI_Irecv(void *b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Request *r) {
  ft_comm *cm = C(c);
  int i = new_req();
  g_req[i].active = 1; g_req[i].buf = b; g_req[i].bytes = (size_t)n * dsize(t);
  g_req[i].src_world = cm->wr[s]; g_req[i].tag = tag; g_req[i].ctx = cm->ctx;
  g_req[i].st.MPI_SOURCE = s;
  *r = i;
  return MPI_SUCCESS;
}
int MPI_Wait(MPI_Request *r, MPI_Status *st) {
  if (*r == MPI_REQUEST_NULL) return MPI_SUCCESS;
  ft_req *q = &g_req[*r];
  if (q->active == 1) {
    int src_local = q->st.MPI_SOURCE;
    ft_recv(q->buf, q->bytes, q->src_world, q->tag, q->ctx, &q->st);
    q->st.MPI_SOURCE = src_local;
  }
  if (st) *st = q->st;
  q->active = 0;
  *r = MPI_REQUEST_NULL;
  return MPI_SUCCESS;
}
int MPI_Waitall(int n, MPI_Request *r, MPI_Status *st) {
  for (int i = 0; i < n; ++i) MPI_Wait(&r[i], st ? &st[i] : NULL);
  return MPI_SUCCESS;
}
int MPI_Get_count(const MPI_Status *s, MPI_Datatype t, int *c) { *c = s->ftmpi_bytes / dsize(t); return MPI_SUCCESS; }

/* ----------------------------------------------------------- collectives */
enum { T_BAR = -100, T_BCAST = -101, T_RED = -102, T_GATH = -103, T_A2A = -104, T_SCAN = -105, T_SCAT = -106 };

int MPI_Barrier(MPI_Comm c) {
  ft_comm *cm = C(c);
  char z = 0;
  if (cm->size == 1) return MPI_SUCCESS;
  if (cm->rank == 0) {
    for (int r = 1; r < cm->size; ++r) ft_recv(&z, 1, cm->wr[r], T_BAR, cm->ctx, NULL);
    for (int r = 1; r < cm->size; ++r) ft_send(&z, 1, cm->wr[r], T_BAR, cm->ctx);
  } else {
    ft_send(&z, 1, cm->wr[0], T_BAR, cm->ctx);
    ft_recv(&z, 1, cm->wr[0], T_BAR, cm->ctx, NULL);
  }
  return MPI_SUCCESS;
}
int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c) {
  ft_comm *cm = C(c);
  size_t bytes = (size_t)n * dsize(t);
  if (cm->rank == root) { for (int r = 0; r < cm->size; ++r) if (r != root) ft_send(b, bytes, cm->wr[r], T_BCAST, cm->ctx); }
  else ft_recv(b, bytes, cm->wr[root], T_BCAST, cm->ctx, NULL);
  return MPI_SUCCESS;
}

#define RED_LOOP(T) { T *a = (T *)acc; const T *x = (const T *)in; \
  for (int i = 0; i < n; ++i) { switch (op) { \
    case MPI_SUM: a[i] = a[i] + x[i]; break; \
    case MPI_MIN: if (x[i] < a[i]) a[i] = x[i]; break; \
    case MPI_MAX: if (x[i] > a[i]) a[i] = x[i]; break; \
    case MPI_LOR: a[i] = (a[i] || x[i]); break; \
    case MPI_LAND: a[i] = (a[i] && x[i]); break; \
    default: ft_die("unsupported reduction op"); } } }
typedef struct { double v; int i; } ft_double_int;
typedef struct { int v; int i; } ft_2int;
typedef struct { float v; int i; } ft_float_int;
#define LOC_LOOP(T) { T *a = (T *)acc; const T *x = (const T *)in; \
  for (int i = 0; i < n; ++i) { \
    if (op == MPI_MINLOC) { if (x[i].v < a[i].v || (x[i].v == a[i].v && x[i].i < a[i].i)) a[i] = x[i]; } \
    else if (op == MPI_MAXLOC) { if (x[i].v > a[i].v || (x[i].v == a[i].v && x[i].i < a[i].i)) a[i] = x[i]; } \
    else ft_die("unsupported op for pair type"); } }
static void combine(void *acc, const void *in, int n, MPI_Datatype t, MPI_Op op) {
  switch (t) {
    case MPI_INT: RED_LOOP(int) break;
    case MPI_UNSIGNED: RED_LOOP(unsigned) break;
    case MPI_FLOAT: RED_LOOP(float) break;
    case MPI_DOUBLE: RED_LOOP(double) break;
    case MPI_LONG: RED_LOOP(long) break;
    case MPI_LONG_LONG_INT: RED_LOOP(long long) break;
    case MPI_UNSIGNED_LONG: RED_LOOP(unsigned long) break;
    case MPI_CHAR: RED_LOOP(char) break;
    case MPI_SHORT: RED_LOOP(short) break;
    case MPI_DOUBLE_INT: LOC_LOOP(ft_double_int) break;
    case MPI_2INT: LOC_LOOP(ft_2int) break;
    case MPI_FLOAT_INT: LOC_LOOP(ft_float_int) break;
    default: ft_die("unsupported reduction type");
  }
}
int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) {
  ft_comm *cm = C(c);
  size_t bytes = (size_t)n * dsize(t);
  const void *mine = (s == MPI_IN_PLACE) ? r : s;
  if (cm->rank != root) { ft_send(mine, bytes, cm->wr[root], T_RED, cm->ctx); return MPI_SUCCESS; }
  char *acc = (char *)malloc(bytes + 1), *tmp = (char *)malloc(bytes + 1);
  for (int q = 0; q < cm->size; ++q) { /* ascending rank order */
    const void *in;
    if (q == root) in = mine; else { ft_recv(tmp, bytes, cm->wr[q], T_RED, cm->ctx, NULL); in = tmp; }
    if (q == 0) memcpy(acc, in, bytes); else combine(acc, in, n, t, op);
  }
  memcpy(r, acc, bytes);
  free(acc); free(tmp);
  return MPI_SUCCESS;
}
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  if (C(c)->rank == 0) MPI_Reduce(s, r, n, t, op, 0, c);
  else MPI_Reduce(s == MPI_IN_PLACE ? r : s, NULL, n, t, op, 0, c);
  return MPI_Bcast(r, n, t, 0, c);
}
int MPI_Scan(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  ft_comm *cm = C(c);
  size_t bytes = (size_t)n * dsize(t);
  if (s != MPI_IN_PLACE) memcpy(r, s, bytes);
  if (cm->rank > 0) {
    char *tmp = (char *)malloc(bytes + 1), *mine = (char *)malloc(bytes + 1);
    ft_recv(tmp, bytes, cm->wr[cm->rank - 1], T_SCAN, cm->ctx, NULL);
    memcpy(mine, r, bytes);
    memcpy(r, tmp, bytes);
    combine(r, mine, n, t, op); /* prefix (lower ranks) op mine */
    free(tmp); free(mine);
  }
  if (cm->rank + 1 < cm->size) ft_send(r, bytes, cm->wr[cm->rank + 1], T_SCAN, cm->ctx);
  return MPI_SUCCESS;
}
int MPI_Gatherv(const void *s, int sn, MPI_Datatype st, void *r, const int *rc, const int *rd, MPI_Datatype rt, int root, MPI_Comm c) {
  ft_comm *cm = C(c);
  if (cm->rank != root) { ft_send(s, (size_t)sn * dsize(st), cm->wr[root], T_GATH, cm->ctx); return MPI_SUCCESS; }
  for (int q = 0; q < cm->size; ++q) {
    char *dst = (char *)r + (size_t)rd[q] * dsize(rt);
    if (q == root) { if (s != MPI_IN_PLACE) memcpy(dst, s, (size_t)sn * dsize(st)); }
    else ft_recv(dst, (size_t)rc[q] * dsize(rt), cm->wr[q], T_GATH, cm->ctx, NULL);
  }
  return MPI_SUCCESS;
}
int MPI_Gather(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, int root, MPI_Comm c) {
  ft_comm *cm = C(c);
  int *rc = (int *)malloc(sizeof(int) * cm->size), *rd = (int *)malloc(sizeof(int) * cm->size);
  for (int q = 0; q < cm->size; ++q) { rc[q] = rn; rd[q] = q * rn; }
  MPI_Gatherv(s, sn, st, r, rc, rd, rt, root, c);
  free(rc); free(rd);
  return MPI_SUCCESS;
}
int MPI_Allgatherv(const void *s, int sn, MPI_Datatype st, void *r, const int *rc, const int *rd, MPI_Datatype rt, MPI_Comm c) {
  ft_comm *cm = C(c);
  if (s == MPI_IN_PLACE) {
    if (cm->rank == 0) MPI_Gatherv(MPI_IN_PLACE, 0, rt, r, rc, rd, rt, 0, c);
    else MPI_Gatherv((char *)r + (size_t)rd[cm->rank] * dsize(rt), rc[cm->rank], rt, r, rc, rd, rt, 0, c);
  } else {
    MPI_Gatherv(s, sn, st, r, rc, rd, rt, 0, c);
  }
  /* broadcast the assembled extent piecewise (displacements may be sparse) */
  for (int q = 0; q < cm->size; ++q)
    MPI_Bcast((char *)r + (size_t)rd[q] * dsize(rt), rc[q], rt, 0, c);
  return MPI_SUCCESS;
}
int MPI_Allgather(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, MPI_Comm c) {
  ft_comm *cm = C(c);
  int *rc = (int *)malloc(sizeof(int) * cm->size), *rd = (int *)malloc(sizeof(int) * cm->size);
  for (int q = 0; q < cm->size; ++q) { rc[q] = rn; rd[q] = q * rn; }
  MPI_Allgatherv(s, sn, st, r, rc, rd, rt, c);
  free(rc); free(rd);
  return MPI_SUCCESS;
}
int MPI_Alltoallv(const void *s, const int *sc, const int *sd, MPI_Datatype st, void *r, const int *rc, const int *rd, MPI_Datatype rt, MPI_Comm c) {
  ft_comm *cm = C(c);
  if (s == MPI_IN_PLACE) { /* send data is taken from the receive buffer */
    for (int q = 0; q < cm->size; ++q)
      ft_send((const char *)r + (size_t)rd[q] * dsize(rt), (size_t)rc[q] * dsize(rt), cm->wr[q], T_A2A, cm->ctx);
  } else {
    for (int q = 0; q < cm->size; ++q)
      ft_send((const char *)s + (size_t)sd[q] * dsize(st), (size_t)sc[q] * dsize(st), cm->wr[q], T_A2A, cm->ctx);
  }
  for (int q = 0; q < cm->size; ++q)
    ft_recv((char *)r + (size_t)rd[q] * dsize(rt), (size_t)rc[q] * dsize(rt), cm->wr[q], T_A2A, cm->ctx, NULL);
  return MPI_SUCCESS;
}
int MPI_Alltoall(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, MPI_Comm c) {
  ft_comm *cm = C(c);
  int *sc = (int *)malloc(sizeof(int) * cm->size * 4);
  int *sd = sc + cm->size, *rc = sd + cm->size, *rd = rc + cm->size;
  for (int q = 0; q < cm->size; ++q) { sc[q] = sn; sd[q] = q * sn; rc[q] = rn; rd[q] = q * rn; }
  MPI_Alltoallv(s, sc, sd, st, r, rc, rd, rt, c);
  free(sc);
  return MPI_SUCCESS;
}
int MPI_Scatterv(const void *s, const int *sc, const int *sd, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, int root, MPI_Comm c) {
  ft_comm *cm = C(c);
  if (cm->rank == root) {
    for (int q = 0; q < cm->size; ++q) {
      const char *src = (const char *)s + (size_t)sd[q] * dsize(st);
      if (q == root) { if (r != MPI_IN_PLACE) memcpy(r, src, (size_t)sc[q] * dsize(st)); }
      else ft_send(src, (size_t)sc[q] * dsize(st), cm->wr[q], T_SCAT, cm->ctx);
    }
  } else ft_recv(r, (size_t)rn * dsize(rt), cm->wr[root], T_SCAT, cm->ctx, NULL);
  return MPI_SUCCESS;
}

/* ---------------------------------------------------------- communicators */
static int agree_ctx(MPI_Comm parent) {
  int mine = g_next_ctx, mx = 0;
  MPI_Allreduce(&mine, &mx, 1, MPI_INT, MPI_MAX, parent);
  g_next_ctx = mx + 1;
  return mx;
}
int MPI_Comm_dup(MPI_Comm c, MPI_Comm *n) {
  ft_comm *cm = C(c);
  int ctx = agree_ctx(c);
  int id = new_comm_slot();
  g_comm[id].used = 1; g_comm[id].ctx = ctx; g_comm[id].size = cm->size; g_comm[id].rank = cm->rank;
  g_comm[id].wr = (int *)malloc(sizeof(int) * cm->size);
  memcpy(g_comm[id].wr, cm->wr, sizeof(int) * cm->size);
  *n = id;
  return MPI_SUCCESS;
}
int MPI_Comm_free(MPI_Comm *c) {
  if (*c > 0 && *c < FT_MAXCOMM && g_comm[*c].used) { free(g_comm[*c].wr); g_comm[*c].used = 0; }
  *c = MPI_COMM_NULL;
  return MPI_SUCCESS;
}
int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm *n) {
  ft_comm *cm = C(c);
  int ctx = agree_ctx(c);
  int P = cm->size;
  int *all = (int *)malloc(sizeof(int) * 2 * P);
  int mine[2] = {color, key};
  MPI_Allgather(mine, 2, MPI_INT, all, 2, MPI_INT, c);
  if (color == MPI_UNDEFINED) { free(all); *n = MPI_COMM_NULL; return MPI_SUCCESS; }
  int *members = (int *)malloc(sizeof(int) * P), cnt = 0;
  for (int q = 0; q < P; ++q) if (all[2 * q] == color) members[cnt++] = q;
  /* stable sort by key, ties by parent rank */
  for (int i = 1; i < cnt; ++i) {
    int m = members[i], j = i - 1;
    while (j >= 0 && all[2 * members[j] + 1] > all[2 * m + 1]) { members[j + 1] = members[j]; --j; }
    members[j + 1] = m;
  }
  int id = new_comm_slot();
  g_comm[id].used = 1; g_comm[id].ctx = ctx; g_comm[id].size = cnt;
  g_comm[id].wr = (int *)malloc(sizeof(int) * cnt);
  for (int i = 0; i < cnt; ++i) { g_comm[id].wr[i] = cm->wr[members[i]]; if (members[i] == cm->rank) g_comm[id].rank = i; }
  free(all); free(members);
  *n = id;
  return MPI_SUCCESS;
}

/* ------------------------------------------------------------- MPI-IO log */
struct ftmpi_file { int fd; MPI_Comm comm; };
int MPI_Info_create(MPI_Info *i) { *i = 0; return MPI_SUCCESS; }
int MPI_Info_set(MPI_Info i, const char *k, const char *v) { (void)i; (void)k; (void)v; return MPI_SUCCESS; }
int MPI_Info_free(MPI_Info *i) { *i = 0; return MPI_SUCCESS; }
int MPI_File_open(MPI_Comm c, const char *name, int mode, MPI_Info info, MPI_File *f) {
  (void)info;
  ft_comm *cm = C(c);
  int ok = 1;
  if (cm->rank == 0) {
    int flags = O_WRONLY | O_APPEND;
    if (mode & MPI_MODE_CREATE) flags |= O_CREAT;
    if (mode & MPI_MODE_EXCL) flags |= O_EXCL;
    int fd = open(name, flags, 0644);
    if (fd < 0) ok = 0; else close(fd);
  }
  MPI_Bcast(&ok, 1, MPI_INT, 0, c);
  if (!ok) { *f = NULL; return 1; }
  struct ftmpi_file *h = (struct ftmpi_file *)malloc(sizeof(*h));
  h->fd = open(name, O_WRONLY | O_APPEND);
  h->comm = c;
  *f = h;
  return h->fd >= 0 ? MPI_SUCCESS : 1;
}
int MPI_File_close(MPI_File *f) { if (*f) { close((*f)->fd); free(*f); *f = NULL; } return MPI_SUCCESS; }
int MPI_File_delete(const char *n, MPI_Info i) { (void)i; unlink(n); return MPI_SUCCESS; }
int MPI_File_write_shared(MPI_File f, const void *b, int n, MPI_Datatype t, MPI_Status *s) {
  (void)s;
  if (f && f->fd >= 0) { ssize_t w = write(f->fd, b, (size_t)n * dsize(t)); (void)w; }
  return MPI_SUCCESS;
}
int MPI_File_write_ordered(MPI_File f, const void *b, int n, MPI_Datatype t, MPI_Status *s) {
  (void)s;
  if (!f) return MPI_SUCCESS;
  ft_comm *cm = C(f->comm);
  for (int q = 0; q < cm->size; ++q) {
    if (q == cm->rank && f->fd >= 0) { ssize_t w = write(f->fd, b, (size_t)n * dsize(t)); (void)w; }
    MPI_Barrier(f->comm);
  }
  return MPI_SUCCESS;
}
