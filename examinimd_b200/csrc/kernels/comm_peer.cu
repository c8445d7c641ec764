// comm_peer.cu -- CommMPI::update_halo (src/comm_types/comm_mpi.cpp:382-423) without a transport library on the critical
// path: the pack kernel of a phase stores the shifted positions STRAIGHT INTO THE NEIGHBOUR'S GHOST ROWS over NVLink
// (CUDA IPC mappings of the neighbour's position arrays, exchanged once per re-neighboring), then raises a sequence-number
// flag in the neighbour's memory; the receiver's stream waits for the flags of a dimension before it goes on (the next
// dimension forwards ghosts of this one, as in the reference's dimension-ordered protocol).  A reverse flag ("consumed":
// I am done reading the ghosts of the previous refresh, and this is the position array that is current on my side) keeps a
// fast rank from overwriting rows that a slow neighbour still reads.  Per decomposed dimension: one push kernel + one
// single-thread wait kernel on the module stream -- ~10 us instead of a ~100 us NCCL send/recv group (profiles/, round 2).
#include "common.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace emd;

namespace {

struct PeerFlags {     // one block per rank, written by its neighbours with system-scope stores
  int arrived[6];      // [phase p] by the sender of the phase-p message I receive: sequence number of the last refresh that has landed
  int consumed[6];     // [phase p] by the receiver of MY phase-p message: the refresh before which it stopped reading its ghosts
  int curbuf[6];       // [phase p] by the same: which of its two registered position arrays is current over there
  int pad[14];
};
static_assert(sizeof(PeerFlags) == 128, "PeerFlags is one 128-byte line");

struct Record {        // what a rank publishes after a re-neighboring
  cudaIpcMemHandle_t flags, x[2];
  int ghost_begin[6];
  int pad[2];
};

__device__ __forceinline__ int ld_acquire_sys(const int *p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void st_release_sys(int *p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

struct SignalArgs { int *consumed[6]; int *curbuf[6]; int valid[6]; };
struct PushPhase {
  const int *idx;        // my pack list of the phase (source rows, ghosts of earlier dimensions included)
  int count;
  double shift;          // box length, applied by the rank on the global boundary only (comm_mpi.h:164,177)
  double *dst[2];        // the receiver's two position arrays
  int dst_begin;         // its first ghost row of this phase
  const int *my_consumed, *my_curbuf; // in my flag block, written by the receiver
  int *peer_arrived;     // in the receiver's flag block
};

struct PreWait { const int *flag[4]; int n; }; // arrived flags of the earlier decomposed dimensions (this one forwards their ghosts)

__global__ void __launch_bounds__(256) peer_push_kernel(PushPhase a, PushPhase b, int dim, const double *__restrict__ x, int q, unsigned *done, unsigned long long *dbg,
                                                        PreWait pre, SignalArgs sig, int sig_cur) {
  __shared__ int s_cur[2];
  const unsigned long long t0 = gtime();
  if (threadIdx.x < 2) {
    const PushPhase &ph = threadIdx.x == 0 ? a : b;
    while (ld_acquire_sys(ph.my_consumed) < q) __nanosleep(200);
    s_cur[threadIdx.x] = ld_acquire_sys(ph.my_curbuf) & 1;
  } else if (blockIdx.x == 0 && threadIdx.x >= 64 && threadIdx.x < 70 && sig.valid[threadIdx.x - 64]) {
    // (first push of a refresh) tell the senders of my ghosts that the previous ones are no longer read, and which array is current
    st_release_sys(sig.curbuf[threadIdx.x - 64], sig_cur);
    st_release_sys(sig.consumed[threadIdx.x - 64], q);
  } else if (threadIdx.x >= 32 && threadIdx.x < 32 + pre.n) {
    while (ld_acquire_sys(pre.flag[threadIdx.x - 32]) < q) __nanosleep(200);
    __threadfence_system();
  }
  __syncthreads();
  const unsigned long long t1 = gtime();
  const long long total = (long long)a.count + b.count;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const bool second = e >= a.count;
    const PushPhase &ph = second ? b : a;
    const long long ii = second ? e - a.count : e;
    const size_t i = (size_t)ph.idx[ii];
    double p[3] = {x[3 * i], x[3 * i + 1], x[3 * i + 2]};
    p[dim] += ph.shift;
    double *d = ph.dst[s_cur[second ? 1 : 0]] + 3 * ((size_t)ph.dst_begin + (size_t)ii);
    d[0] = p[0]; d[1] = p[1]; d[2] = p[2];
  }
  const unsigned long long t2 = gtime();
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (dbg && blockIdx.x == 0) { atomicAdd(dbg + 0, t1 - t0); atomicAdd(dbg + 1, t2 - t1); atomicAdd(dbg + 2, gtime() - t2); atomicAdd(dbg + 3, 1ull); }
    if (atomicAdd(done, 1u) == gridDim.x - 1) { // the last block: every store of this launch is visible system-wide
      *done = 0u;
      __threadfence_system();
      st_release_sys(a.peer_arrived, q);
      st_release_sys(b.peer_arrived, q);
    }
  }
}

} // namespace

struct emd_peer {
  emd_net *net = nullptr;
  emd_ctx *ctx = nullptr;
  int nranks = 1, rank = 0;
  PeerFlags *d_flags = nullptr;
  unsigned *d_done = nullptr;          // [3] block counters of the push kernels
  unsigned long long *d_dbg = nullptr; // EMD_PEER_DEBUG=1: accumulated nanoseconds of the kernel stages
  Record mine;
  std::vector<Record> all;             // last published records, by rank
  std::vector<PeerFlags *> flags_of;   // mapped flag blocks, by rank (nullptr: not a neighbour / not mapped yet)
  std::vector<double *> x_of[2];       // mapped position arrays, by rank
  std::vector<Record> mapped;          // the records the mappings above were opened from
  double *reg_x[2] = {nullptr, nullptr};
  int seq = 0;
  bool ready = false;
  int pending_mask = 0; // phases of the refresh in progress whose arrival nobody has waited for yet
  SignalArgs sig;       // the "consumed" message of the refresh in progress, sent by its first push kernel
  int sig_cur = 0;
  bool sig_pending = false;
};

extern "C" {

int emd_peer_create(emd_peer **out, emd_net *net, emd_ctx *ctx, int nranks, int rank) {
  if (!out || !net || !ctx) { set_error("emd_peer_create: bad arguments"); return 1; }
  emd_peer *p = new emd_peer();
  p->net = net; p->ctx = ctx; p->nranks = nranks; p->rank = rank;
  EMD_CUDA(cudaMalloc((void **)&p->d_flags, sizeof(PeerFlags)));
  EMD_CUDA(cudaMemset(p->d_flags, 0, sizeof(PeerFlags)));
  EMD_CUDA(cudaMalloc((void **)&p->d_done, 4 * sizeof(unsigned)));
  EMD_CUDA(cudaMemset(p->d_done, 0, 4 * sizeof(unsigned)));
  if (getenv("EMD_PEER_DEBUG") && atoi(getenv("EMD_PEER_DEBUG"))) { EMD_CUDA(cudaMalloc((void **)&p->d_dbg, 64)); EMD_CUDA(cudaMemset(p->d_dbg, 0, 64)); }
  memset(&p->mine, 0, sizeof p->mine);
  EMD_CUDA(cudaIpcGetMemHandle(&p->mine.flags, p->d_flags));
  p->all.resize(nranks); p->mapped.resize(nranks);
  for (auto &r : p->mapped) memset(&r, 0, sizeof r);
  p->flags_of.assign(nranks, nullptr);
  p->x_of[0].assign(nranks, nullptr); p->x_of[1].assign(nranks, nullptr);
  *out = p;
  return 0;
}

void emd_peer_destroy(emd_peer *p) {
  if (!p) return;
  if (p->d_dbg) {
    unsigned long long h[8];
    cudaMemcpy(h, p->d_dbg, 64, cudaMemcpyDeviceToHost);
    fprintf(stderr, "emd_peer[rank %d]: push: wait-consumed %.1f us, copy %.1f us, fence+flag %.1f us (%llu launches)\n", p->rank,
            h[3] ? h[0] / 1e3 / h[3] : 0.0, h[3] ? h[1] / 1e3 / h[3] : 0.0, h[3] ? h[2] / 1e3 / h[3] : 0.0, h[3]);
    cudaFree(p->d_dbg);
  }
  for (int r = 0; r < p->nranks; r++) {
    if (p->flags_of[r]) cudaIpcCloseMemHandle(p->flags_of[r]);
    for (int k = 0; k < 2; k++) if (p->x_of[k][r]) cudaIpcCloseMemHandle(p->x_of[k][r]);
  }
  if (p->d_flags) cudaFree(p->d_flags);
  if (p->d_done) cudaFree(p->d_done);
  delete p;
}

// Collective, after CommMPI::exchange_halo: every rank publishes its two position arrays (whole cudaMalloc allocations) and the
// first ghost row of each phase; the mappings of the neighbours are (re)opened when an array was reallocated.
int emd_peer_publish(emd_peer *p, const emd_decomp *dec, double *d_x0, double *d_x1, const int ghost_begin[6]) {
  if (!p || !dec) { set_error("emd_peer_publish: bad arguments"); return 1; }
  p->ready = false;
  EMD_CUDA(cudaStreamSynchronize(p->ctx->stream));
  if (d_x0 != p->reg_x[0]) { EMD_CUDA(cudaIpcGetMemHandle(&p->mine.x[0], d_x0)); p->reg_x[0] = d_x0; }
  if (d_x1 != p->reg_x[1]) { EMD_CUDA(cudaIpcGetMemHandle(&p->mine.x[1], d_x1)); p->reg_x[1] = d_x1; }
  for (int k = 0; k < 6; k++) p->mine.ghost_begin[k] = ghost_begin[k];
  if (emd_net_allgather_bytes(p->net, &p->mine, (int)sizeof(Record), p->all.data())) return 1;
  for (int ph = 0; ph < 6; ph++) {
    if (dec->grid[ph / 2] <= 1) continue;
    const int peers[2] = {dec->neighbor_send[ph], dec->neighbor_recv[ph]};
    for (int r : peers) {
      if (r < 0 || r >= p->nranks || r == p->rank) { set_error("emd_peer_publish: bad neighbour %d", r); return 1; }
      const Record &rec = p->all[r];
      Record &old = p->mapped[r];
      if (!p->flags_of[r]) {
        EMD_CUDA(cudaIpcOpenMemHandle((void **)&p->flags_of[r], rec.flags, cudaIpcMemLazyEnablePeerAccess));
        old.flags = rec.flags;
      }
      // the two arrays trade places at every cell sort (System::swap_sorted): look the handles up among the open mappings
      // before opening anything (cudaIpcOpenMemHandle costs milliseconds)
      double *have[2] = {p->x_of[0][r], p->x_of[1][r]}, *want[2] = {nullptr, nullptr};
      bool used[2] = {false, false};
      for (int k = 0; k < 2; k++)
        for (int j = 0; j < 2; j++)
          if (!want[k] && have[j] && !used[j] && memcmp(&old.x[j], &rec.x[k], sizeof(cudaIpcMemHandle_t)) == 0) { want[k] = have[j]; used[j] = true; }
      for (int j = 0; j < 2; j++)
        if (have[j] && !used[j]) EMD_CUDA(cudaIpcCloseMemHandle(have[j]));
      for (int k = 0; k < 2; k++) {
        if (!want[k]) EMD_CUDA(cudaIpcOpenMemHandle((void **)&want[k], rec.x[k], cudaIpcMemLazyEnablePeerAccess));
        p->x_of[k][r] = want[k];
        old.x[k] = rec.x[k];
      }
    }
  }
  p->ready = true;
  return 0;
}

int emd_peer_ready(const emd_peer *p) { return p && p->ready; }

// start of a refresh: tell the senders of my ghosts that the previous ones are no longer read, and which array is current
int emd_peer_begin_update(emd_peer *p, const emd_decomp *dec, const double *d_x_current) {
  if (!p || !p->ready) { set_error("emd_peer_begin_update: not published"); return 1; }
  const int cur = d_x_current == p->reg_x[0] ? 0 : d_x_current == p->reg_x[1] ? 1 : -1;
  if (cur < 0) { set_error("emd_peer_begin_update: the current position array is not one of the two published ones"); return 1; }
  p->seq++;
  p->pending_mask = 0;
  SignalArgs &a = p->sig;
  memset(&a, 0, sizeof a);
  for (int ph = 0; ph < 6; ph++) {
    if (dec->grid[ph / 2] <= 1) continue;
    PeerFlags *sender = p->flags_of[dec->neighbor_recv[ph]]; // the rank whose phase-ph message I receive
    a.consumed[ph] = &sender->consumed[ph]; a.curbuf[ph] = &sender->curbuf[ph]; a.valid[ph] = 1;
  }
  p->sig_cur = cur;
  p->sig_pending = true;
  return 0;
}

// both phases of one decomposed dimension: push my border atoms into the neighbours' ghost rows, wait for theirs
int emd_peer_wait_all(emd_peer *p) {
  if (!p) { set_error("emd_peer_wait_all: no peer transport"); return 1; }
  if (!p->pending_mask) return 0;
  if (emd_ctx_set_halo_gate(p->ctx, p->d_flags->arrived, p->seq, p->pending_mask) || emd_ctx_halo_gate_wait(p->ctx)) return 1;
  p->pending_mask = 0;
  return 0;
}

int emd_peer_update_dim(emd_peer *p, const emd_decomp *dec, const double domain[3], int dim, const double *d_x, const int *d_pack_idx_a,
                        int count_a, const int *d_pack_idx_b, int count_b, int defer_wait) {
  if (!p || !p->ready || dim < 0 || dim > 2 || dec->grid[dim] <= 1) { set_error("emd_peer_update_dim: bad arguments"); return 1; }
  PushPhase ph[2];
  const int *idx[2] = {d_pack_idx_a, d_pack_idx_b};
  const int cnt[2] = {count_a, count_b};
  for (int k = 0; k < 2; k++) {
    const int phase = 2 * dim + k, upper = (k == 0);
    const int S = dec->neighbor_send[phase];
    ph[k].idx = idx[k]; ph[k].count = cnt[k];
    ph[k].shift = 0.0;
    if (upper && dec->pos[dim] == dec->grid[dim] - 1) ph[k].shift = -domain[dim];
    if (!upper && dec->pos[dim] == 0) ph[k].shift = domain[dim];
    ph[k].dst[0] = p->x_of[0][S]; ph[k].dst[1] = p->x_of[1][S];
    ph[k].dst_begin = p->all[S].ghost_begin[phase];
    ph[k].my_consumed = &p->d_flags->consumed[phase]; ph[k].my_curbuf = &p->d_flags->curbuf[phase];
    ph[k].peer_arrived = &p->flags_of[S]->arrived[phase];
  }
  const long long total = (long long)count_a + count_b;
  const int grid = (int)std::max(1LL, std::min(64LL, (total + 255) / 256));
  PreWait pre;
  memset(&pre, 0, sizeof pre);
  for (int e = 0; e < 6 && pre.n < 4; e++)
    if ((p->pending_mask >> e) & 1) pre.flag[pre.n++] = &p->d_flags->arrived[e];
  SignalArgs sig;
  memset(&sig, 0, sizeof sig);
  if (p->sig_pending) { sig = p->sig; p->sig_pending = false; }
  EMD_LAUNCH(p->ctx, peer_push_kernel, grid, 256, 0, ph[0], ph[1], dim, d_x, p->seq, p->d_done + dim, p->d_dbg, pre, sig, p->sig_cur);
  p->pending_mask |= 3 << (2 * dim);
  if (defer_wait) return emd_ctx_set_halo_gate(p->ctx, p->d_flags->arrived, p->seq, p->pending_mask); // the consumer waits
  return emd_peer_wait_all(p);
}

} // extern "C"
