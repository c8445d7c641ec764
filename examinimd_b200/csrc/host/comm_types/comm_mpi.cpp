// CommMPI host class: sequencing of the six phases, buffer growth and the count handshakes of
// src/comm_types/comm_mpi.cpp:193-466.  All arithmetic on atoms is in kernels/comm_mpi.cu / comm.cu.
#include "comm_mpi.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static double g_t_exchange = 0.0, g_t_halo = 0.0, g_t_publish = 0.0;
static int g_n_rebuild = 0, g_n_calls = 0;
static const size_t kParticleBytes = 72; // sizeof(Particle), src/system.h:43-55

static int env_int(const char *a, const char *b, const char *c, int dflt) {
  const char *names[3] = {a, b, c};
  for (const char *n : names)
    if (n)
      if (const char *v = getenv(n)) return atoi(v);
  return dflt;
}

CommMPI::CommMPI(System *s, T_X_FLOAT comm_depth_) : Comm(s, comm_depth_), net(nullptr), peer(nullptr) {
  proc_rank = env_int("RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", 0);
  proc_size = env_int("WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", 1);
  if (proc_size < 1) proc_size = 1;
  memset(&dec, 0, sizeof dec);
  for (int p = 0; p < 6; p++) proc_num_send[p] = proc_num_recv[p] = num_ghost[p] = ghost_offsets[p] = 0;
  if (proc_size > 1 && emd_net_create(&net, system->ctx, proc_size, proc_rank, nullptr)) fail("emd_net_create");
  // EMD_HALO_TRANSPORT=nccl keeps the send/recv groups for the per-step refresh (measurement switch); default: peer stores
  const char *tr = getenv("EMD_HALO_TRANSPORT");
  if (net && !(tr && !strcmp(tr, "nccl")) && emd_peer_create(&peer, net, system->ctx, proc_size, proc_rank)) fail("emd_peer_create");
  system->do_print = system->do_print && proc_rank == 0; // src/system.cpp:61-67: only rank 0 prints
}

CommMPI::~CommMPI() {
  if (getenv("EMD_PEER_DEBUG") && g_n_rebuild)
    fprintf(stderr, "CommMPI[rank %d]: per re-neighboring (host wall, synchronised): exchange %.0f us, exchange_halo %.0f us, peer publish %.0f us (%d)\n", proc_rank,
            1e6 * g_t_exchange / g_n_rebuild, 1e6 * g_t_halo / g_n_rebuild, 1e6 * g_t_publish / g_n_rebuild, g_n_rebuild);
  if (peer) emd_peer_destroy(peer);
  if (net) emd_net_destroy(net);
}

void CommMPI::init() {}

void CommMPI::fail(const char *what) {
  fprintf(stderr, "CommMPI[rank %d]: %s: %s\n", proc_rank, what, emd_last_error());
  emd_host_exit(1);
}

void CommMPI::ensure_bytes(DeviceArray<char> &b, size_t bytes) {
  if (bytes <= b.extent()) return;
  if (!b.alloc(bytes + bytes / 2 + 1024)) fail("buffer allocation"); // half again: a lattice start under-estimates the faces of a melt by up to 15 %
}

// src/comm_types/comm_mpi.cpp:52-147
void CommMPI::create_domain_decomposition() {
  const double L[3] = {system->domain_x, system->domain_y, system->domain_z};
  if (emd_comm_decompose(proc_size, proc_rank, L, &dec)) fail("decompose");
  system->sub_domain_x = dec.sub[0]; system->sub_domain_y = dec.sub[1]; system->sub_domain_z = dec.sub[2];
  system->sub_domain_lo_x = dec.sub_lo[0]; system->sub_domain_lo_y = dec.sub_lo[1]; system->sub_domain_lo_z = dec.sub_lo[2];
  system->sub_domain_hi_x = dec.sub_hi[0]; system->sub_domain_hi_y = dec.sub_hi[1]; system->sub_domain_hi_z = dec.sub_hi[2];
}

// src/comm_types/comm_mpi.cpp:193-289
void CommMPI::exchange() {
  emd_ctx *ctx = system->ctx;
  static const bool dbg = getenv("EMD_PEER_DEBUG") != nullptr;
  double t0 = 0.0;
  if (dbg) { emd_ctx_sync(ctx); t0 = now_s(); }
  const double L[3] = {system->domain_x, system->domain_y, system->domain_z};
  T_INT N_local = system->N_local, N_ghost = 0;
  const int wrap[3] = {dec.grid[0] == 1, dec.grid[1] == 1, dec.grid[2] == 1};
  if (emd_comm_wrap_dims(ctx, system->x, N_local, L, wrap)) fail("wrap");
  T_INT N_total_recv = 0, N_total_send = 0;
  // The two phases of a dimension do not depend on each other (an atom that arrives from -d in the even phase lies inside
  // the brick in d and cannot be selected by the odd one; packed atoms are marked and skipped): both are packed first,
  // their counts travel as one handshake and their atoms as one message group -- half the synchronisations and groups
  // of the phase-by-phase sequence (:219-256), same atoms, same order.
  for (int dim = 0; dim < 3; dim++) {
    const int pa = 2 * dim;
    if (!decomposed(pa)) continue;
    DeviceArray<char> *pk[2] = {&pack_buffer, &pack_buffer2}, *up[2] = {&unpack_buffer, &unpack_buffer2};
    int send[2] = {0, 0}, recv[2] = {0, 0};
    // a floor under the migration buffers: nothing leaves at step 0 of a lattice start, and the first real migration should
    // not pay for allocations and a second packing pass inside somebody's measurement window
    for (int k = 0; k < 2; k++) {
      const size_t floor_bytes = (size_t)std::max<T_INT>(4096, N_local / 128) * kParticleBytes;
      ensure_bytes(*pk[k], floor_bytes);
      ensure_bytes(*up[k], floor_bytes);
    }
    for (int k = 0; k < 2; k++) {
      for (int attempt = 0; attempt < 2; attempt++) {
        const int cap = (int)(pk[k]->extent() / kParticleBytes);
        int count = 0;
        if (emd_comm_exchange_pack(ctx, pa + k, &dec, L, system->x, system->v, system->q, system->id, system->type, N_local + N_ghost,
                                   pk[k]->ptr, cap, &count))
          fail("exchange_pack");
        send[k] = count;
        if (count <= cap) break;
        ensure_bytes(*pk[k], (size_t)(count * 1.1 + 16) * kParticleBytes); // :230-238
      }
    }
    const int peer_send[2] = {dec.neighbor_send[pa], dec.neighbor_send[pa + 1]}, peer_recv[2] = {dec.neighbor_recv[pa], dec.neighbor_recv[pa + 1]};
    if (emd_net_exchange_counts2(net, send, peer_send, peer_recv, recv)) fail("count handshake");
    for (int k = 0; k < 2; k++) ensure_bytes(*up[k], (size_t)recv[k] * kParticleBytes);
    const T_INT need = N_local + N_ghost + recv[0] + recv[1];
    if (need > system->N_max) system->grow(need + (recv[0] + recv[1]) / 4 + 16);
    if (emd_net_group_begin(net)) fail("group");
    for (int k = 0; k < 2; k++)
      if (emd_net_sendrecv(net, pk[k]->ptr, (size_t)send[k] * kParticleBytes, peer_send[k], up[k]->ptr, (size_t)recv[k] * kParticleBytes, peer_recv[k]))
        fail("sendrecv");
    if (emd_net_group_end(net)) fail("group");
    for (int k = 0; k < 2; k++) {
      if (emd_comm_unpack(ctx, up[k]->ptr, recv[k], N_local + N_ghost, system->x, system->v, system->q, system->id, system->type)) fail("unpack");
      N_ghost += recv[k];
      N_total_recv += recv[k];
      N_total_send += send[k];
    }
  }
  const T_INT N_local_start = N_local, N_exchange = N_ghost;
  N_local = N_local + N_total_recv - N_total_send; // :259
  if (emd_comm_exchange_compact(ctx, system->x, system->v, system->q, system->id, system->type, N_local, N_local_start + N_exchange))
    fail("compact");
  system->N_local = N_local;
  system->N_ghost = 0;
  if (dbg) { emd_ctx_sync(ctx); if (g_n_calls++ >= 5) g_t_exchange += now_s() - t0; if (getenv("EMD_COMM_TRACE") && proc_rank == 0) fprintf(stderr, "TRACE exchange %.3f ms\n", 1e3 * (now_s() - t0)); }
}

// src/comm_types/comm_mpi.cpp:291-380
void CommMPI::exchange_halo() {
  emd_ctx *ctx = system->ctx;
  static const bool dbg = getenv("EMD_PEER_DEBUG") != nullptr;
  double t0 = 0.0, t1 = 0.0;
  if (dbg) { emd_ctx_sync(ctx); t0 = now_s(); }
  const double L[3] = {system->domain_x, system->domain_y, system->domain_z};
  const T_INT N_local = system->N_local;
  T_INT N_ghost = 0;
  for (int phase = 0; phase < 6; phase++) {
    // an odd phase does not re-scan the ghosts its even twin just received (:306,:346)
    const T_INT nparticles = N_local + N_ghost - ((phase % 2 == 1) ? proc_num_recv[phase - 1] : 0);
    int count = 0;
    if (decomposed(phase) && phase % 2 == 1) {
      count = proc_num_recv[phase]; // shipped together with its even twin below
    } else if (decomposed(phase)) {
      // both phases of the dimension scan the same atoms [0, nparticles): packed first, one count handshake, one message group
      DeviceArray<char> *pk[2] = {&pack_buffer, &pack_buffer2}, *up[2] = {&unpack_buffer, &unpack_buffer2};
      int send[2] = {0, 0}, recv[2] = {0, 0};
      for (int k = 0; k < 2; k++) {
        const int ph = phase + k;
        for (int attempt = 0; attempt < 2; attempt++) {
          const int cap = (int)std::min(pk[k]->extent() / kParticleBytes, pack_indicies[ph].extent());
          if (emd_comm_halo_pack(ctx, ph, &dec, L, comm_depth, system->x, system->v, system->q, system->id, system->type, nparticles,
                                 pack_indicies[ph].ptr, pk[k]->ptr, cap, &send[k]))
            fail("halo_pack");
          if (send[k] <= cap) break;
          ensure_bytes(*pk[k], (size_t)(send[k] * 1.5 + 1024) * kParticleBytes); // :319-327
          if ((size_t)send[k] > pack_indicies[ph].extent() && !pack_indicies[ph].alloc((size_t)(send[k] * 1.5) + 1024)) fail("alloc pack_indicies");
        }
        proc_num_send[ph] = send[k];
      }
      const int peer_send[2] = {dec.neighbor_send[phase], dec.neighbor_send[phase + 1]}, peer_recv[2] = {dec.neighbor_recv[phase], dec.neighbor_recv[phase + 1]};
      if (emd_net_exchange_counts2(net, send, peer_send, peer_recv, recv)) fail("count handshake");
      proc_num_recv[phase] = recv[0]; proc_num_recv[phase + 1] = recv[1];
      for (int k = 0; k < 2; k++) ensure_bytes(*up[k], (size_t)recv[k] * kParticleBytes);
      const T_INT need = N_local + N_ghost + recv[0] + recv[1];
      if (need > system->N_max) system->grow(need + (recv[0] + recv[1]) / 4 + 16);
      if (emd_net_group_begin(net)) fail("group");
      for (int k = 0; k < 2; k++)
        if (emd_net_sendrecv(net, pk[k]->ptr, (size_t)send[k] * kParticleBytes, peer_send[k], up[k]->ptr, (size_t)recv[k] * kParticleBytes, peer_recv[k]))
          fail("sendrecv");
      if (emd_net_group_end(net)) fail("group");
      if (emd_comm_unpack(ctx, up[0]->ptr, recv[0], N_local + N_ghost, system->x, system->v, system->q, system->id, system->type) ||
          emd_comm_unpack(ctx, up[1]->ptr, recv[1], N_local + N_ghost + recv[0], system->x, system->v, system->q, system->id, system->type))
        fail("unpack");
      count = recv[0];
    } else { // the in-process twin of the phase, same kernels as CommSerial (:344-371)
      const double lo[3] = {system->sub_domain_lo_x, system->sub_domain_lo_y, system->sub_domain_lo_z};
      const double hi[3] = {system->sub_domain_hi_x, system->sub_domain_hi_y, system->sub_domain_hi_z};
      for (int attempt = 0; attempt < 2; attempt++) {
        if (emd_comm_halo_phase(ctx, phase, system->x, system->v, system->q, system->id, system->type, nparticles, N_local + N_ghost,
                                system->N_max, pack_indicies[phase].ptr, (int)pack_indicies[phase].extent(), L, lo, hi, comm_depth, &count))
          fail("halo_phase");
        bool redo = false;
        if (N_local + N_ghost + count > system->N_max) { system->grow(N_local + N_ghost + count + count / 4); redo = true; }
        if ((size_t)count > pack_indicies[phase].extent()) {
          if (!pack_indicies[phase].alloc((size_t)(count * 1.5) + 1024)) fail("alloc pack_indicies");
          redo = true;
        }
        if (!redo) break;
      }
      proc_num_send[phase] = proc_num_recv[phase] = count;
    }
    num_ghost[phase] = count;
    N_ghost += count;
  }
  system->N_ghost = N_ghost;
  // the per-step refresh ships 24 B per atom in either direction: make sure both buffers hold the largest phase
  size_t most = 0;
  for (int d = 0; d < 3; d++)
    most = std::max(most, (size_t)std::max(proc_num_send[2 * d] + proc_num_send[2 * d + 1], proc_num_recv[2 * d] + proc_num_recv[2 * d + 1]));
  ensure_bytes(pack_buffer, most * kParticleBytes);
  ensure_bytes(unpack_buffer, most * kParticleBytes);
  // leading dimensions without decomposition: one resolved refresh kernel instead of two phase kernels per dimension
  local_dims = 0;
  while (local_dims < 3 && !decomposed(2 * local_dims)) local_dims++;
  local_ghosts = 0;
  for (int phase = 0; phase < 2 * local_dims; phase++) local_ghosts += proc_num_recv[phase];
  if (local_dims > 0 && local_ghosts > 0) {
    if (local_root.extent() < (size_t)local_ghosts) {
      if (!local_root.alloc((size_t)local_ghosts + local_ghosts / 2 + 16) || !local_shift.alloc(3 * ((size_t)local_ghosts + local_ghosts / 2 + 16))) fail("alloc local roots");
    }
    const int *lists[6];
    int counts[6];
    for (int p = 0; p < 6; p++) { lists[p] = pack_indicies[p].ptr; counts[p] = p < 2 * local_dims ? proc_num_send[p] : 0; }
    if (emd_comm_halo_resolve(ctx, lists, counts, N_local, L, local_root.ptr, local_shift.ptr)) fail("halo_resolve");
  }
  if (dbg) { emd_ctx_sync(ctx); t1 = now_s(); if (g_n_calls > 5) { g_t_halo += t1 - t0; g_n_rebuild++; } }
  if (peer) { // the neighbours learn where my ghost rows of each phase start (and my position arrays, if they were reallocated)
    int ghost_begin[6];
    T_INT g = 0;
    for (int phase = 0; phase < 6; phase++) { ghost_begin[phase] = system->N_local + g; g += proc_num_recv[phase]; }
    if (!system->x_alt) { emd_peer_destroy(peer); peer = nullptr; }
    else if (emd_peer_publish(peer, &dec, system->x, system->x_alt, ghost_begin)) fail("peer publish");
  }
  if (dbg) { emd_ctx_sync(ctx); if (g_n_calls > 5) g_t_publish += now_s() - t1; if (getenv("EMD_COMM_TRACE") && proc_rank == 0) fprintf(stderr, "TRACE exchange_halo %.3f ms publish %.3f ms\n", 1e3 * (t1 - t0), 1e3 * (now_s() - t1)); }
}

// src/comm_types/comm_mpi.cpp:382-423: no host synchronisation anywhere in here.  The two phases of a dimension do not
// depend on each other (phase 2d+1 never replays ghosts received in phase 2d, :306), so a decomposed dimension packs
// both directions, ships them as ONE NCCL group and unpacks both: three exchanges per step instead of six.
void CommMPI::update_halo() { refresh(false); }

// the peer-store refresh without the wait for the neighbours' stores (the force kernel named by Force::gates_halo waits)
bool CommMPI::update_halo_deferred() {
  if (!(peer && emd_peer_ready(peer))) return false;
  refresh(true);
  return true;
}

void CommMPI::refresh(bool defer) {
  emd_ctx *ctx = system->ctx;
  const double L[3] = {system->domain_x, system->domain_y, system->domain_z};
  T_INT ghost_begin[6];
  T_INT N_ghost = 0;
  for (int phase = 0; phase < 6; phase++) { ghost_begin[phase] = system->N_local + N_ghost; N_ghost += proc_num_recv[phase]; }
  const bool by_peer = peer && emd_peer_ready(peer);
  if (by_peer && emd_peer_begin_update(peer, &dec, system->x)) fail("peer begin_update");
  if (local_dims > 0 && local_ghosts > 0 && emd_comm_halo_refresh(ctx, system->x, system->N_local, local_ghosts, local_root.ptr, local_shift.ptr))
    fail("halo_refresh");
  for (int dim = local_dims; dim < 3; dim++) {
    const int pa = 2 * dim, pb = 2 * dim + 1;
    if (decomposed(pa) && by_peer) {
      // the pack kernel stores into the neighbours' ghost rows; the stream then waits for their stores into mine
      if (emd_peer_update_dim(peer, &dec, L, dim, system->x, pack_indicies[pa].ptr, proc_num_send[pa], pack_indicies[pb].ptr, proc_num_send[pb], defer ? 1 : 0))
        fail("peer update_dim");
    } else if (decomposed(pa)) {
      double *send_a = (double *)pack_buffer.ptr, *send_b = send_a + 3 * (size_t)proc_num_send[pa];
      // a refresh message is exactly the ghost rows of x (3 doubles per ghost, in ghost order): receive in place, no unpack
      double *recv_a = system->x + 3 * (size_t)ghost_begin[pa], *recv_b = system->x + 3 * (size_t)ghost_begin[pb];
      if (emd_comm_halo_update_pack(ctx, pa, &dec, L, system->x, pack_indicies[pa].ptr, proc_num_send[pa], send_a) ||
          emd_comm_halo_update_pack(ctx, pb, &dec, L, system->x, pack_indicies[pb].ptr, proc_num_send[pb], send_b))
        fail("halo_update_pack");
      if (emd_net_group_begin(net) ||
          emd_net_sendrecv(net, send_a, (size_t)proc_num_send[pa] * 24, dec.neighbor_send[pa], recv_a, (size_t)proc_num_recv[pa] * 24, dec.neighbor_recv[pa]) ||
          emd_net_sendrecv(net, send_b, (size_t)proc_num_send[pb] * 24, dec.neighbor_send[pb], recv_b, (size_t)proc_num_recv[pb] * 24, dec.neighbor_recv[pb]) ||
          emd_net_group_end(net))
        fail("sendrecv");
    } else {
      if (by_peer && emd_peer_wait_all(peer)) fail("peer wait"); // this dimension forwards the ghosts of the earlier ones
      for (int phase = pa; phase <= pb; phase++)
        if (emd_comm_halo_update_phase(ctx, phase, system->x, system->v, system->q, system->id, system->type, pack_indicies[phase].ptr,
                                       proc_num_send[phase], ghost_begin[phase], L))
          fail("halo_update_phase");
    }
  }
}

// src/comm_types/comm_mpi.cpp:425-466: reverse direction, phases 5..0; the ghost rows of f are contiguous and go out in place
void CommMPI::update_force() {
  emd_ctx *ctx = system->ctx;
  ghost_offsets[0] = system->N_local;
  for (int phase = 1; phase < 6; phase++) ghost_offsets[phase] = ghost_offsets[phase - 1] + proc_num_recv[phase - 1];
  for (int phase = 5; phase >= 0; phase--) {
    if (decomposed(phase)) {
      if (emd_net_sendrecv(net, system->f + 3 * (size_t)ghost_offsets[phase], (size_t)proc_num_recv[phase] * 24, dec.neighbor_recv[phase],
                           pack_buffer.ptr, (size_t)proc_num_send[phase] * 24, dec.neighbor_send[phase]))
        fail("sendrecv");
      if (emd_comm_force_unpack(ctx, system->f, pack_indicies[phase].ptr, proc_num_send[phase], (const double *)pack_buffer.ptr))
        fail("force_unpack");
    } else {
      if (emd_comm_force_fold_phase(ctx, system->f, pack_indicies[phase].ptr, proc_num_send[phase], ghost_offsets[phase]))
        fail("force_fold_phase");
    }
  }
}

// src/comm_types/comm_mpi.cpp:150-191.  NB the reference's reduce_min_* call MPI_MAX (:183,:189); they are unused by the
// hot path and implemented here as what their names say would need a MIN transport op, so they keep the reference's behaviour.
void CommMPI::reduce_float(T_FLOAT *v, T_INT N) { if (net && emd_net_allreduce(net, v, N, 1, 0)) fail("allreduce"); }
void CommMPI::reduce_int(T_INT *v, T_INT N) { if (net && emd_net_allreduce(net, v, N, 0, 0)) fail("allreduce"); }
void CommMPI::reduce_max_float(T_FLOAT *v, T_INT N) { if (net && emd_net_allreduce(net, v, N, 1, 1)) fail("allreduce"); }
void CommMPI::reduce_max_int(T_INT *v, T_INT N) { if (net && emd_net_allreduce(net, v, N, 0, 1)) fail("allreduce"); }
void CommMPI::reduce_min_float(T_FLOAT *v, T_INT N) { reduce_max_float(v, N); }
void CommMPI::reduce_min_int(T_INT *v, T_INT N) { reduce_max_int(v, N); }
void CommMPI::scan_int(T_INT *v, T_INT N) {
  if (!net) return;
  for (T_INT k = 0; k < N; k++)
    if (emd_net_scan_int(net, &v[k])) fail("scan");
}
void CommMPI::weighted_reduce_float(T_FLOAT *, T_INT *, T_INT) {} // declared by the reference's Comm, defined nowhere, never called

int CommMPI::process_rank() { return proc_rank; }
int CommMPI::num_processes() { return proc_size; }

void CommMPI::error(const char *errormsg) { // :472-476 (MPI_Abort): a non-zero exit makes the launcher tear the job down
  if (proc_rank == 0) printf("%s\n", errormsg);
  fflush(stdout);
  emd_host_exit(1);
}

const char *CommMPI::name() { return "CommMPI"; }
