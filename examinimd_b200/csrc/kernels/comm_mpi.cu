// comm_mpi.cu -- multi-GPU communication: pack/unpack kernels and the NCCL transport.
// Replaces CommMPI (src/comm_types/comm_mpi.cpp:52-466, functors src/comm_types/comm_mpi.h:134-497):
// 3-D brick decomposition, one process per GPU, six dimension-ordered phases per operation.
//
//   exchange      leavers (x > hi / x < lo, strict) are packed as 72-byte Particles and marked type = -1,
//                 shipped, appended; the holes are then filled from the tail (:193-289)
//   exchange_halo border atoms within comm_depth of a face are packed (shifted by the box length only
//                 on the rank at the global boundary), shipped, appended as ghosts; the source
//                 indices are kept for the per-step refresh (:291-380)
//   update_halo   positions only, 24 B per ghost, replaying the saved indices (:382-423)
//   update_force  reverse direction, phases 5..0, f[src] += (:425-466)
//
// Differences that stay behind the API: every compaction is a STABLE flag/scan/scatter instead of an
// atomic counter (ghost and migration order = ascending source index, so a run is reproducible);
// messages are NCCL send/recv pairs grouped per phase ON THE MODULE STREAM over NVLink (the
// reference hands device pointers to blocking MPI_Send/MPI_Wait); the per-step refresh and the
// force fold never touch the host: counts are known from the last ghost build, so pack -> send/recv
// -> unpack is pure stream order.  Only the ghost build / migration read a count back per phase,
// as the reference does (:229,:318).
// NCCL is loaded with dlopen the first time a transport is created: single-GPU runs never need it.
#include "common.cuh"
#include <arpa/inet.h>
#include <dlfcn.h>
#include <netdb.h>
#include <nccl.h>
#include <sys/socket.h>
#include <unistd.h>
#include <cstdlib>

using namespace emd;

namespace {

struct Particle { // src/system.h:43-55, the wire format: 72 bytes = 18 ints
  double x, y, z, vx, vy, vz, mass, q;
  int id, type;
};
static_assert(sizeof(Particle) == 72, "Particle must stay 72 bytes");

struct Atoms { double *x, *v, *q; int *id, *type; };

__device__ __forceinline__ Particle get_particle(const Atoms &a, size_t i) { // src/system.h:103-111
  Particle p;
  p.x = a.x[3 * i]; p.y = a.x[3 * i + 1]; p.z = a.x[3 * i + 2];
  p.vx = a.v[3 * i]; p.vy = a.v[3 * i + 1]; p.vz = a.v[3 * i + 2];
  p.mass = 0.0; p.q = a.q[i]; p.id = a.id[i]; p.type = a.type[i];
  return p;
}
__device__ __forceinline__ void set_particle(const Atoms &a, size_t i, const Particle &p) { // src/system.h:114-120
  a.x[3 * i] = p.x; a.x[3 * i + 1] = p.y; a.x[3 * i + 2] = p.z;
  a.v[3 * i] = p.vx; a.v[3 * i + 1] = p.vy; a.v[3 * i + 2] = p.vz;
  a.q[i] = p.q; a.id[i] = p.id; a.type[i] = p.type;
}

// TagExchangeSelf (comm_mpi.h:134-153): periodic wrap only in the dimensions that are NOT decomposed
__global__ void __launch_bounds__(256) wrap_dims_kernel(double *__restrict__ x, long long n3, double Lx, double Ly, double Lz, int wx,
                                                        int wy, int wz) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n3) return;
  const int d = (int)(e % 3);
  if (!(d == 0 ? wx : (d == 1 ? wy : wz))) return;
  const double L = d == 0 ? Lx : (d == 1 ? Ly : Lz);
  const double xo = x[e];
  double xn = xo;
  if (xo > L) xn -= L;
  if (xo < 0) xn += L;
  if (xn != xo) x[e] = xn;
}

// flags for one phase: MODE 0 = migration (TagExchangePack: type >= 0 and strictly outside the face),
// MODE 1 = halo (TagHaloPack: within comm_depth of the face, inclusive)
template <int MODE>
__global__ void __launch_bounds__(256) flag_kernel(const double *__restrict__ x, const int *__restrict__ type, int n_scan, int dim,
                                                   int upper, double thr, int *__restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_scan) return;
  const double xi = x[3 * (size_t)i + dim];
  if (MODE == 0) flags[i] = (type[i] >= 0) && (upper ? (xi > thr) : (xi < thr));
  else flags[i] = upper ? (xi >= thr) : (xi <= thr);
}

template <int MODE>
__global__ void __launch_bounds__(256) pack_kernel(Atoms a, int n_scan, const int *__restrict__ flags, const int *__restrict__ offsets,
                                                   int dim, double shift, int *__restrict__ pack_idx, Particle *__restrict__ buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_scan || !flags[i]) return;
  const int slot = offsets[i];
  Particle p = get_particle(a, (size_t)i);
  if (dim == 0) p.x += shift; else if (dim == 1) p.y += shift; else p.z += shift;
  buf[slot] = p;
  if (pack_idx) pack_idx[slot] = i;
  if (MODE == 0) a.type[i] = -1; // the atom has left (comm_mpi.h:163)
}

__global__ void __launch_bounds__(256) unpack_kernel(Atoms a, int dst_begin, int count, const Particle *__restrict__ buf) { // TagUnpack
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  set_particle(a, (size_t)dst_begin + i, buf[i]);
}

// hole filling after a migration (TagExchangeCreateDestList / TagExchangeCompact, comm_mpi.h:229-250): the c-th hole
// below n_new (ascending) takes the c-th live atom found walking DOWN from the end
__global__ void __launch_bounds__(256) hole_flag_kernel(const int *__restrict__ type, int n_new, int n_end, int *__restrict__ hole,
                                                        int *__restrict__ live) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_new) hole[i] = type[i] < 0;
  const int ntail = n_end - n_new;
  if (i < ntail) live[i] = type[n_end - 1 - i] >= 0;
}
__global__ void __launch_bounds__(256) hole_list_kernel(const int *__restrict__ hole, const int *__restrict__ hole_off, int n_new,
                                                        int *__restrict__ dest_list) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_new && hole[i]) dest_list[hole_off[i]] = i;
}
__global__ void __launch_bounds__(256) hole_fill_kernel(Atoms a, const int *__restrict__ live, const int *__restrict__ live_off, int n_new,
                                                        int n_end, const int *__restrict__ dest_list) {
  const int ii = blockIdx.x * blockDim.x + threadIdx.x;
  if (ii >= n_end - n_new || !live[ii]) return;
  const size_t src = (size_t)n_end - 1 - ii, dst = (size_t)dest_list[live_off[ii]];
  set_particle(a, dst, get_particle(a, src));
}

__global__ void __launch_bounds__(256) update_pack_kernel(const double *__restrict__ x, const int *__restrict__ pack_idx, int count, int dim,
                                                          double shift, double *__restrict__ buf) { // TagHaloUpdatePack
  const int ii = blockIdx.x * blockDim.x + threadIdx.x;
  if (ii >= count) return;
  const size_t i = (size_t)pack_idx[ii];
  double p[3] = {x[3 * i], x[3 * i + 1], x[3 * i + 2]};
  p[dim] += shift;
  buf[3 * (size_t)ii] = p[0]; buf[3 * (size_t)ii + 1] = p[1]; buf[3 * (size_t)ii + 2] = p[2];
}
__global__ void __launch_bounds__(256) update_unpack_kernel(double *__restrict__ x, int ghost_begin, int count,
                                                            const double *__restrict__ buf) { // TagHaloUpdateUnpack
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 3LL * count) return;
  x[3 * (size_t)ghost_begin + e] = buf[e];
}
__global__ void __launch_bounds__(256) force_unpack_kernel(double *__restrict__ f, const int *__restrict__ pack_idx, int count,
                                                           const double *__restrict__ buf) { // TagHaloForceUnpack
  const int ii = blockIdx.x * blockDim.x + threadIdx.x;
  if (ii >= count) return;
  const size_t i = (size_t)pack_idx[ii]; // a source index appears once per phase
  f[3 * i] += buf[3 * (size_t)ii]; f[3 * i + 1] += buf[3 * (size_t)ii + 1]; f[3 * i + 2] += buf[3 * (size_t)ii + 2];
}

int phase_geometry(int phase, const emd_decomp *d, const double domain[3], double depth, bool halo, int *dim, int *upper, double *thr,
                   double *shift) {
  if (phase < 0 || phase > 5 || !d) { set_error("comm: bad phase %d", phase); return 1; }
  *dim = phase / 2; *upper = (phase % 2 == 0);
  if (halo) *thr = *upper ? d->sub_hi[*dim] - depth : d->sub_lo[*dim] + depth; // comm_mpi.h:309,322,...
  else *thr = *upper ? d->sub_hi[*dim] : d->sub_lo[*dim];                      // comm_mpi.h:159,172,...
  // the box length is applied only by the rank that sits on the global boundary (comm_mpi.h:164,177,...)
  *shift = 0.0;
  if (*upper && d->pos[*dim] == d->grid[*dim] - 1) *shift = -domain[*dim];
  if (!*upper && d->pos[*dim] == 0) *shift = domain[*dim];
  return 0;
}

// ------------------------------------------------------------------------------ NCCL via dlopen
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
} g_nccl;

int load_nccl() {
  if (g_nccl.lib) return 0;
  const char *names[] = {getenv("EMD_NCCL_LIB"), "libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
  void *h = nullptr;
  for (const char *n : names) {
    if (!n || !*n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) { set_error("comm: cannot load NCCL (libnccl.so.2): %s", dlerror()); return 1; }
#define EMD_SYM(field, name)                                                                     \
  do {                                                                                           \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                                  \
    if (!g_nccl.field) { set_error("comm: NCCL symbol %s missing", name); dlclose(h); return 1; } \
  } while (0)
  EMD_SYM(GetUniqueId, "ncclGetUniqueId"); EMD_SYM(CommInitRank, "ncclCommInitRank"); EMD_SYM(CommDestroy, "ncclCommDestroy");
  EMD_SYM(Send, "ncclSend"); EMD_SYM(Recv, "ncclRecv"); EMD_SYM(AllReduce, "ncclAllReduce"); EMD_SYM(AllGather, "ncclAllGather");
  EMD_SYM(GroupStart, "ncclGroupStart"); EMD_SYM(GroupEnd, "ncclGroupEnd"); EMD_SYM(GetErrorString, "ncclGetErrorString");
#undef EMD_SYM
  g_nccl.lib = h;
  return 0;
}

#define EMD_NCCL(call)                                                                                      \
  do {                                                                                                      \
    ncclResult_t r_ = (call);                                                                               \
    if (r_ != ncclSuccess) { set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); return 1; } \
  } while (0)

// TCP hand-over of the 128-byte NCCL id from rank 0 (MASTER_ADDR : MASTER_PORT + 29), for hosts that
// have no channel of their own (the standalone ExaMiniMD binary under torchrun / any RANK/WORLD_SIZE launcher)
int tcp_share_id(int rank, int nranks, ncclUniqueId *id) {
  const char *addr = getenv("MASTER_ADDR");
  const char *port_s = getenv("EMD_RENDEZVOUS_PORT");
  int port = port_s ? atoi(port_s) : (getenv("MASTER_PORT") ? atoi(getenv("MASTER_PORT")) + 29 : 29529);
  if (!addr || !*addr) addr = "127.0.0.1";
  if (rank == 0) {
    int ls = socket(AF_INET, SOCK_STREAM, 0);
    if (ls < 0) { set_error("comm: socket() failed"); return 1; }
    int one = 1;
    setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
    sockaddr_in sa;
    memset(&sa, 0, sizeof sa);
    sa.sin_family = AF_INET; sa.sin_addr.s_addr = htonl(INADDR_ANY); sa.sin_port = htons((unsigned short)port);
    if (bind(ls, (sockaddr *)&sa, sizeof sa) != 0 || listen(ls, nranks) != 0) { close(ls); set_error("comm: cannot listen on port %d", port); return 1; }
    for (int k = 1; k < nranks; k++) {
      int c = accept(ls, nullptr, nullptr);
      if (c < 0) { close(ls); set_error("comm: accept() failed"); return 1; }
      size_t off = 0;
      while (off < sizeof *id) { ssize_t w = write(c, (const char *)id + off, sizeof *id - off); if (w <= 0) break; off += (size_t)w; }
      close(c);
    }
    close(ls);
    return 0;
  }
  addrinfo hints, *res = nullptr;
  memset(&hints, 0, sizeof hints);
  hints.ai_family = AF_INET; hints.ai_socktype = SOCK_STREAM;
  char ps[16];
  snprintf(ps, sizeof ps, "%d", port);
  if (getaddrinfo(addr, ps, &hints, &res) != 0 || !res) { set_error("comm: cannot resolve %s", addr); return 1; }
  for (int attempt = 0; attempt < 1200; attempt++) { // up to ~120 s
    int c = socket(AF_INET, SOCK_STREAM, 0);
    if (c >= 0 && connect(c, res->ai_addr, res->ai_addrlen) == 0) {
      size_t off = 0;
      while (off < sizeof *id) { ssize_t r = read(c, (char *)id + off, sizeof *id - off); if (r <= 0) break; off += (size_t)r; }
      close(c);
      freeaddrinfo(res);
      if (off == sizeof *id) return 0;
      set_error("comm: short read of the NCCL id");
      return 1;
    }
    if (c >= 0) close(c);
    usleep(100000);
  }
  freeaddrinfo(res);
  set_error("comm: rank %d could not reach rank 0 at %s:%d", rank, addr, port);
  return 1;
}

} // namespace

struct emd_net {
  emd_ctx *ctx = nullptr;
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  int *d_small = nullptr; // 64 ints / 32 doubles of device staging for counts and scalar reductions
  bool in_group = false;  // between emd_net_group_begin / _end: message pairs join one NCCL group
  char *d_gather = nullptr; size_t gather_cap = 0; // emd_net_allgather_bytes staging
};

extern "C" {

// CommMPI::create_domain_decomposition (comm_mpi.cpp:52-147): host arithmetic only, no GPU needed
int emd_comm_decompose(int nranks, int rank, const double domain[3], emd_decomp *out) {
  if (!out || nranks < 1 || rank < 0 || rank >= nranks) { set_error("emd_comm_decompose: bad arguments"); return 1; }
  const double area_xy = domain[0] * domain[1], area_xz = domain[0] * domain[2], area_yz = domain[1] * domain[2];
  double smallest_surface = 2.0 * (area_xy + area_xz + area_yz);
  int g[3] = {1, 1, 1};
  for (int ipx = 1; ipx <= nranks; ipx++) {
    if (nranks % ipx) continue;
    const int nremain = nranks / ipx;
    for (int ipy = 1; ipy <= nremain; ipy++) {
      if (nremain % ipy) continue;
      const int ipz = nremain / ipy;
      const double surface = area_xy / ipx / ipy + area_xz / ipx / ipz + area_yz / ipy / ipz;
      if (surface < smallest_surface) { smallest_surface = surface; g[0] = ipx; g[1] = ipy; g[2] = ipz; } // strict: first found wins
    }
  }
  out->nranks = nranks; out->rank = rank;
  for (int d = 0; d < 3; d++) out->grid[d] = g[d];
  out->pos[2] = rank / (g[0] * g[1]);
  out->pos[1] = (rank % (g[0] * g[1])) / g[0];
  out->pos[0] = rank % g[0];
  const int stride[3] = {1, g[0], g[0] * g[1]};
  for (int d = 0; d < 3; d++) {
    if (g[d] > 1) {
      out->neighbor_send[2 * d + 1] = (out->pos[d] > 0) ? rank - stride[d] : rank + stride[d] * (g[d] - 1);
      out->neighbor_send[2 * d] = (out->pos[d] < g[d] - 1) ? rank + stride[d] : rank - stride[d] * (g[d] - 1);
    } else {
      out->neighbor_send[2 * d] = out->neighbor_send[2 * d + 1] = -1;
    }
    out->neighbor_recv[2 * d] = out->neighbor_send[2 * d + 1];
    out->neighbor_recv[2 * d + 1] = out->neighbor_send[2 * d];
    out->sub[d] = domain[d] / g[d];
    out->sub_lo[d] = out->pos[d] * out->sub[d];
    out->sub_hi[d] = (out->pos[d] + 1) * out->sub[d];
  }
  return 0;
}

int emd_comm_wrap_dims(emd_ctx *ctx, double *d_x, int n_local, const double domain[3], const int wrap_dim[3]) {
  if (n_local <= 0 || !(wrap_dim[0] || wrap_dim[1] || wrap_dim[2])) return 0;
  const long long n3 = 3LL * n_local;
  EMD_LAUNCH(ctx, wrap_dims_kernel, grid_for(n3, 256), 256, 0, d_x, n3, domain[0], domain[1], domain[2], wrap_dim[0], wrap_dim[1], wrap_dim[2]);
  return 0;
}

static int pack_common(emd_ctx *ctx, bool halo, int phase, const emd_decomp *dec, const double domain[3], double depth, double *d_x,
                       double *d_v, double *d_q, int *d_id, int *d_type, int n_scan, int *d_pack_idx, void *d_buf, int capacity,
                       int *h_count) {
  int dim, upper;
  double thr, shift;
  if (phase_geometry(phase, dec, domain, depth, halo, &dim, &upper, &thr, &shift)) return 1;
  int count = 0;
  if (n_scan > 0) {
    // (sized for emd_comm_exchange_compact as well, which first runs when the first atom migrates: no allocation then)
    if (ctx->s_a.ensure(sizeof(int) * (4 * (size_t)n_scan + 4096))) return 1;
    int *flags = ctx->s_a.as<int>(), *offsets = flags + n_scan, *d_total = offsets + n_scan;
    if (halo) EMD_LAUNCH(ctx, (flag_kernel<1>), grid_for(n_scan, 256), 256, 0, d_x, d_type, n_scan, dim, upper, thr, flags);
    else EMD_LAUNCH(ctx, (flag_kernel<0>), grid_for(n_scan, 256), 256, 0, d_x, d_type, n_scan, dim, upper, thr, flags);
    if (exclusive_scan_int(ctx, flags, offsets, n_scan, d_total)) return 1;
    EMD_CUDA(cudaMemcpyAsync(ctx->h_pinned, d_total, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    EMD_CUDA(cudaStreamSynchronize(ctx->stream)); // comm_mpi.cpp:229 / :318
    count = ctx->h_pinned[0];
    // nothing is written unless everything fits: the caller grows its buffers and calls again (:230-238, :319-327)
    if (count > 0 && count <= capacity) {
      Atoms a = {d_x, d_v, d_q, d_id, d_type};
      if (halo) EMD_LAUNCH(ctx, (pack_kernel<1>), grid_for(n_scan, 256), 256, 0, a, n_scan, flags, offsets, dim, shift, d_pack_idx, (Particle *)d_buf);
      else EMD_LAUNCH(ctx, (pack_kernel<0>), grid_for(n_scan, 256), 256, 0, a, n_scan, flags, offsets, dim, shift, d_pack_idx, (Particle *)d_buf);
    }
  }
  if (h_count) *h_count = count;
  return 0;
}

int emd_comm_exchange_pack(emd_ctx *ctx, int phase, const emd_decomp *dec, const double domain[3], double *d_x, double *d_v, double *d_q,
                           int *d_id, int *d_type, int n_scan, void *d_pack_buffer, int capacity, int *h_count) {
  return pack_common(ctx, false, phase, dec, domain, 0.0, d_x, d_v, d_q, d_id, d_type, n_scan, nullptr, d_pack_buffer, capacity, h_count);
}

int emd_comm_halo_pack(emd_ctx *ctx, int phase, const emd_decomp *dec, const double domain[3], double comm_depth, double *d_x, double *d_v,
                       double *d_q, int *d_id, int *d_type, int n_scan, int *d_pack_indicies, void *d_pack_buffer, int capacity,
                       int *h_count) {
  return pack_common(ctx, true, phase, dec, domain, comm_depth, d_x, d_v, d_q, d_id, d_type, n_scan, d_pack_indicies, d_pack_buffer, capacity,
                     h_count);
}

int emd_comm_unpack(emd_ctx *ctx, const void *d_unpack_buffer, int count, int dst_begin, double *d_x, double *d_v, double *d_q, int *d_id,
                    int *d_type) {
  if (count <= 0) return 0;
  Atoms a = {d_x, d_v, d_q, d_id, d_type};
  EMD_LAUNCH(ctx, unpack_kernel, grid_for(count, 256), 256, 0, a, dst_begin, count, (const Particle *)d_unpack_buffer);
  return 0;
}

int emd_comm_exchange_compact(emd_ctx *ctx, double *d_x, double *d_v, double *d_q, int *d_id, int *d_type, int n_new, int n_end) {
  const int ntail = n_end - n_new;
  if (ntail <= 0 || n_new <= 0) return 0;
  if (ctx->s_a.ensure(sizeof(int) * (3 * (size_t)n_new + 2 * (size_t)ntail + 8))) return 1;
  int *hole = ctx->s_a.as<int>(), *hole_off = hole + n_new, *dest = hole_off + n_new, *live = dest + n_new, *live_off = live + ntail;
  const int nmax = n_new > ntail ? n_new : ntail;
  EMD_LAUNCH(ctx, hole_flag_kernel, grid_for(nmax, 256), 256, 0, d_type, n_new, n_end, hole, live);
  if (exclusive_scan_int(ctx, hole, hole_off, n_new, nullptr)) return 1;
  if (exclusive_scan_int(ctx, live, live_off, ntail, nullptr)) return 1;
  EMD_LAUNCH(ctx, hole_list_kernel, grid_for(n_new, 256), 256, 0, hole, hole_off, n_new, dest);
  Atoms a = {d_x, d_v, d_q, d_id, d_type};
  EMD_LAUNCH(ctx, hole_fill_kernel, grid_for(ntail, 256), 256, 0, a, live, live_off, n_new, n_end, dest);
  return 0;
}

int emd_comm_halo_update_pack(emd_ctx *ctx, int phase, const emd_decomp *dec, const double domain[3], const double *d_x,
                              const int *d_pack_indicies, int count, double *d_buffer) {
  int dim, upper;
  double thr, shift;
  if (phase_geometry(phase, dec, domain, 0.0, true, &dim, &upper, &thr, &shift)) return 1;
  if (count <= 0) return 0;
  EMD_LAUNCH(ctx, update_pack_kernel, grid_for(count, 256), 256, 0, d_x, d_pack_indicies, count, dim, shift, d_buffer);
  return 0;
}

int emd_comm_halo_update_unpack(emd_ctx *ctx, double *d_x, int ghost_begin, int count, const double *d_buffer) {
  if (count <= 0) return 0;
  EMD_LAUNCH(ctx, update_unpack_kernel, grid_for(3LL * count, 256), 256, 0, d_x, ghost_begin, count, d_buffer);
  return 0;
}

int emd_comm_force_unpack(emd_ctx *ctx, double *d_f, const int *d_pack_indicies, int count, const double *d_buffer) {
  if (count <= 0) return 0;
  EMD_LAUNCH(ctx, force_unpack_kernel, grid_for(count, 256), 256, 0, d_f, d_pack_indicies, count, d_buffer);
  return 0;
}

// ------------------------------------------------------------------------------ transport
int emd_net_unique_id(void *out128) {
  if (load_nccl()) return 1;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  EMD_NCCL(g_nccl.GetUniqueId((ncclUniqueId *)out128));
  return 0;
}

int emd_net_create(emd_net **out, emd_ctx *ctx, int nranks, int rank, const void *unique_id128) {
  if (!out || !ctx || nranks < 1 || rank < 0 || rank >= nranks) { set_error("emd_net_create: bad arguments"); return 1; }
  if (load_nccl()) return 1;
  ncclUniqueId id;
  if (unique_id128) memcpy(&id, unique_id128, sizeof id);
  else {
    if (rank == 0) EMD_NCCL(g_nccl.GetUniqueId(&id));
    if (nranks > 1 && tcp_share_id(rank, nranks, &id)) return 1;
  }
  emd_net *n = new emd_net();
  n->ctx = ctx; n->nranks = nranks; n->rank = rank;
  EMD_CUDA(cudaSetDevice(ctx->device));
  EMD_NCCL(g_nccl.CommInitRank(&n->comm, nranks, id, rank));
  EMD_CUDA(cudaMalloc((void **)&n->d_small, 256 + 8 * (size_t)nranks));
  *out = n;
  return 0;
}

void emd_net_destroy(emd_net *n) {
  if (!n) return;
  if (n->comm) g_nccl.CommDestroy(n->comm);
  if (n->d_small) cudaFree(n->d_small);
  if (n->d_gather) cudaFree(n->d_gather);
  delete n;
}

// one phase's message pair, stream-ordered on the context stream (replaces MPI_Irecv/MPI_Send/MPI_Wait,
// comm_mpi.cpp:245-251, 333-337, 401-404, 449-452); either side may be empty
int emd_net_sendrecv(emd_net *n, const void *d_send, unsigned long long send_bytes, int peer_send, void *d_recv,
                     unsigned long long recv_bytes, int peer_recv) {
  if (!n) { set_error("emd_net_sendrecv: no transport"); return 1; }
  if (send_bytes == 0 && recv_bytes == 0) return 0;
  if (!n->in_group) EMD_NCCL(g_nccl.GroupStart());
  if (recv_bytes) EMD_NCCL(g_nccl.Recv(d_recv, recv_bytes, ncclChar, peer_recv, n->comm, n->ctx->stream));
  if (send_bytes) EMD_NCCL(g_nccl.Send(d_send, send_bytes, ncclChar, peer_send, n->comm, n->ctx->stream));
  if (!n->in_group) { EMD_NCCL(g_nccl.GroupEnd()); n->ctx->launches++; }
  return 0;
}

// Several message pairs as ONE NCCL group = one fused send/recv kernel: the +d and -d phases of a dimension are
// independent of each other (an odd phase never scans the ghosts of its even twin, comm_mpi.cpp:306), so the per-step
// refresh needs three exchanges, not six.
int emd_net_group_begin(emd_net *n) {
  if (!n || n->in_group) { set_error("emd_net_group_begin: no transport or nested group"); return 1; }
  EMD_NCCL(g_nccl.GroupStart());
  n->in_group = true;
  return 0;
}
int emd_net_group_end(emd_net *n) {
  if (!n || !n->in_group) { set_error("emd_net_group_end: no open group"); return 1; }
  n->in_group = false;
  EMD_NCCL(g_nccl.GroupEnd());
  n->ctx->launches++;
  return 0;
}

// the count handshake of a phase (tag 100001 messages, comm_mpi.cpp:235-238 / :325-328); synchronises
int emd_net_exchange_count(emd_net *n, int send_count, int peer_send, int peer_recv, int *h_recv_count) {
  if (!n) { set_error("emd_net_exchange_count: no transport"); return 1; }
  emd_ctx *c = n->ctx;
  c->h_pinned[8] = send_count;
  EMD_CUDA(cudaMemcpyAsync(n->d_small, c->h_pinned + 8, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  EMD_NCCL(g_nccl.GroupStart());
  EMD_NCCL(g_nccl.Recv(n->d_small + 1, 1, ncclInt, peer_recv, n->comm, c->stream));
  EMD_NCCL(g_nccl.Send(n->d_small, 1, ncclInt, peer_send, n->comm, c->stream));
  EMD_NCCL(g_nccl.GroupEnd());
  EMD_CUDA(cudaMemcpyAsync(c->h_pinned + 9, n->d_small + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  EMD_CUDA(cudaStreamSynchronize(c->stream));
  *h_recv_count = c->h_pinned[9];
  return 0;
}

// the count handshakes of the two phases of one dimension as ONE message group and ONE synchronisation
int emd_net_exchange_counts2(emd_net *n, const int send_count[2], const int peer_send[2], const int peer_recv[2], int h_recv_count[2]) {
  if (!n) { set_error("emd_net_exchange_counts2: no transport"); return 1; }
  emd_ctx *c = n->ctx;
  c->h_pinned[8] = send_count[0]; c->h_pinned[9] = send_count[1];
  EMD_CUDA(cudaMemcpyAsync(n->d_small, c->h_pinned + 8, 2 * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  EMD_NCCL(g_nccl.GroupStart());
  for (int k = 0; k < 2; k++) {
    EMD_NCCL(g_nccl.Recv(n->d_small + 2 + k, 1, ncclInt, peer_recv[k], n->comm, c->stream));
    EMD_NCCL(g_nccl.Send(n->d_small + k, 1, ncclInt, peer_send[k], n->comm, c->stream));
  }
  EMD_NCCL(g_nccl.GroupEnd());
  EMD_CUDA(cudaMemcpyAsync(c->h_pinned + 10, n->d_small + 2, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  EMD_CUDA(cudaStreamSynchronize(c->stream));
  h_recv_count[0] = c->h_pinned[10]; h_recv_count[1] = c->h_pinned[11];
  return 0;
}

// MPI_Allreduce(IN_PLACE) on HOST values (comm_mpi.cpp:156-191): is_double 0 = int, 1 = double; op 0 = sum, 1 = max
int emd_net_allreduce(emd_net *n, void *h_values, int count, int is_double, int op) {
  if (!n) { set_error("emd_net_allreduce: no transport"); return 1; }
  if (count <= 0) return 0;
  emd_ctx *c = n->ctx;
  const size_t esz = is_double ? sizeof(double) : sizeof(int);
  if (esz * (size_t)count > 128) { set_error("emd_net_allreduce: at most %d values per call", (int)(128 / esz)); return 1; }
  void *stage = (void *)(c->h_pinned + 16);
  memcpy(stage, h_values, esz * (size_t)count);
  void *dbuf = (void *)(n->d_small + 8);
  EMD_CUDA(cudaMemcpyAsync(dbuf, stage, esz * (size_t)count, cudaMemcpyHostToDevice, c->stream));
  EMD_NCCL(g_nccl.AllReduce(dbuf, dbuf, (size_t)count, is_double ? ncclDouble : ncclInt, op == 1 ? ncclMax : ncclSum, n->comm, c->stream));
  EMD_CUDA(cudaMemcpyAsync(stage, dbuf, esz * (size_t)count, cudaMemcpyDeviceToHost, c->stream));
  EMD_CUDA(cudaStreamSynchronize(c->stream));
  memcpy(h_values, stage, esz * (size_t)count);
  return 0;
}

// MPI_Scan(IN_PLACE, SUM) of one int per rank (comm_mpi.cpp:150-154): all-gather + inclusive prefix on the host
int emd_net_scan_int(emd_net *n, int *h_value) {
  if (!n) { set_error("emd_net_scan_int: no transport"); return 1; }
  emd_ctx *c = n->ctx;
  if (n->nranks > 32) { set_error("emd_net_scan_int: more than 32 ranks"); return 1; }
  int *stage = c->h_pinned + 16;
  stage[0] = *h_value;
  int *dsend = n->d_small + 8, *dall = n->d_small + 64;
  EMD_CUDA(cudaMemcpyAsync(dsend, stage, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  EMD_NCCL(g_nccl.AllGather(dsend, dall, 1, ncclInt, n->comm, c->stream));
  EMD_CUDA(cudaMemcpyAsync(stage, dall, sizeof(int) * (size_t)n->nranks, cudaMemcpyDeviceToHost, c->stream));
  EMD_CUDA(cudaStreamSynchronize(c->stream));
  int s = 0;
  for (int r = 0; r <= n->rank; r++) s += stage[r];
  *h_value = s;
  return 0;
}

// MPI_Allgather of a small record per rank (host memory in, host memory out); synchronises
int emd_net_allgather_bytes(emd_net *n, const void *h_in, int nbytes, void *h_out_all) {
  if (!n || nbytes <= 0 || (nbytes & 3)) { set_error("emd_net_allgather_bytes: bad arguments"); return 1; }
  emd_ctx *c = n->ctx;
  const size_t need = (size_t)nbytes * ((size_t)n->nranks + 1);
  if (need > n->gather_cap) {
    if (n->d_gather) cudaFree(n->d_gather);
    n->d_gather = nullptr; n->gather_cap = 0;
    EMD_CUDA(cudaMalloc((void **)&n->d_gather, need));
    n->gather_cap = need;
  }
  char *dsend = n->d_gather, *dall = n->d_gather + nbytes;
  EMD_CUDA(cudaMemcpyAsync(dsend, h_in, (size_t)nbytes, cudaMemcpyHostToDevice, c->stream));
  EMD_NCCL(g_nccl.AllGather(dsend, dall, (size_t)nbytes, ncclChar, n->comm, c->stream));
  EMD_CUDA(cudaMemcpyAsync(h_out_all, dall, (size_t)nbytes * (size_t)n->nranks, cudaMemcpyDeviceToHost, c->stream));
  EMD_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int emd_net_barrier(emd_net *n) {
  int one = 1;
  return emd_net_allreduce(n, &one, 1, 0, 0);
}

} // extern "C"
