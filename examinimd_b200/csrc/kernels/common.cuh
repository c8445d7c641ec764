// common.cuh -- context, error plumbing, scratch memory and scan helper shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <string>
#include "emd_b200.h"

namespace emd {

void set_error(const char *fmt, ...);

#define EMD_CUDA(call)                                                                     \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      emd::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

// grow-only device scratch buffer
struct Scratch {
  void *p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { p = nullptr; cap = 0; set_error("scratch cudaMalloc(%zu) -> %s", want, cudaGetErrorString(e)); return 1; }
    cap = want;
    return 0;
  }
  template <class T> T *as() { return reinterpret_cast<T *>(p); }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

constexpr int kMaxTypesConst = 12; // MAX_TYPES_STACKPARAMS, src/types.h:68

struct LJParams {
  int ntypes = 0;
  double lj1[kMaxTypesConst * kMaxTypesConst];
  double lj2[kMaxTypesConst * kMaxTypesConst];
  double cutsq[kMaxTypesConst * kMaxTypesConst];
};

} // namespace emd

struct emd_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  unsigned long long launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaStream_t side_stream = nullptr, main_stream = nullptr; // emd_ctx_side_*: main_stream != nullptr while the side stream is current
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool fork_marked = false;
  int side_sms = 0; // SMs of the side stream's green context (0: it shares the whole device)
  int num_sms = 148;
  // scratch
  emd::Scratch s_a, s_b, s_c, s_scan; // general-purpose device scratch
  int *h_pinned = nullptr;            // small pinned staging area for scalar read-backs (64 ints)
  // LJ parameters
  emd::LJParams lj;
  unsigned long long lj_version = 0; // bumped by emd_force_lj_set_params (device copies of the table are uploaded once per version)
  // Halo gate (comm_peer.cu): the neighbours' stores into my ghost rows are announced by sequence flags; a kernel that can
  // wait for them itself (the LJ tile force: only before its first tile that reads a ghost) takes the gate along, every
  // other consumer calls emd_ctx_halo_gate_wait first.
  const int *gate_flags = nullptr; // [6] arrived flags, by phase
  int gate_seq = 0, gate_mask = 0; // wanted sequence number; phases that count
  bool gate_pending = false;
  double *d_lj_tables = nullptr; // for ntypes > 12: [3][ntypes][ntypes]
  int lj_tables_ntypes = 0;
};

namespace emd {

// exclusive scan of d_in[0..n) into d_out[0..n); if d_total != nullptr, *d_total = sum.
// d_in may equal d_out.  Three launches (reduce / scan of block sums / scan) -- see scan.cu.
int exclusive_scan_int(emd_ctx *ctx, const int *d_in, int *d_out, int n, int *d_total);

inline int grid_for(long long n, int block) { return (int)((n + block - 1) / block); }

#define EMD_LAUNCH(ctx, kernel, grid, block, smem, ...)                   \
  do {                                                                    \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);      \
    (ctx)->launches++;                                                    \
    EMD_CUDA(cudaGetLastError());                                         \
  } while (0)

} // namespace emd
