// Kokkos_Core.hpp -- a small host-only stand-in for the part of the Kokkos 3.x API that ExaMiniMD
// uses, written for ONE purpose: compiling the UNMODIFIED reference sources under /root/reference
// into oracle/_ref/ so that the CPU oracle (oracle/*.c) and the CUDA product can be checked against
// the reference's own code.  TEST INFRASTRUCTURE ONLY -- nothing in examinimd_b200/ includes it.
//
// Kokkos itself (kokkos/kokkos, required ">= 3.0" by the reference's CMakeLists.txt:7) is not
// installed in this image and cannot be fetched.  What is restated here is its documented host
// semantics: Views are reference-counted LayoutRight arrays (the host default), every policy runs
// on the calling thread in index order (the reference's 1-thread behaviour, which is the
// deterministic order the oracle pins), or -- with -DKOKKOS_SHIM_OPENMP -- RangePolicy/league
// loops are split over OpenMP threads with Atomic-trait views and atomic_fetch_add made atomic,
// which is what the reference's OpenMP back-end does.  Team policies have team_size 1 and vector
// length 1, exactly as Kokkos' Serial/OpenMP back-ends give on CPUs.
#ifndef KOKKOS_SHIM_CORE_HPP
#define KOKKOS_SHIM_CORE_HPP

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>
#ifdef KOKKOS_SHIM_OPENMP
#include <omp.h>
#endif

#include <Kokkos_Macros.hpp>
#include <limits>
#include <cstdint>

namespace Kokkos {

struct LayoutRight {};
struct LayoutLeft {};
struct HostSpace { typedef HostSpace memory_space; static const char *name() { return "Host"; } };

struct Serial {
  typedef HostSpace memory_space;
  typedef Serial execution_space;
  static int concurrency() { return 1; }
  void print_configuration(std::ostream &os, bool = false) const { os << "Kokkos shim: Serial (host, 1 thread)\n"; }
  static const char *name() { return "Serial"; }
  void fence() const {}
};
struct OpenMP {
  typedef HostSpace memory_space;
  typedef OpenMP execution_space;
  static int concurrency() {
#ifdef KOKKOS_SHIM_OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
  }
  void print_configuration(std::ostream &os, bool = false) const { os << "Kokkos shim: OpenMP (host, " << concurrency() << " threads)\n"; }
  static const char *name() { return "OpenMP"; }
  void fence() const {}
};
#ifdef KOKKOS_SHIM_OPENMP
typedef OpenMP DefaultExecutionSpace;
#else
typedef Serial DefaultExecutionSpace;
#endif
typedef DefaultExecutionSpace DefaultHostExecutionSpace;

enum MemoryTraitsFlags { Unmanaged = 0x01, RandomAccess = 0x02, Atomic = 0x04, Restrict = 0x08, Aligned = 0x10 };
template <unsigned F>
struct MemoryTraits { enum : unsigned { flags = F }; };

struct ALL_t { constexpr ALL_t operator()() const { return ALL_t(); } };
constexpr ALL_t ALL = ALL_t();
struct AUTO_t {};
constexpr AUTO_t AUTO = AUTO_t();

template <class A, class B>
struct pair {
  A first; B second;
  pair() : first(), second() {}
  pair(const A &a, const B &b) : first(a), second(b) {}
  template <class C, class D> pair(const std::pair<C, D> &p) : first(p.first), second(p.second) {}
  bool operator==(const pair &o) const { return first == o.first && second == o.second; }
};

inline void initialize(int &, char **) {}
inline void initialize() {}
inline void finalize() {}
inline void fence() {}
inline void fence(const std::string &) {}
[[noreturn]] inline void abort(const char *msg) { fprintf(stderr, "Kokkos::abort: %s\n", msg); ::abort(); }

namespace Profiling {
inline void pushRegion(const std::string &) {}
inline void popRegion() {}
} // namespace Profiling

class Timer {
  std::chrono::steady_clock::time_point t0;
public:
  Timer() { reset(); }
  void reset() { t0 = std::chrono::steady_clock::now(); }
  double seconds() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

// ------------------------------------------------------------------------------ atomics
template <class T>
inline T atomic_fetch_add(T *p, const T &v) {
#ifdef KOKKOS_SHIM_OPENMP
  if constexpr (std::is_integral<T>::value) return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
  else { T old; _Pragma("omp critical(kokkos_shim_atomic)") { old = *p; *p = old + v; } return old; }
#else
  T old = *p; *p = old + v; return old;
#endif
}
template <class T>
inline void atomic_add(T *p, const T &v) {
#ifdef KOKKOS_SHIM_OPENMP
  _Pragma("omp atomic") *p += v;
#else
  *p += v;
#endif
}
template <class T>
inline void atomic_increment(T *p) { atomic_add(p, T(1)); }

// element handle of a View with the Atomic memory trait
template <class T>
struct AtomicRef {
  T *p;
  explicit AtomicRef(T *p_) : p(p_) {}
  operator T() const { return *p; }
  T operator=(const T &v) const { *p = v; return v; }
  void operator+=(const T &v) const { atomic_add(p, v); }
  void operator-=(const T &v) const { atomic_add(p, T(-v)); }
  T operator++(int) const { return atomic_fetch_add(p, T(1)); }
  T operator++() const { return atomic_fetch_add(p, T(1)) + T(1); }
};

// --------------------------------------------------------------------------------- View
namespace Impl {
// DataType grammar: T, T*, T**, T*[3], const T*[3], T***[3] ...  In C++ `T*[3]` is "array of 3 (T*)", i.e. the static
// extents are the OUTERMOST type constructors; Kokkos puts them AFTER the dynamic (pointer) extents in index order.
template <class D> struct DataAnalysis { typedef D value_type; enum { rank = 0, dyn = 0 }; static void statics(size_t *, int) {} };
template <class D> struct DataAnalysis<D *> {
  typedef typename DataAnalysis<D>::value_type value_type;
  enum { rank = DataAnalysis<D>::rank + 1, dyn = DataAnalysis<D>::dyn + 1 };
  static void statics(size_t *e, int pos) { DataAnalysis<D>::statics(e, pos); }
};
template <class D, size_t N> struct DataAnalysis<D[N]> {
  typedef typename DataAnalysis<D>::value_type value_type;
  enum { rank = DataAnalysis<D>::rank + 1, dyn = DataAnalysis<D>::dyn };
  static void statics(size_t *e, int pos) { e[pos] = N; DataAnalysis<D>::statics(e, pos + 1); }
};
template <class... P> struct HasAtomic : std::false_type {};
template <unsigned F, class... P> struct HasAtomic<MemoryTraits<F>, P...> : std::integral_constant<bool, (F & Atomic) != 0 || HasAtomic<P...>::value> {};
template <class X, class... P> struct HasAtomic<X, P...> : HasAtomic<P...> {};
} // namespace Impl

struct ScratchSpace; // forward (team scratch allocator)

template <class DataType, class... Props>
class View {
public:
  typedef Impl::DataAnalysis<DataType> analysis;
  typedef typename analysis::value_type value_type;
  typedef typename std::remove_const<value_type>::type non_const_value_type;
  typedef const non_const_value_type const_value_type;
  enum { rank = analysis::rank, dynamic_rank = analysis::dyn, is_atomic = Impl::HasAtomic<Props...>::value };
  enum { Rank = rank };
  typedef View HostMirror;
  typedef HostSpace memory_space;
  typedef DefaultExecutionSpace execution_space;
  typedef LayoutRight array_layout;
  typedef typename std::conditional<is_atomic, AtomicRef<value_type>, value_type &>::type reference_type;
  typedef value_type *pointer_type;
  typedef int size_type;

  value_type *ptr_;
  size_t ext_[8];
  std::shared_ptr<void> track_;
  std::string label_;

  View() : ptr_(nullptr) { for (int d = 0; d < 8; d++) ext_[d] = d < rank ? 0 : 1; set_statics(); }

  // allocating constructor: label + dynamic extents
  explicit View(const std::string &label, size_t n0 = 0, size_t n1 = 0, size_t n2 = 0, size_t n3 = 0, size_t n4 = 0, size_t n5 = 0,
                size_t n6 = 0, size_t n7 = 0) : ptr_(nullptr), label_(label) {
    set_extents(n0, n1, n2, n3, n4, n5, n6, n7);
    allocate();
  }
  explicit View(const char *label, size_t n0 = 0, size_t n1 = 0, size_t n2 = 0, size_t n3 = 0, size_t n4 = 0, size_t n5 = 0, size_t n6 = 0,
                size_t n7 = 0) : View(std::string(label), n0, n1, n2, n3, n4, n5, n6, n7) {}
  // wrapping constructor: existing memory
  View(value_type *p, size_t n0 = 0, size_t n1 = 0, size_t n2 = 0, size_t n3 = 0, size_t n4 = 0, size_t n5 = 0, size_t n6 = 0, size_t n7 = 0)
      : ptr_(p) { set_extents(n0, n1, n2, n3, n4, n5, n6, n7); }
  // scratch constructor (team.team_scratch(level), extents...)
  inline View(const ScratchSpace &s, size_t n0 = 0, size_t n1 = 0, size_t n2 = 0, size_t n3 = 0, size_t n4 = 0, size_t n5 = 0, size_t n6 = 0,
              size_t n7 = 0);

  // converting copy: same rank, compatible value type (adds const / changes traits)
  template <class D2, class... P2, class = typename std::enable_if<(int)View<D2, P2...>::rank == (int)rank &&
      std::is_convertible<typename View<D2, P2...>::value_type *, value_type *>::value>::type>
  View(const View<D2, P2...> &o) : ptr_(o.ptr_), track_(o.track_), label_(o.label_) { for (int d = 0; d < 8; d++) ext_[d] = o.ext_[d]; }

  // rank-1 sub-range (view, pair) and rank-0 element (view, index)
  template <class D2, class... P2, class I, class = typename std::enable_if<(int)View<D2, P2...>::rank == 1 && rank == 1>::type>
  View(const View<D2, P2...> &o, const std::pair<I, I> &r) : ptr_(o.ptr_ + r.first), track_(o.track_), label_(o.label_) {
    for (int d = 0; d < 8; d++) ext_[d] = 1;
    ext_[0] = (size_t)(r.second - r.first);
  }
  template <class D2, class... P2, class I, class = typename std::enable_if<(int)View<D2, P2...>::rank == 1 && rank == 1>::type>
  View(const View<D2, P2...> &o, const Kokkos::pair<I, I> &r) : ptr_(o.ptr_ + r.first), track_(o.track_), label_(o.label_) {
    for (int d = 0; d < 8; d++) ext_[d] = 1;
    ext_[0] = (size_t)(r.second - r.first);
  }
  template <class D2, class... P2, class I, class = typename std::enable_if<(int)View<D2, P2...>::rank == 1 && rank == 0 && std::is_integral<I>::value>::type>
  View(const View<D2, P2...> &o, const I &i) : ptr_(o.ptr_ + i), track_(o.track_), label_(o.label_) { for (int d = 0; d < 8; d++) ext_[d] = 1; }

  // row of a rank-2 view: View<T*>(v2, i, ALL)
  template <class D2, class... P2, class I, class = typename std::enable_if<(int)View<D2, P2...>::rank == 2 && rank == 1 && std::is_integral<I>::value>::type>
  View(const View<D2, P2...> &o, const I &i, const ALL_t &) : ptr_(o.ptr_ + (size_t)i * o.ext_[1]), track_(o.track_), label_(o.label_) {
    for (int d = 0; d < 8; d++) ext_[d] = 1;
    ext_[0] = o.ext_[1];
  }

  size_t extent(int d) const { return d < 8 ? ext_[d] : 1; }
  int extent_int(int d) const { return (int)extent(d); }
  size_t size() const { size_t s = 1; for (int d = 0; d < rank; d++) s *= ext_[d]; return s; }
  size_t span() const { return size(); }
  value_type *data() const { return ptr_; }
  value_type *ptr_on_device() const { return ptr_; }
  const std::string &label() const { return label_; }
  bool is_allocated() const { return ptr_ != nullptr; }
  int use_count() const { return (int)track_.use_count(); }
  size_t stride(int d) const { size_t s = 1; for (int k = d + 1; k < rank; k++) s *= ext_[k]; return s; }
  size_t stride_0() const { return stride(0); }
  size_t stride_1() const { return stride(1); }

  static size_t shmem_size(size_t n0 = 1, size_t n1 = 1, size_t n2 = 1, size_t n3 = 1, size_t n4 = 1, size_t n5 = 1, size_t n6 = 1, size_t n7 = 1) {
    View tmp;
    tmp.set_extents(n0, n1, n2, n3, n4, n5, n6, n7);
    return tmp.size() * sizeof(value_type) + 8; // + alignment slack, like Kokkos
  }

  reference_type make_ref(size_t off) const {
    if constexpr (is_atomic) return AtomicRef<value_type>(ptr_ + off);
    else return ptr_[off];
  }
  reference_type operator()() const { return make_ref(0); }
  template <class I0> reference_type operator()(const I0 &i0) const { return make_ref((size_t)i0); }
  template <class I0> reference_type operator[](const I0 &i0) const { return make_ref((size_t)i0); }
  template <class I0, class I1> reference_type operator()(const I0 &i0, const I1 &i1) const { return make_ref((size_t)i0 * ext_[1] + (size_t)i1); }
  template <class I0, class I1, class I2> reference_type operator()(const I0 &i0, const I1 &i1, const I2 &i2) const {
    return make_ref(((size_t)i0 * ext_[1] + (size_t)i1) * ext_[2] + (size_t)i2);
  }
  template <class I0, class I1, class I2, class I3> reference_type operator()(const I0 &i0, const I1 &i1, const I2 &i2, const I3 &i3) const {
    return make_ref((((size_t)i0 * ext_[1] + (size_t)i1) * ext_[2] + (size_t)i2) * ext_[3] + (size_t)i3);
  }
  template <class I0, class I1, class I2, class I3, class I4>
  reference_type operator()(const I0 &i0, const I1 &i1, const I2 &i2, const I3 &i3, const I4 &i4) const {
    return make_ref(((((size_t)i0 * ext_[1] + (size_t)i1) * ext_[2] + (size_t)i2) * ext_[3] + (size_t)i3) * ext_[4] + (size_t)i4);
  }
  template <class I0, class I1, class I2, class I3, class I4, class I5>
  reference_type operator()(const I0 &i0, const I1 &i1, const I2 &i2, const I3 &i3, const I4 &i4, const I5 &i5) const {
    return make_ref((((((size_t)i0 * ext_[1] + (size_t)i1) * ext_[2] + (size_t)i2) * ext_[3] + (size_t)i3) * ext_[4] + (size_t)i4) * ext_[5] + (size_t)i5);
  }

  // used by resize/realloc and the scratch constructor
  void set_extents(size_t n0, size_t n1, size_t n2, size_t n3, size_t n4, size_t n5, size_t n6, size_t n7) {
    const size_t n[8] = {n0, n1, n2, n3, n4, n5, n6, n7};
    for (int d = 0; d < 8; d++) ext_[d] = 1;
    for (int d = 0; d < dynamic_rank; d++) ext_[d] = n[d];
    set_statics();
  }
  void allocate() {
    const size_t bytes = size() * sizeof(value_type);
    void *p = bytes ? calloc(bytes, 1) : nullptr; // Kokkos zero-initialises
    if (bytes && !p) Kokkos::abort("Kokkos shim: allocation failed");
    track_ = std::shared_ptr<void>(p, free);
    ptr_ = static_cast<value_type *>(p);
  }

private:
  void set_statics() { analysis::statics(ext_, dynamic_rank); }
};

// ------------------------------------------------------------------- scratch memory
struct ScratchSpace {
  mutable char *cur; char *end;
  ScratchSpace() : cur(nullptr), end(nullptr) {}
  ScratchSpace(char *b, char *e) : cur(b), end(e) {}
  void *get(size_t bytes) const {
    uintptr_t a = (reinterpret_cast<uintptr_t>(cur) + 7) & ~uintptr_t(7);
    char *p = reinterpret_cast<char *>(a);
    if (p + bytes > end) Kokkos::abort("Kokkos shim: team scratch exhausted");
    cur = p + bytes;
    return p;
  }
};
template <class DataType, class... Props>
inline View<DataType, Props...>::View(const ScratchSpace &s, size_t n0, size_t n1, size_t n2, size_t n3, size_t n4, size_t n5, size_t n6, size_t n7)
    : ptr_(nullptr) {
  set_extents(n0, n1, n2, n3, n4, n5, n6, n7);
  ptr_ = static_cast<value_type *>(s.get(size() * sizeof(value_type)));
}

// ------------------------------------------------------------ mirrors, copies, resizing
template <class V> inline V create_mirror_view(const V &v) { return v; }
template <class S, class V> inline V create_mirror_view(const S &, const V &v) { return v; }
template <class V> inline V create_mirror(const V &v) { V m(v.label(), v.extent(0), v.extent(1), v.extent(2), v.extent(3), v.extent(4), v.extent(5), v.extent(6), v.extent(7)); return m; }

template <class D1, class... P1, class D2, class... P2>
inline void deep_copy(const View<D1, P1...> &dst, const View<D2, P2...> &src) {
  typedef typename View<D1, P1...>::non_const_value_type T;
  if ((const void *)dst.data() == (const void *)src.data()) return;
  if (dst.size() != src.size()) Kokkos::abort("Kokkos shim: deep_copy extent mismatch");
  T *d = const_cast<T *>(dst.data());
  for (size_t k = 0, n = dst.size(); k < n; k++) d[k] = src.data()[k];
}
template <class D1, class... P1>
inline void deep_copy(const View<D1, P1...> &dst, const typename View<D1, P1...>::non_const_value_type &v) {
  typedef typename View<D1, P1...>::non_const_value_type T;
  T *d = const_cast<T *>(dst.data());
  const size_t n = dst.size();
#ifdef KOKKOS_SHIM_OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (size_t k = 0; k < n; k++) d[k] = v;
}
template <class T, class D2, class... P2, class = typename std::enable_if<View<D2, P2...>::rank == 0 && std::is_arithmetic<T>::value>::type>
inline void deep_copy(T &dst, const View<D2, P2...> &src) { dst = *src.data(); }
template <class E, class D1, class... P1, class D2, class... P2, class = typename E::execution_space>
inline void deep_copy(const E &, const View<D1, P1...> &dst, const View<D2, P2...> &src) { deep_copy(dst, src); }

template <class D, class... P>
inline void realloc(View<D, P...> &v, size_t n0 = 0, size_t n1 = 0, size_t n2 = 0, size_t n3 = 0, size_t n4 = 0, size_t n5 = 0, size_t n6 = 0, size_t n7 = 0) {
  v = View<D, P...>(v.label(), n0, n1, n2, n3, n4, n5, n6, n7);
}
template <class D, class... P>
inline void resize(View<D, P...> &v, size_t n0 = 0, size_t n1 = 0, size_t n2 = 0, size_t n3 = 0, size_t n4 = 0, size_t n5 = 0, size_t n6 = 0, size_t n7 = 0) {
  typedef View<D, P...> V;
  V nv(v.label(), n0, n1, n2, n3, n4, n5, n6, n7);
  bool same = true;
  for (int d = 0; d < V::rank; d++) same = same && nv.extent(d) == v.extent(d);
  if (same) return;
  // copy the overlapping index box (Kokkos::resize preserves content)
  size_t m[8], idx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  size_t total = 1;
  for (int d = 0; d < 8; d++) { m[d] = d < V::rank ? std::min(nv.extent(d), v.extent(d)) : 1; total *= m[d]; }
  typename V::non_const_value_type *dst = const_cast<typename V::non_const_value_type *>(nv.data());
  for (size_t c = 0; c < total; c++) {
    size_t so = 0, dof = 0;
    for (int d = 0; d < V::rank; d++) { so = so * v.extent(d) + idx[d]; dof = dof * nv.extent(d) + idx[d]; }
    dst[dof] = v.data()[so];
    for (int d = V::rank - 1; d >= 0; d--) { if (++idx[d] < m[d]) break; idx[d] = 0; }
  }
  v = nv;
}

// subview: (rows range, ALL) of a rank-2 view and (row, ALL) of a rank-2 view -- the two forms the reference uses
template <class D, class... P, class I>
inline View<D, P...> subview(const View<D, P...> &v, const std::pair<I, I> &r, ALL_t) {
  static_assert(View<D, P...>::rank == 2, "shim subview(range, ALL): rank-2 only");
  View<D, P...> s(v);
  s.ptr_ = v.ptr_ + (size_t)r.first * v.extent(1);
  s.ext_[0] = (size_t)(r.second - r.first);
  return s;
}
template <class D, class... P, class I>
inline View<D, P...> subview(const View<D, P...> &v, const Kokkos::pair<I, I> &r, ALL_t) { return subview(v, std::pair<I, I>(r.first, r.second), ALL); }
template <class D, class... P, class I, class = typename std::enable_if<std::is_integral<I>::value>::type>
inline View<typename View<D, P...>::value_type *, LayoutRight> subview(const View<D, P...> &v, const I &row, ALL_t) {
  static_assert(View<D, P...>::rank == 2, "shim subview(row, ALL): rank-2 only");
  View<typename View<D, P...>::value_type *, LayoutRight> s;
  s.ptr_ = v.ptr_ + (size_t)row * v.extent(1);
  s.ext_[0] = v.extent(1);
  s.track_ = v.track_;
  return s;
}

// single elements as rank-0 views: subview(v1, i), subview(v2, i, j)
template <class D, class... P, class I, class = typename std::enable_if<std::is_integral<I>::value && View<D, P...>::rank == 1>::type>
inline View<typename View<D, P...>::value_type> subview(const View<D, P...> &v, const I &i) {
  View<typename View<D, P...>::value_type> s;
  s.ptr_ = v.ptr_ + (size_t)i;
  s.track_ = v.track_;
  return s;
}
template <class D, class... P, class I, class J, class = typename std::enable_if<std::is_integral<I>::value && std::is_integral<J>::value && View<D, P...>::rank == 2>::type>
inline View<typename View<D, P...>::value_type> subview(const View<D, P...> &v, const I &i, const J &j) {
  View<typename View<D, P...>::value_type> s;
  s.ptr_ = v.ptr_ + (size_t)i * v.extent(1) + (size_t)j;
  s.track_ = v.track_;
  return s;
}

// ------------------------------------------------------------------------------ policies
template <class T> struct IndexType { typedef T type; };
struct Dynamic {}; struct Static {};
template <class T> struct Schedule {};
struct ParallelForTag {}; struct ParallelReduceTag {}; struct ParallelScanTag {};

namespace Impl {
template <class T> struct IsPolicyTrait : std::false_type {};
template <class T> struct IsPolicyTrait<IndexType<T>> : std::true_type {};
template <class T> struct IsPolicyTrait<Schedule<T>> : std::true_type {};
template <> struct IsPolicyTrait<Serial> : std::true_type {};
template <> struct IsPolicyTrait<OpenMP> : std::true_type {};
template <class... P> struct WorkTagOf { typedef void type; };
template <class A, class... P> struct WorkTagOf<A, P...> {
  typedef typename std::conditional<IsPolicyTrait<A>::value, typename WorkTagOf<P...>::type, A>::type type;
};
} // namespace Impl

template <class... Props>
struct RangePolicy {
  typedef typename Impl::WorkTagOf<Props...>::type work_tag;
  typedef long member_type;
  long b, e;
  RangePolicy(long b_, long e_) : b(b_), e(e_) {}
  long begin() const { return b; }
  long end() const { return e; }
};

struct PerTeamValue { size_t v; };
struct PerThreadValue { size_t v; };
inline PerTeamValue PerTeam(size_t v) { return PerTeamValue{v}; }
inline PerThreadValue PerThread(size_t v) { return PerThreadValue{v}; }

class HostTeamMember {
public:
  int league_rank_, league_size_;
  char *team_b, *team_e, *thread_b, *thread_e;
  mutable ScratchSpace team_s, thread_s;
  HostTeamMember(int lr, int ls, char *tb, char *te, char *hb, char *he)
      : league_rank_(lr), league_size_(ls), team_b(tb), team_e(te), thread_b(hb), thread_e(he), team_s(tb, te), thread_s(hb, he) {}
  int league_rank() const { return league_rank_; }
  int league_size() const { return league_size_; }
  int team_rank() const { return 0; }
  int team_size() const { return 1; }
  void team_barrier() const {}
  const ScratchSpace &team_scratch(int) const { return team_s; }
  const ScratchSpace &thread_scratch(int) const { return thread_s; }
  const ScratchSpace &team_shmem() const { return team_s; }
};
struct PerTeamTag { const HostTeamMember &t; };
struct PerThreadTag { const HostTeamMember &t; };
inline PerTeamTag PerTeam(const HostTeamMember &t) { return PerTeamTag{t}; }
inline PerThreadTag PerThread(const HostTeamMember &t) { return PerThreadTag{t}; }
template <class F> inline void single(const PerTeamTag &, const F &f) { f(); }
template <class F> inline void single(const PerThreadTag &, const F &f) { f(); }
template <class F, class T> inline void single(const PerTeamTag &, const F &f, T &v) { f(v); }
template <class F, class T> inline void single(const PerThreadTag &, const F &f, T &v) { f(v); }

template <class... Props>
struct TeamPolicy {
  typedef typename Impl::WorkTagOf<Props...>::type work_tag;
  typedef HostTeamMember member_type;
  int league, team, vec;
  size_t scratch_team[2], scratch_thread[2];
  void init(int l) { league = l; team = 1; vec = 1; scratch_team[0] = scratch_team[1] = scratch_thread[0] = scratch_thread[1] = 0; }
  TeamPolicy(int l, int, int = 1) { init(l); }
  TeamPolicy(int l, AUTO_t, int = 1) { init(l); }
  TeamPolicy(int l, int, AUTO_t) { init(l); }
  TeamPolicy(int l, AUTO_t, AUTO_t) { init(l); }
  int league_size() const { return league; }
  int team_size() const { return 1; }
  TeamPolicy &set_scratch_size(int level, const PerTeamValue &a) { scratch_team[level] = a.v; return *this; }
  TeamPolicy &set_scratch_size(int level, const PerThreadValue &a) { scratch_thread[level] = a.v; return *this; }
  TeamPolicy &set_scratch_size(int level, const PerTeamValue &a, const PerThreadValue &b) { scratch_team[level] = a.v; scratch_thread[level] = b.v; return *this; }
  TeamPolicy &set_scratch_size(int level, const PerThreadValue &b, const PerTeamValue &a) { scratch_team[level] = a.v; scratch_thread[level] = b.v; return *this; }
  template <class F, class Tag> int team_size_max(const F &, const Tag &) const { return 1; }
  template <class F, class Tag> int team_size_recommended(const F &, const Tag &) const { return 1; }
};

struct TeamRange { long b, e; };
inline TeamRange TeamThreadRange(const HostTeamMember &, long n) { return TeamRange{0, n}; }
inline TeamRange TeamThreadRange(const HostTeamMember &, long b, long e) { return TeamRange{b, e}; }
inline TeamRange ThreadVectorRange(const HostTeamMember &, long n) { return TeamRange{0, n}; }
inline TeamRange ThreadVectorRange(const HostTeamMember &, long b, long e) { return TeamRange{b, e}; }
inline TeamRange TeamVectorRange(const HostTeamMember &, long n) { return TeamRange{0, n}; }

// reducers
template <class T> struct Max { T &ref; typedef T value_type; explicit Max(T &r) : ref(r) {}
  static T identity() { return std::numeric_limits<T>::lowest(); } static void join(T &a, const T &b) { if (b > a) a = b; } };
template <class T> struct Min { T &ref; typedef T value_type; explicit Min(T &r) : ref(r) {}
  static T identity() { return std::numeric_limits<T>::max(); } static void join(T &a, const T &b) { if (b < a) a = b; } };
template <class T> struct Sum { T &ref; typedef T value_type; explicit Sum(T &r) : ref(r) {}
  static T identity() { return T(); } static void join(T &a, const T &b) { a += b; } };
namespace Impl {
template <class R> struct IsReducer : std::false_type {};
template <class T> struct IsReducer<Max<T>> : std::true_type {};
template <class T> struct IsReducer<Min<T>> : std::true_type {};
template <class T> struct IsReducer<Sum<T>> : std::true_type {};

template <class Tag, class F, class... A>
inline typename std::enable_if<std::is_void<Tag>::value>::type call(const F &f, A &&...a) { f(std::forward<A>(a)...); }
template <class Tag, class F, class... A>
inline typename std::enable_if<!std::is_void<Tag>::value>::type call(const F &f, A &&...a) { f(Tag(), std::forward<A>(a)...); }
} // namespace Impl

// -------- nested (team-level) patterns: serial loops, as on every Kokkos CPU back-end
template <class F> inline void parallel_for(const TeamRange &r, const F &f) { for (long i = r.b; i < r.e; i++) f(i); }
template <class F, class T, class = typename std::enable_if<!Impl::IsReducer<T>::value>::type>
inline void parallel_reduce(const TeamRange &r, const F &f, T &result) { T v = T(); for (long i = r.b; i < r.e; i++) f(i, v); result = v; }
template <class F, class T>
inline void parallel_reduce(const TeamRange &r, const F &f, const Max<T> &red) { T v = Max<T>::identity(); for (long i = r.b; i < r.e; i++) f(i, v); red.ref = v; }
template <class F> inline void parallel_scan(const TeamRange &r, const F &f) {
  // value type of the reference's only nested scan is int (force_snap_neigh_impl.h:633-656)
  int acc = 0;
  for (long i = r.b; i < r.e; i++) f(i, acc, true);
}

// ------------------------------------------------ top-level patterns: RangePolicy
template <class... P, class F>
inline void parallel_for(const RangePolicy<P...> &p, const F &f) {
  typedef typename RangePolicy<P...>::work_tag Tag;
#ifdef KOKKOS_SHIM_OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (long i = p.b; i < p.e; i++) Impl::call<Tag>(f, i);
}
template <class... P, class F> inline void parallel_for(const std::string &, const RangePolicy<P...> &p, const F &f) { parallel_for(p, f); }
template <class I, class F, class = typename std::enable_if<std::is_integral<I>::value>::type>
inline void parallel_for(const I &n, const F &f) { parallel_for(RangePolicy<>(0, (long)n), f); }
template <class I, class F, class = typename std::enable_if<std::is_integral<I>::value>::type>
inline void parallel_for(const std::string &, const I &n, const F &f) { parallel_for(RangePolicy<>(0, (long)n), f); }

template <class... P, class F, class T, class = typename std::enable_if<!Impl::IsReducer<T>::value>::type>
inline void parallel_reduce(const RangePolicy<P...> &p, const F &f, T &result) {
  typedef typename RangePolicy<P...>::work_tag Tag;
  T total = T();
#ifdef KOKKOS_SHIM_OPENMP
#pragma omp parallel
  {
    T mine = T();
#pragma omp for schedule(static) nowait
    for (long i = p.b; i < p.e; i++) Impl::call<Tag>(f, i, mine);
#pragma omp critical(kokkos_shim_reduce)
    total += mine;
  }
#else
  for (long i = p.b; i < p.e; i++) Impl::call<Tag>(f, i, total);
#endif
  result = total;
}
template <class... P, class F, class R, class = typename std::enable_if<Impl::IsReducer<R>::value>::type, class = void>
inline void parallel_reduce(const RangePolicy<P...> &p, const F &f, const R &red) {
  typedef typename RangePolicy<P...>::work_tag Tag;
  typename R::value_type v = R::identity();
  for (long i = p.b; i < p.e; i++) Impl::call<Tag>(f, i, v);
  red.ref = v;
}
template <class... P, class F, class T> inline void parallel_reduce(const std::string &, const RangePolicy<P...> &p, const F &f, T &&r) { parallel_reduce(p, f, std::forward<T>(r)); }
template <class I, class F, class T, class = typename std::enable_if<std::is_integral<I>::value>::type>
inline void parallel_reduce(const I &n, const F &f, T &&r) { parallel_reduce(RangePolicy<>(0, (long)n), f, std::forward<T>(r)); }
template <class I, class F, class T, class = typename std::enable_if<std::is_integral<I>::value>::type>
inline void parallel_reduce(const std::string &, const I &n, const F &f, T &&r) { parallel_reduce(RangePolicy<>(0, (long)n), f, std::forward<T>(r)); }

// exclusive-scan protocol: f(i, update, final) is called with final = true once per i in order
template <class... P, class F>
inline void parallel_scan(const RangePolicy<P...> &p, const F &f) {
  typedef typename RangePolicy<P...>::work_tag Tag;
  // every scan functor of the reference accumulates T_INT
  int acc = 0;
  for (long i = p.b; i < p.e; i++) Impl::call<Tag>(f, i, acc, true);
}
template <class... P, class F> inline void parallel_scan(const std::string &, const RangePolicy<P...> &p, const F &f) { parallel_scan(p, f); }
template <class I, class F, class = typename std::enable_if<std::is_integral<I>::value>::type>
inline void parallel_scan(const std::string &, const I &n, const F &f) { parallel_scan(RangePolicy<>(0, (long)n), f); }

// ------------------------------------------------ top-level patterns: TeamPolicy
template <class... P, class F>
inline void parallel_for(const TeamPolicy<P...> &p, const F &f) {
  typedef typename TeamPolicy<P...>::work_tag Tag;
  const size_t tb = p.scratch_team[0] + p.scratch_team[1] + 64, hb = p.scratch_thread[0] + p.scratch_thread[1] + 64;
#ifdef KOKKOS_SHIM_OPENMP
#pragma omp parallel
#endif
  {
    std::vector<char> team_buf(tb), thread_buf(hb);
#ifdef KOKKOS_SHIM_OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
    for (int l = 0; l < p.league; l++) {
      HostTeamMember m(l, p.league, team_buf.data(), team_buf.data() + tb, thread_buf.data(), thread_buf.data() + hb);
      Impl::call<Tag>(f, m);
    }
  }
}
template <class... P, class F> inline void parallel_for(const std::string &, const TeamPolicy<P...> &p, const F &f) { parallel_for(p, f); }

template <class... P, class F, class T>
inline void parallel_reduce(const TeamPolicy<P...> &p, const F &f, T &result) {
  typedef typename TeamPolicy<P...>::work_tag Tag;
  const size_t tb = p.scratch_team[0] + p.scratch_team[1] + 64, hb = p.scratch_thread[0] + p.scratch_thread[1] + 64;
  T total = T();
#ifdef KOKKOS_SHIM_OPENMP
#pragma omp parallel
#endif
  {
    std::vector<char> team_buf(tb), thread_buf(hb);
    T mine = T();
#ifdef KOKKOS_SHIM_OPENMP
#pragma omp for schedule(dynamic, 1) nowait
#endif
    for (int l = 0; l < p.league; l++) {
      HostTeamMember m(l, p.league, team_buf.data(), team_buf.data() + tb, thread_buf.data(), thread_buf.data() + hb);
      Impl::call<Tag>(f, m, mine);
    }
#ifdef KOKKOS_SHIM_OPENMP
#pragma omp critical(kokkos_shim_reduce)
#endif
    total += mine;
  }
  result = total;
}
template <class... P, class F, class T> inline void parallel_reduce(const std::string &, const TeamPolicy<P...> &p, const F &f, T &r) { parallel_reduce(p, f, r); }

} // namespace Kokkos
#endif
