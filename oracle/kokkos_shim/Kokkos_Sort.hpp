// Kokkos_Sort.hpp (shim) -- Kokkos::BinSort / Kokkos::BinOp3D as BinningKKSort uses them
// (src/binning_types/binning_kksort.h:44-47, binning_kksort.cpp:103-137).  Semantics restated from
// Kokkos 3.x algorithms/src/Kokkos_Sort.hpp:
//   BinOp3D(max_bins[3], min[3], max[3]): mul_[d] = max_bins[d] / (max[d] - min[d]);
//     bin(keys,i) = ((int(mul0*(k(i,0)-min0)) * max_bins1 + int(mul1*(k(i,1)-min1))) * max_bins2) + int(mul2*(k(i,2)-min2))
//   BinSort(keys, bin_op, sort_within_bins = false); create_permute_vector(): histogram, exclusive
//     scan, then sort_order[offset[bin] + count[bin]++] = i  (arrival order = ascending i on one thread);
//   sort(values): values(i,:) = old values(sort_order(i),:) through a scratch copy.
// The two histogram/placement passes are kept serial so the permutation is the reference's
// deterministic 1-thread order even in the OpenMP build; the gather in sort() is parallel.
#ifndef KOKKOS_SHIM_SORT_HPP
#define KOKKOS_SHIM_SORT_HPP
#include <Kokkos_Core.hpp>
namespace Kokkos {

template <class KeyViewType>
struct BinOp3D {
  int max_bins_[3];
  double mul_[3];
  typename KeyViewType::non_const_value_type range_[3];
  typename KeyViewType::non_const_value_type min_[3];
  BinOp3D() {}
  BinOp3D(int max_bins__[], typename KeyViewType::const_value_type min[], typename KeyViewType::const_value_type max[]) {
    for (int d = 0; d < 3; d++) {
      max_bins_[d] = max_bins__[d];
      mul_[d] = 1.0 * max_bins__[d] / (max[d] - min[d]);
      range_[d] = max[d] - min[d];
      min_[d] = min[d];
    }
  }
  template <class ViewType>
  int bin(ViewType &keys, const int &i) const {
    return int((((int(mul_[0] * (keys(i, 0) - min_[0])) * max_bins_[1]) + int(mul_[1] * (keys(i, 1) - min_[1]))) * max_bins_[2]) +
               int(mul_[2] * (keys(i, 2) - min_[2])));
  }
  int max_bins() const { return max_bins_[0] * max_bins_[1] * max_bins_[2]; }
};

template <class KeyViewType, class BinSortOp, class Space = DefaultExecutionSpace, class SizeType = int>
class BinSort {
public:
  typedef View<SizeType *> offset_type;
  typedef View<int *> bin_count_type;
  typedef offset_type perm_type;
private:
  KeyViewType keys;
  BinSortOp bin_op;
  offset_type bin_offsets, sort_order;
  bin_count_type bin_count;
  int range_begin, range_end;
public:
  BinSort() : range_begin(0), range_end(0) {}
  BinSort(const KeyViewType &keys_, const BinSortOp &op, bool = false) : keys(keys_), bin_op(op), range_begin(0), range_end((int)keys_.extent(0)) {
    bin_count = bin_count_type("Kokkos::SortImpl::BinSortFunctor::bin_count", op.max_bins());
    bin_offsets = offset_type("Kokkos::SortImpl::BinSortFunctor::bin_offsets", op.max_bins());
    sort_order = offset_type("PermutationVector", range_end - range_begin);
  }
  void create_permute_vector() {
    const int nbins = bin_op.max_bins();
    for (int b = 0; b < nbins; b++) bin_count(b) = 0;
    for (int i = range_begin; i < range_end; i++) bin_count(bin_op.bin(keys, i))++;
    int acc = 0;
    for (int b = 0; b < nbins; b++) { bin_offsets(b) = acc; acc += bin_count(b); }
    for (int b = 0; b < nbins; b++) bin_count(b) = 0;
    for (int i = range_begin; i < range_end; i++) {
      const int b = bin_op.bin(keys, i);
      const int c = bin_count(b)++;
      sort_order(bin_offsets(b) + c) = i;
    }
  }
  template <class ValuesViewType>
  void sort(const ValuesViewType &values) const {
    typedef typename ValuesViewType::non_const_value_type T;
    const size_t n = (size_t)(range_end - range_begin);
    size_t cols = 1;
    for (int d = 1; d < (int)ValuesViewType::rank; d++) cols *= values.extent(d);
    std::vector<T> scratch(n * cols);
    T *v = const_cast<T *>(values.data());
#ifdef KOKKOS_SHIM_OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (size_t i = 0; i < n; i++) {
      const size_t s = (size_t)sort_order(i);
      for (size_t c = 0; c < cols; c++) scratch[i * cols + c] = v[s * cols + c];
    }
#ifdef KOKKOS_SHIM_OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (size_t k = 0; k < n * cols; k++) v[k] = scratch[k];
  }
  offset_type get_permute_vector() const { return sort_order; }
  offset_type get_bin_offsets() const { return bin_offsets; }
  bin_count_type get_bin_count() const { return bin_count; }
};
} // namespace Kokkos
#endif
