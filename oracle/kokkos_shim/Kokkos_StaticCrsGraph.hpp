// Kokkos_StaticCrsGraph.hpp (shim) -- the container the reference's NeighListCSR derives from
// (src/neighbor_types/neighbor_csr.h:81-124): two views, row_map[nrows+1] and entries[nnz].
#ifndef KOKKOS_SHIM_STATICCRSGRAPH_HPP
#define KOKKOS_SHIM_STATICCRSGRAPH_HPP
#include <Kokkos_Core.hpp>
namespace Kokkos {
template <class DataType, class Layout, class Space, class MemTraits = void, class SizeType = int>
class StaticCrsGraph {
public:
  typedef DataType data_type;
  typedef SizeType size_type;
  typedef View<const SizeType *, LayoutRight> row_map_type;
  typedef View<DataType *, LayoutRight> entries_type;
  entries_type entries;
  row_map_type row_map;
  StaticCrsGraph() {}
  template <class E, class R> StaticCrsGraph(const E &entries_, const R &row_map_) : entries(entries_), row_map(row_map_) {}
  size_t numRows() const { return row_map.extent(0) ? row_map.extent(0) - 1 : 0; }
};
} // namespace Kokkos
#endif
