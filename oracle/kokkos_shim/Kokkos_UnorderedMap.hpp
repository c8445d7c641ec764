// Kokkos_UnorderedMap.hpp (shim) -- the insert-only set NeighborCSRMapConstr uses
// (src/neighbor_types/neighbor_csr_map_constr.h:101,158,221): fixed-capacity open addressing,
// insert() fails when full so that the caller's rehash-and-redo loop works as with Kokkos.
#ifndef KOKKOS_SHIM_UNORDEREDMAP_HPP
#define KOKKOS_SHIM_UNORDEREDMAP_HPP
#include <Kokkos_Core.hpp>
namespace Kokkos {
struct UnorderedMapInsertResult {
  bool ok, existed;
  bool failed() const { return !ok; }
  bool success() const { return ok && !existed; }
  bool existing() const { return existed; }
};
template <class Key, class Value, class Space = HostSpace>
class UnorderedMap {
  struct Store { std::vector<Key> keys; std::vector<char> used; size_t count = 0; bool failed = false; };
  std::shared_ptr<Store> s;
  static size_t hash(const Key &k) { return (size_t)k.first * 0x9E3779B97F4A7C15ull ^ ((size_t)k.second * 0xC2B2AE3D27D4EB4Full); }
public:
  typedef UnorderedMapInsertResult insert_result;
  UnorderedMap(size_t cap = 0) : s(new Store) { rehash(cap); }
  size_t capacity() const { return s->keys.size(); }
  size_t size() const { return s->count; }
  bool rehash(size_t cap) { s->keys.assign(cap, Key()); s->used.assign(cap, 0); s->count = 0; s->failed = false; return true; }
  void clear() { std::fill(s->used.begin(), s->used.end(), 0); s->count = 0; s->failed = false; }
  bool failed_insert() const { return s->failed; }
  insert_result insert(const Key &k) const {
    insert_result r{false, false};
    const size_t cap = s->keys.size();
    if (cap == 0) { s->failed = true; return r; }
#ifdef KOKKOS_SHIM_OPENMP
#pragma omp critical(kokkos_shim_umap)
#endif
    {
      size_t h = hash(k) % cap;
      for (size_t probe = 0; probe < cap; probe++, h = (h + 1 == cap ? 0 : h + 1)) {
        if (!s->used[h]) {
          if (s->count * 10 >= cap * 9) break; // keep some head room like Kokkos (fails before 100 % full)
          s->used[h] = 1; s->keys[h] = k; s->count++; r.ok = true; break;
        }
        if (s->keys[h] == k) { r.ok = true; r.existed = true; break; }
      }
      if (!r.ok) s->failed = true;
    }
    return r;
  }
  bool valid_at(size_t i) const { return s->used[i] != 0; }
  Key key_at(size_t i) const { return s->keys[i]; }
};
} // namespace Kokkos
#endif
