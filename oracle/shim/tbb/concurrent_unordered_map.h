// Test infrastructure only (see oracle/README in the header of oracle/Makefile): stands in for <tbb/concurrent_unordered_map.h>,
// which the reference's map_gpu.hpp:53 / filter_gpu.hpp:53 include for the key -> state table of the keyed-stateful GPU operators
// (find / insert / end from several replica threads). TBB is not installed in this image. A node-based std::unordered_map behind a
// mutex has the guarantees those call sites need: element addresses are stable, find and insert are serialised.
#pragma once
#include <mutex>
#include <unordered_map>
#include <utility>
namespace tbb {
template <class K, class V, class H = std::hash<K>, class E = std::equal_to<K>>
class concurrent_unordered_map {
    std::unordered_map<K, V, H, E> m;
    mutable std::mutex mu;
public:
    using iterator = typename std::unordered_map<K, V, H, E>::iterator;
    concurrent_unordered_map() { m.reserve(1u << 16); }
    iterator find(const K &k) { std::lock_guard<std::mutex> l(mu); return m.find(k); }
    iterator end() { return m.end(); }
    template <class P> std::pair<iterator, bool> insert(P &&p) { std::lock_guard<std::mutex> l(mu); return m.insert(std::forward<P>(p)); }
    size_t size() const { std::lock_guard<std::mutex> l(mu); return m.size(); }
};
}
