// Test stub (tests/test_reference_seam.py): just enough of tbb::concurrent_hash_map for the
// reference's KMerCounter.cpp to COMPILE against host/GPUHandler.h. TBB is an un-vendored,
// unpinned dependency of the reference; it is not part of the hot path.
#pragma once
#include <map>
#include <utility>
namespace tbb {
template <class K, class V, class H>
class concurrent_hash_map {
public:
    typedef std::map<K, V> map_t;
    typedef typename map_t::iterator iterator;
    class accessor {
    public:
        std::pair<const K, V> *operator->() { return p; }
        std::pair<const K, V> *p = nullptr;
    };
    bool emplace(accessor &a, const K &k, const V &v) {
        auto r = m.emplace(k, v);
        a.p = &*r.first;
        return r.second;
    }
    size_t size() const { return m.size(); }
    iterator begin() { return m.begin(); }
    iterator end() { return m.end(); }
private:
    map_t m;
};
}  // namespace tbb
