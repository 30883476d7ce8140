// TEST INFRASTRUCTURE - not tsl::robin_map.  The interface FORM uses (find / insert / at /
// erase / iteration with it.value()) over std::unordered_map.  Only the iteration order
// differs from the real container, which the canonical rules R4 / R5 / R7 make irrelevant
// except on exact distance ties.
#pragma once
#include <functional>
#include <unordered_map>
namespace tsl {
template <typename K, typename V, typename H = std::hash<K>> class robin_map {
  using Base = std::unordered_map<K, V, H>;

public:
  class iterator : public Base::iterator {
  public:
    iterator() = default;
    iterator(typename Base::iterator it) : Base::iterator(it) {}
    V &value() const { return (*this)->second; }
    const K &key() const { return (*this)->first; }
  };
  class const_iterator : public Base::const_iterator {
  public:
    const_iterator() = default;
    const_iterator(typename Base::const_iterator it) : Base::const_iterator(it) {}
    const_iterator(iterator it) : Base::const_iterator(static_cast<typename Base::iterator>(it)) {}
    const V &value() const { return (*this)->second; }
    const K &key() const { return (*this)->first; }
  };
  iterator begin() { return m.begin(); }
  iterator end() { return m.end(); }
  const_iterator begin() const { return m.begin(); }
  const_iterator end() const { return m.end(); }
  iterator find(const K &k) { return m.find(k); }
  const_iterator find(const K &k) const { return m.find(k); }
  std::pair<iterator, bool> insert(std::pair<K, V> p) {
    auto r = m.insert(std::move(p));
    return {iterator(r.first), r.second};
  }
  V &at(const K &k) { return m.at(k); }
  const V &at(const K &k) const { return m.at(k); }
  void erase(iterator it) { m.erase(static_cast<typename Base::iterator>(it)); }
  std::size_t erase(const K &k) { return m.erase(k); }
  std::size_t size() const { return m.size(); }
  bool empty() const { return m.empty(); }
  void clear() { m.clear(); }

private:
  Base m;
};
} // namespace tsl
