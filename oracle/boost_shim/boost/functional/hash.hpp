// Stand-in for <boost/functional/hash.hpp>: hash containers only need *a* hash; iteration order over them is not part
// of the encoder's output (computeFeaturesColoredSimple walks vectors).
#ifndef HELLO_ORACLE_BOOST_HASH_SHIM
#define HELLO_ORACLE_BOOST_HASH_SHIM
#include <cstddef>
#include <functional>
#include <string>
#include <utility>
namespace boost {
template <class T> inline void hash_combine(std::size_t& seed, const T& v) {
    seed ^= std::hash<T>()(v) + 0x9e3779b97f4a7c15ULL + (seed << 6) + (seed >> 2);
}
template <class T> struct hash { std::size_t operator()(const T& v) const { return std::hash<T>()(v); } };
template <class A, class B> struct hash<std::pair<A, B>> {
    std::size_t operator()(const std::pair<A, B>& p) const {
        std::size_t seed = 0;
        hash_combine(seed, p.first);
        hash_combine(seed, p.second);
        return seed;
    }
};
}
#endif
