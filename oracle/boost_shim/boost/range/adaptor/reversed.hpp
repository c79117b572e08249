#ifndef HELLO_ORACLE_BOOST_REVERSED_SHIM
#define HELLO_ORACLE_BOOST_REVERSED_SHIM
namespace boost { namespace adaptors {
template <class C> struct shim_reversed {
    C& c;
    auto begin() const -> decltype(c.rbegin()) { return c.rbegin(); }
    auto end() const -> decltype(c.rend()) { return c.rend(); }
};
template <class C> shim_reversed<C> reverse(C& c) { return shim_reversed<C>{c}; }
}}
#endif
