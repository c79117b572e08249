// Stand-in for <boost/python.hpp> -- TEST INFRASTRUCTURE (oracle/), NOT A PRODUCT PATH.
//
// The reference's feature encoder (c++/src/AlleleSearcherLiteFiltered.cpp) is written against Boost.Python only for
// its argument containers (p::list of reads / qualities / CIGAR tuples) and its numpy return value.  Boost is not in
// this image, so oracle/Makefile compiles the reference's UNMODIFIED sources against this header instead: a tiny value
// tree (int / float / bool / str / list) with the handful of Boost.Python names those sources use.  Nothing here takes
// part in the encoder's arithmetic; it only carries the inputs in and the uint8 array out.
#ifndef HELLO_ORACLE_BOOST_PYTHON_SHIM
#define HELLO_ORACLE_BOOST_PYTHON_SHIM
#include <algorithm>        // the real Boost headers pull these in transitively and the reference relies on that
#include <cstddef>
#include <iostream>
#include <iterator>
#include <map>
#include <set>
#include <sstream>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace boost { namespace python {

class object;
struct shim_node {
    enum Kind { NONE, INT, FLOAT, BOOL, STR, LIST } kind = NONE;
    long long i = 0;
    double f = 0;
    std::string s;
    std::vector<object> items;
};

class object {
public:
    std::shared_ptr<shim_node> n;
    object() : n(std::make_shared<shim_node>()) {}
    template <class T, typename std::enable_if<std::is_integral<T>::value && !std::is_same<T, bool>::value, int>::type = 0>
    object(T v) : n(std::make_shared<shim_node>()) { n->kind = shim_node::INT; n->i = (long long)v; }
    object(bool v) : n(std::make_shared<shim_node>()) { n->kind = shim_node::BOOL; n->i = v; }
    object(double v) : n(std::make_shared<shim_node>()) { n->kind = shim_node::FLOAT; n->f = v; }
    object(const char* v) : n(std::make_shared<shim_node>()) { n->kind = shim_node::STR; n->s = v; }
    object(const std::string& v) : n(std::make_shared<shim_node>()) { n->kind = shim_node::STR; n->s = v; }
    object operator[](std::size_t k) const;
};

inline object object::operator[](std::size_t k) const {
    if (n->kind != shim_node::LIST || k >= n->items.size()) throw std::out_of_range("shim list index");
    return n->items[k];
}

class list : public object {
public:
    list() { n->kind = shim_node::LIST; }
    list(const object& o) : object(o) { if (n->kind != shim_node::LIST) throw std::invalid_argument("shim: not a list"); }
    template <class T> void append(const T& v) { n->items.push_back(object(v)); }
    void append(const object& v) { n->items.push_back(v); }
};

class tuple : public list {
public:
    tuple() {}
    tuple(const object& o) : list(o) {}
};

template <class A> tuple make_tuple(const A& a) { tuple t; t.append(a); return t; }
template <class A, class B> tuple make_tuple(const A& a, const B& b) { tuple t; t.append(a); t.append(b); return t; }
template <class A, class B, class C> tuple make_tuple(const A& a, const B& b, const C& c) {
    tuple t; t.append(a); t.append(b); t.append(c); return t;
}

inline long len(const object& o) {
    if (o.n->kind == shim_node::LIST) return (long)o.n->items.size();
    if (o.n->kind == shim_node::STR) return (long)o.n->s.size();
    throw std::invalid_argument("shim: len() of a scalar");
}

template <class T, class Enable = void> struct shim_convert {
    static T get(const object&) { throw std::invalid_argument("shim: extract<> of an unsupported type"); }
};
template <class T> struct shim_convert<T, typename std::enable_if<std::is_arithmetic<T>::value>::type> {
    static T get(const object& o) {
        switch (o.n->kind) {
            case shim_node::INT: case shim_node::BOOL: return (T)o.n->i;
            case shim_node::FLOAT: return (T)o.n->f;
            default: throw std::invalid_argument("shim: not a number");
        }
    }
};
template <> struct shim_convert<std::string> {
    static std::string get(const object& o) {
        if (o.n->kind != shim_node::STR) throw std::invalid_argument("shim: not a string");
        return o.n->s;
    }
};

template <class T> struct extract {
    object o;
    extract(const object& o_) : o(o_) {}
    operator T() const { return shim_convert<T>::get(o); }
    T operator()() const { return shim_convert<T>::get(o); }
};

template <class T> struct stl_input_iterator {
    using iterator_category = std::input_iterator_tag;
    using value_type = T;
    using difference_type = std::ptrdiff_t;
    using pointer = const T*;
    using reference = T;
    object o;
    std::size_t k = 0;
    bool end = true;
    stl_input_iterator() {}
    stl_input_iterator(const object& o_) : o(o_), k(0), end(false) { end = (std::size_t)len(o) == 0; }
    T operator*() const { return shim_convert<T>::get(o[k]); }
    stl_input_iterator& operator++() { ++k; if (k >= (std::size_t)len(o)) end = true; return *this; }
    stl_input_iterator operator++(int) { stl_input_iterator t = *this; ++*this; return t; }
    bool operator==(const stl_input_iterator& b) const { return end == b.end && (end || k == b.k); }
    bool operator!=(const stl_input_iterator& b) const { return !(*this == b); }
};

}}  // namespace boost::python
#endif
