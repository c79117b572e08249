// Stand-in for <boost/python/numpy.hpp> (see boost/python.hpp in this directory): a zero-initialised byte buffer with a shape.
#ifndef HELLO_ORACLE_BOOST_NUMPY_SHIM
#define HELLO_ORACLE_BOOST_NUMPY_SHIM
#include <boost/python.hpp>
namespace boost { namespace python { namespace numpy {
struct dtype {
    std::size_t itemsize = 1;
    template <class T> static dtype get_builtin() { dtype d; d.itemsize = sizeof(T); return d; }
};
class ndarray {
public:
    std::shared_ptr<std::vector<char>> buf;
    std::vector<long> dims;
    char* get_data() const { return buf->data(); }
    long shape(int k) const { return dims[k]; }
    int get_nd() const { return (int)dims.size(); }
};
inline ndarray zeros(const tuple& shape, const dtype& dt) {
    ndarray a;
    std::size_t total = dt.itemsize;
    for (long k = 0; k < len(shape); ++k) { long d = extract<long>(shape[k]); a.dims.push_back(d); total *= (std::size_t)d; }
    a.buf = std::make_shared<std::vector<char>>(total, 0);
    return a;
}
inline void initialize() {}
}}}
#endif
