// Stand-in for Boost.Log: the reference only writes debug lines; they are discarded.
#ifndef HELLO_ORACLE_BOOST_LOG_SHIM
#define HELLO_ORACLE_BOOST_LOG_SHIM
#include <ostream>
namespace boost { namespace log {
struct shim_null_stream {
    template <class T> shim_null_stream& operator<<(const T&) { return *this; }
    shim_null_stream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
namespace sources { template <class L> struct severity_logger {}; }
namespace sinks {}
}}
#define BOOST_LOG_SEV(lg, sev) ::boost::log::shim_null_stream()
#endif
