#include <boost/log/common.hpp>
