#include <boost/python.hpp>
