// forwarding header: a case file written against the reference includes "Rectangle.hpp"; all declarations live in veritas_host.hpp
#include "veritas_host.hpp"
