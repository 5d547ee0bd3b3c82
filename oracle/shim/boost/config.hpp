// oracle/shim: minimal stand-in for <boost/config.hpp> (TEST INFRASTRUCTURE ONLY).
// The reference relies on transitive std includes that real Boost drags in.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
namespace boost {}
