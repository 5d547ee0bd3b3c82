/* include/compat/hector.hpp -- the reference's umbrella header (inst/include/hector.hpp) for code
 * that is compiled, UNMODIFIED, against the B200 engine instead of libhector.a:
 *
 *   g++ -I include/compat -I include  src/main.cpp  -L hector_b200 -lhector_b200
 *
 * Every header in this directory carries the name of the reference header it stands in for and
 * forwards to the C++ facade include/hector_b200_core.hpp with `namespace Hector` as an alias of
 * `namespace hector_b200`.  Callers covered: src/main.cpp (the CLI), src/rcpp_hector.cpp (the R
 * glue; needs an Rcpp.h), misc/main-api.cpp-style embedding code.  See INTEGRATION.md. */
#ifndef HECTOR_B200_COMPAT_HECTOR_HPP
#define HECTOR_B200_COMPAT_HECTOR_HPP
#include "avisitor.hpp"
#include "component_data.hpp"
#include "core.hpp"
#include "csv_outputstream_visitor.hpp"
#include "csv_tracking_visitor.hpp"
#include "h_exception.hpp"
#include "h_reader.hpp"
#include "h_util.hpp"
#include "ini_to_core_reader.hpp"
#include "logger.hpp"
#include "message_data.hpp"
#include "unitval.hpp"
#endif
