// TEST INFRASTRUCTURE ONLY. boost::shared_ptr stand-in (Utils.h:17-19 aliases std::tr1 to boost).
#pragma once
#include <memory>
namespace boost { using std::shared_ptr; }
