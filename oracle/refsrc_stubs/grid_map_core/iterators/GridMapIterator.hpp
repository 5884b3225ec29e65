/* Stand-in for <grid_map_core/iterators/GridMapIterator.hpp> (absent third-party header): see amb_refsrc_deps.h.  TEST INFRASTRUCTURE. */
#include <amb_refsrc_deps.h>
