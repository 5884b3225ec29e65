/* Stand-in for <grid_map_msgs/GridMap.h> (absent third-party header): see amb_refsrc_deps.h.  TEST INFRASTRUCTURE. */
#include <amb_refsrc_deps.h>
