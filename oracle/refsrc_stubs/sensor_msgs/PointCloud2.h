/* Stand-in for <sensor_msgs/PointCloud2.h> (absent third-party header): see amb_refsrc_stereo_deps.h.  TEST INFRASTRUCTURE. */
#include <amb_refsrc_stereo_deps.h>
