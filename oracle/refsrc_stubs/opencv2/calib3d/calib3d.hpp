/* Stand-in for <opencv2/calib3d/calib3d.hpp> (absent third-party header): see amb_refsrc_stereo_deps.h.  TEST INFRASTRUCTURE. */
#include <amb_refsrc_stereo_deps.h>
