/* Stand-in for <cv_bridge/cv_bridge.h> (absent third-party header): see amb_refsrc_stereo_deps.h.  TEST INFRASTRUCTURE. */
#include <amb_refsrc_stereo_deps.h>
