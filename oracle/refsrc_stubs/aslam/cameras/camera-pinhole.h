/* Stand-in for <aslam/cameras/camera-pinhole.h> (absent third-party header): see amb_refsrc_deps.h.  TEST INFRASTRUCTURE. */
#include <amb_refsrc_deps.h>
