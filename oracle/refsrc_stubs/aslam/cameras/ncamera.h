/* Stand-in for <aslam/cameras/ncamera.h> (absent third-party header): see amb_refsrc_deps.h.  TEST INFRASTRUCTURE. */
#include <amb_refsrc_deps.h>
