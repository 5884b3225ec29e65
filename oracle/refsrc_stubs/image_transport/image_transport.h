/* Stand-in for <image_transport/image_transport.h> (absent third-party header): see amb_refsrc_stereo_deps.h.  TEST INFRASTRUCTURE. */
#include <amb_refsrc_stereo_deps.h>
