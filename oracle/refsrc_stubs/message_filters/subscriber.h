/* Stand-in for <message_filters/subscriber.h> (absent third-party header): see amb_refsrc_stereo_deps.h.  TEST INFRASTRUCTURE. */
#include <amb_refsrc_stereo_deps.h>
