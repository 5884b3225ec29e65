/* Stand-in for <opencv2/highgui/highgui.hpp> (absent third-party header): see amb_refsrc_deps.h.  TEST INFRASTRUCTURE. */
#include <amb_refsrc_deps.h>
