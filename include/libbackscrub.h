/* include/libbackscrub.h — same interface under the flat name (out-of-tree users that say
 * `#include "libbackscrub.h"`); see include/lib/libbackscrub.h. */
#include "lib/libbackscrub.h"
