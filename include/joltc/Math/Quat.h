/* Forwarding header: the engine's `#include <joltc/Math/Quat.h>` resolves to the libgpx-backed subset. */
#include "../../joltc_gpx.h"
