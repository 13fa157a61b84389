/* Forwarding header: the engine's `#include <joltc/Math/Vector3.h>` resolves to the libgpx-backed subset. */
#include "../../joltc_gpx.h"
