/* Forwarding header: the engine's `#include <joltc/Math/Transform.h>` resolves to the libgpx-backed subset. */
#include "../../joltc_gpx.h"
