/* Forwarding header: the engine's `#include <joltc/Physics/Body/Body.h>` resolves to the libgpx-backed subset. */
#include "../../../joltc_gpx.h"
