/* Forwarding header: the engine's `#include <joltc/Physics/Collision/NarrowPhaseQuery.h>` resolves to the libgpx-backed subset. */
#include "../../../joltc_gpx.h"
