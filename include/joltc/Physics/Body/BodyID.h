/* Forwarding header: the engine's `#include <joltc/Physics/Body/BodyID.h>` resolves to the libgpx-backed subset. */
#include "../../../joltc_gpx.h"
