/* Forwarding header: the engine's `#include <joltc/constants.h>` resolves to the libgpx-backed subset. */
#include "../joltc_gpx.h"
