/* Forwarding header: the engine's `#include <joltc/types.h>` resolves to the libgpx-backed subset. */
#include "../joltc_gpx.h"
