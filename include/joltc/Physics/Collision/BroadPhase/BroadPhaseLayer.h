/* Forwarding header: the engine's `#include <joltc/Physics/Collision/BroadPhase/BroadPhaseLayer.h>` resolves to the libgpx-backed subset. */
#include "../../../../joltc_gpx.h"
