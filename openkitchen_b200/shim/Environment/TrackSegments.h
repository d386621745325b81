// forwarding header: the reference's Environment/TrackSegments.h surface lives in OpenKitchenB200.hpp
#pragma once
#include "../OpenKitchenB200.hpp"
