// forwarding header: the reference's Environment/RaceTrack.h surface lives in OpenKitchenB200.hpp
#pragma once
#include "../OpenKitchenB200.hpp"
