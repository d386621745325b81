// forwarding header: the reference's Environment/CollisionChecker.h surface lives in OpenKitchenB200.hpp
#pragma once
#include "../OpenKitchenB200.hpp"
