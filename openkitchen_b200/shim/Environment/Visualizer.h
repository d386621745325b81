// forwarding header: the reference's Environment/Visualizer.h surface lives in OpenKitchenB200.hpp
#pragma once
#include "../OpenKitchenB200.hpp"
