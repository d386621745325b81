// forwarding header: the reference's Environment/Agent.h surface lives in OpenKitchenB200.hpp
#pragma once
#include "../OpenKitchenB200.hpp"
