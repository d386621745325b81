// forwarding header: the reference's Environment/ScreenGrabber.h surface lives in OpenKitchenB200.hpp
#pragma once
#include "../OpenKitchenB200.hpp"
