// stand-in (see xtensor/containers/xarray.hpp); intentionally empty
#pragma once
#include <xtensor/containers/xarray.hpp>
