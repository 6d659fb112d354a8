#pragma once

//! @file
//! @brief umbrella header (same name as the reference's include/libiqo/iqo.hpp)

#include "AreaResizer.hpp"
#include "LinearResizer.hpp"
#include "LanczosResizer.hpp"
