// Everything libiqo's users include with <libiqo/iqo.hpp> (reference: include/libiqo/iqo.hpp).
#ifndef LIBIQO_IQO_HPP
#define LIBIQO_IQO_HPP

#include "LanczosResizer.hpp"
#include "AreaResizer.hpp"
#include "LinearResizer.hpp"

#endif
