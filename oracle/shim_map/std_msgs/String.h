// TEST INFRASTRUCTURE ONLY: forwards to the single stand-in header (see dsm_ros_shim.hpp)
#pragma once
#include <dsm_ros_shim.hpp>
