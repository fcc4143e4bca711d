// TEST INFRASTRUCTURE ONLY.  INTEGRATION.md's three-line patch, applied through the include path: when
// oracle/shim_map_b200 precedes the reference's source directory, surfel_map.h's `#include <fusion_functions.h>`
// lands here and the reference's SurfelMap holds the product's adapter instead of its own FusionFunctions.
#pragma once
#include <Eigen/Eigen>
#include <opencv2/opencv.hpp>
#include <elements.h> // the reference's SurfelElement / Superpixel_seed
#include "dsm_fusion_functions.hpp"
typedef dsm::FusionFunctions FusionFunctions;
