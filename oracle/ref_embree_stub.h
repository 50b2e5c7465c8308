// Scene handle of the Embree stand-in (see ref_embree_stub.cpp).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <vector>
#include "refshim/embree4/rtcore.h"
#include "../include/ospcuda.h"

struct RTCSceneTy
{
  const OcWorld *world = nullptr;             // HOST pointers
  bool volumes = false;                       // the world's volume scene (user geometries = volumes)
  std::vector<std::vector<void *>> userPtr;   // [instance][model | volume] -> geometryUserPtr for user-geometry callbacks
};
